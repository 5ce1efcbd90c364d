/* oracle_internal.h -- shared between the translation units of the CPU oracle (TEST INFRASTRUCTURE). */
#ifndef MARSHMALLOW_ORACLE_INTERNAL_H
#define MARSHMALLOW_ORACLE_INTERNAL_H
#include <stdint.h>
#include "oracle.h"

typedef struct {
    const float *texels;  /* w*h*d*4 floats holding the byte values 0..255 */
    int w, h, d;
    const uint8_t *bytes; /* the same texels as bytes (integer sampler model) */
} ftex;

struct om_scene {
    ftex placement, nightsky, curl, lowres, hires;
    float *store[5];
    uint8_t *bstore[5];
    float cam[40];   /* UniformCameraObject, 160 B: Shader.h:24-29 */
    float sun[29];   /* UniformSunObject,    116 B: SkyManager.h:8-14 */
    float sky[13];   /* UniformSkyObject,     52 B: SkyManager.h:28-36 */
    int filter;      /* OM_FILTER_* */
    int pow_mode;    /* OM_POW_*    */
    int arith;       /* OM_ARITH_*  */
};

typedef struct { uint32_t trips, n2d, n3d, lit; uint32_t *litmask; /* optional: bit k set = loop iteration k was a lit step (k < 256) */
                 uint8_t *powclass; int in_light; /* diagnostics of tools/pow_filter_bound.py: class of the march trip's coverage pow, 256 trips per pixel */ } px_counters;


/* the software sampler of cloud_march_oracle.c (Texture.cpp:29-52, 315-338 in the three filter definitions) */
void om__sample2d(const ftex *t, int filter, float u, float v, float out[4]);
void om__sample3d(const ftex *t, int filter, float u, float v, float w, float out[4]);
/* CC:288-500 for one pixel under the contracted arithmetic definition (cloud_march_oracle_fma.c) */
void om__march_pixel_fma(const struct om_scene *s, int px, int py, int W, int H, float out[4], px_counters *cnt);

#endif
