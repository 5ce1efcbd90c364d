/* oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE; never linked into the product). */
#ifndef MARSHMALLOW_ORACLE_H
#define MARSHMALLOW_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { OM_TEX_PLACEMENT = 0, OM_TEX_NIGHTSKY = 1, OM_TEX_CURL = 2, OM_TEX_LOWRES = 3, OM_TEX_HIRES = 4 };
enum { OM_FILTER_FP32 = 0, OM_FILTER_FIX8 = 1, OM_FILTER_TEXUNIT = 2 /* bit-exact model of the B200 texture unit */ };
enum { OM_POW_DET = 0, OM_POW_LIBM = 1 };
/* arithmetic definition: one rounding per operator (no contraction), or the lexical fused-multiply-add rule of glsl_env_fma.h */
enum { OM_ARITH_IEEE = 0, OM_ARITH_FMA = 1 };
enum { OM_FULL = 0, OM_PHASE16 = 1 };

typedef struct om_scene om_scene;

om_scene *om_scene_create(void);
void om_scene_destroy(om_scene *s);
int om_scene_set_texture(om_scene *s, int slot, const uint8_t *rgba8, int w, int h, int d);
int om_scene_set_uniforms(om_scene *s, const void *camera160, const void *sun116, const void *sky52);
int om_scene_set_modes(om_scene *s, int filter, int pow_mode);
int om_scene_set_arith(om_scene *s, int arith);
float om_det_powf_fma(float x, float y);   /* the contracted definition's deterministic pow (Horner steps fused) */
int om_march(const om_scene *s, int mode, int W, int H, int row_begin, int row_stride, int row_block,
             float *out_rgba32f, uint32_t *counters /* 4 per pixel or NULL */, int nthreads);
void om_set_window(int trips_per_window /* 1 = plain loop; >1 = windowed replay model of the kernel's ray-split mode */);
void om_set_litmask_buffer(uint32_t *buf /* 8 x uint32 per pixel (zeroed by the caller), or NULL */);
int om_sample(const om_scene *s, int slot, int filter, const float *uvw, int n, float *out_rgba);
typedef struct { const om_scene *scene; int filter; } om_sampler_ctx;
void om_sample_callback(void *user /* om_sampler_ctx* */, int slot, const float *uvw, float *out_rgba);
void om_tonemap_rgba8(const float *rgba32f, size_t npix, uint8_t *rgba8);

/* helper known-answer hooks (compute-clouds.comp:65-77,147-177,193-208) */
float om_det_powf(float x, float y);
float om_hgPhase(float cosTheta, float g);
float om_remap(float v, float a, float b, float c, float d);
float om_remapClamped(float v, float a, float b, float c, float d);
float om_cloudLayerDensity(float relativeHeight, float cloudType);
float om_heightBiasCoverage(float coverage, float height);
int om_raySphereIntersection(const float ro[3], const float rd[3], const float sphere[4], float *t);

/* cloud shadow march of the mesh shader (model.frag:240-283) for n world positions; fetches: texture() calls per point or NULL */
int om_cloud_shadow(const om_scene *s, const float *positions_xyz, int n, float *out_density, uint32_t *fetches, int nthreads);

/* post_chain_oracle.c: god-ray.frag:41-76, radialBlur.frag:36-63, tonemap.frag:11-33 */
void om_sun_screen_position(const void *camera160, const void *sun116, float out_xy[2]);
int om_god_ray(const void *camera160, const void *sun116, const float *src_rgba32f, int W, int H, float *dst_rgba32f);
int om_radial_blur(const void *camera160, const void *sun116, const float *src_rgba32f, int W, int H, float *dst_rgba32f);
int om_tonemap_present(const float *src_rgba32f, int W, int H, int bgra, uint8_t *dst_8888);

/* reproject_oracle.c: restatement of reproject.comp:91-152 */
int om_reproject(const void *camera160, const void *cameraPrev160, const float *src_rgba32f, int W, int H, float *dst_rgba32f);

/* curl_noise_oracle.c: restatement of ImageUtils.cpp:25-223 */
void om_generate_curl_noise(uint8_t *rgba8_128x128);
float om_curl_hash(float x, float y, float z);           /* ImageUtils.cpp:25-29 */
int om_curl_hash_index(float x, float y, float z);      /* ImageUtils.cpp:31-34: index into the 12 gradients */

/* noise_volume_oracle.c: CPU statement of OUR hashed-cell generator (no reference code exists) */
void om_build_noise_volumes(uint64_t seed, uint8_t *low128_rgba8, uint8_t *hi32_rgba8);
uint32_t om_noise_hash(uint32_t x, uint32_t y, uint32_t z, uint32_t seed);

#ifdef __cplusplus
}
#endif
#endif
