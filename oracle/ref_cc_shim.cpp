// ref_cc_shim.cpp -- runs the reference's own compute-clouds.comp on the CPU (TEST INFRASTRUCTURE, built into oracle/_ref/).
// The shader text is rewritten lexically by glsl_to_cpp.py into _ref/compute_clouds_gen.inc and compiled here inside the GLSL
// environment of glsl_env.h.  Single-threaded (the shader's uniform blocks are plain globals, as in GLSL).
#include "glsl_env.h"

namespace glsl {
static uvec3 gl_GlobalInvocationID;
static const sampler2D cloudPlacement = {0}, nightSkyMap = {1}, curlNoise = {2};       // slots as in oracle.h (OM_TEX_*)
static const sampler3D lowResCloudShape = {3}, hiResCloudShape = {4};
static const image2D resultImage = {0}, resultImagePrev = {1};
#include "_ref/compute_clouds_gen.inc"
}  // namespace glsl

extern "C" {
// One invocation of main() per (gx, gy) pair: the shader writes pixel (4*gx + o%4, 4*gy + o/4), o = int(sun.color.a), of a
// 1920x1080 image (its hard-coded WIDTH/HEIGHT).  out: 1920*1080*4 floats; written: 1920*1080 flags; fetches: {2D, 3D} totals.
int ref_cc_run(const void *camera160, const void *sun116, const void *sky52, glsl::sample_fn sample, void *user,
               const uint32_t *ids_xy, int n, float *out, uint8_t *written, unsigned long long fetches[2]) {
    using namespace glsl;
    static_assert(sizeof(camera) == 160 && sizeof(sun) == 116 && sizeof(sky) == 52, "uniform blocks must match the engine's structs");
    memcpy((void *)&camera, camera160, 160);
    memcpy((void *)&sun, sun116, 116);
    memcpy((void *)&sky, sky52, 52);
    Env &e = env();
    e.sample = sample; e.user = user; e.out = out; e.written = written; e.out_w = 1920; e.out_h = 1080; e.n2d = e.n3d = 0;
    for (int i = 0; i < n; i++) {
        gl_GlobalInvocationID.x = ids_xy[2 * i]; gl_GlobalInvocationID.y = ids_xy[2 * i + 1]; gl_GlobalInvocationID.z = 0;
        main();
    }
    if (fetches) { fetches[0] = e.n2d; fetches[1] = e.n3d; }
    return 0;
}
}
