// ref_cc_fma_shim.cpp -- the reference's own compute-clouds.comp on the CPU under the CONTRACTED arithmetic definition (TEST
// INFRASTRUCTURE, built into oracle/_ref/libref_cc_fma.so).  Same lexical rewrite as ref_cc_shim.cpp plus `float` -> `Float`
// (glsl_to_cpp.py --float-class); compiled inside glsl_env_fma.h, whose operators fuse every product that is directly added or
// subtracted.  Single-threaded (the shader's uniform blocks are plain globals, as in GLSL).
#include "glsl_env_fma.h"

namespace glslf {
static uvec3 gl_GlobalInvocationID;
static const sampler2D cloudPlacement = {0}, nightSkyMap = {1}, curlNoise = {2};       // slots as in oracle.h (OM_TEX_*)
static const sampler3D lowResCloudShape = {3}, hiResCloudShape = {4};
static const image2D resultImage = {0}, resultImagePrev = {1};
#include "_ref/compute_clouds_fma_gen.inc"
}  // namespace glslf

extern "C" {
// Same contract as ref_cc_run (ref_cc_shim.cpp): one invocation of main() per (gx, gy) pair of a 1920x1080 image.
int ref_cc_fma_run(const void *camera160, const void *sun116, const void *sky52, glslf::sample_fn sample, void *user,
                   const uint32_t *ids_xy, int n, float *out, uint8_t *written, unsigned long long fetches[2]) {
    using namespace glslf;
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
