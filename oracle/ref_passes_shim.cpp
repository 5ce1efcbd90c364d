// ref_passes_shim.cpp -- the reference's own reproject.comp, god-ray.frag, radialBlur.frag, tonemap.frag and model.frag executed on
// the CPU (TEST INFRASTRUCTURE, built into oracle/_ref/libref_passes.so).  Each shader's text is rewritten lexically by
// glsl_to_cpp.py and compiled in its own namespace inside the GLSL environment of glsl_env.h.  Single-threaded.
#include "glsl_env.h"

namespace glsl {
static uvec3 gl_GlobalInvocationID;
static float probe_value;                                                    // see glsl_to_cpp.py --probe

namespace reproject {
static const image2D targetImage = {0}, sourceImage = {1};
#include "_ref/reproject_gen.inc"
}
namespace godray {
static const fsampler2D texColor = {0};
#include "_ref/god-ray_gen.inc"
}
namespace radialblur {
static const fsampler2D texColor = {0};
#include "_ref/radialBlur_gen.inc"
}
namespace tonemap {
static const fsampler2D texColor = {0};
#include "_ref/tonemap_gen.inc"
}
namespace model {
// material textures of the mesh: constant mid-grey (the shadow march does not read them)
struct csampler2D { int id; };
inline vec4 texture(const csampler2D &, const vec2 &) { return vec4(0.5f, 0.5f, 0.5f, 1.0f); }
static const csampler2D texColor = {0}, pbrInfo = {1}, normalMap = {2};
static const sampler2D cloudPlacement = {0};                                 // OM_TEX_PLACEMENT
static const sampler3D lowResCloudShape = {3};                               // OM_TEX_LOWRES
#include "_ref/model_gen.inc"
}
}  // namespace glsl

using namespace glsl;

template <class T> static void load_block(T &dst, const void *src, size_t have) {
    memset((void *)&dst, 0, sizeof(T));
    memcpy((void *)&dst, src, sizeof(T) < have ? sizeof(T) : have);
}
static void bind_images(const float *src, float *out, int w, int h) {
    Env &e = env();
    e.src = src; e.out = out; e.written = nullptr; e.out_w = w; e.out_h = h;
}

extern "C" {

// reproject.comp: one invocation per pixel of a w x h image
int ref_reproject(const void *camera160, const void *cameraPrev160, const float *src, int w, int h, float *dst) {
    load_block(reproject::camera, camera160, 160);
    load_block(reproject::cameraPrev, cameraPrev160, 160);
    bind_images(src, dst, w, h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            gl_GlobalInvocationID.x = (uint)x; gl_GlobalInvocationID.y = (uint)y; gl_GlobalInvocationID.z = 0;
            reproject::main();
        }
    return 0;
}

// the three post passes: one fragment per pixel, fragUV = pixel centre (the quad's UVs interpolated, Geometry.cpp:94-99)
#define FRAGMENT_PASS(NS, LOADS)                                                                              \
    bind_images(src, nullptr, w, h);                                                                          \
    LOADS                                                                                                     \
    for (int y = 0; y < h; y++)                                                                               \
        for (int x = 0; x < w; x++) {                                                                         \
            NS::fragUV = vec2(((float)x + 0.5f) / (float)w, ((float)y + 0.5f) / (float)h);                    \
            NS::main();                                                                                       \
            float *o = dst + 4 * ((size_t)y * w + x);                                                         \
            o[0] = NS::outColor.x; o[1] = NS::outColor.y; o[2] = NS::outColor.z; o[3] = NS::outColor.w;      \
        }                                                                                                     \
    return 0;

int ref_god_ray(const void *camera160, const void *sun116, const float *src, int w, int h, float *dst) {
    FRAGMENT_PASS(godray, load_block(godray::camera, camera160, 160); load_block(godray::sun, sun116, 116);)
}
int ref_radial_blur(const void *camera160, const void *sun116, const float *src, int w, int h, float *dst) {
    FRAGMENT_PASS(radialblur, load_block(radialblur::camera, camera160, 160); load_block(radialblur::sun, sun116, 116);)
}
int ref_tonemap(const float *src, int w, int h, float *dst) {
    FRAGMENT_PASS(tonemap, ;)
}

// model.frag for n fragments at world positions positions_xyz: only the cloud-shadow march (model.frag:240-283) depends on the
// position; its accumDensity is read back through the probe.  The other stage inputs get fixed plausible values.
int ref_cloud_shadow(const void *camera160, const void *sun116, const void *sky52, sample_fn sample, void *user,
                     const float *positions_xyz, int n, float *out_density, unsigned long long fetches[2]) {
    load_block(model::camera, camera160, 160);
    load_block(model::sun, sun116, 116);
    load_block(model::sky, sky52, 52);
    Env &e = env();
    e.sample = sample; e.user = user; e.n2d = e.n3d = 0;
    model::fragColor = vec3(1.0f); model::fragUV = vec2(0.5f, 0.5f);
    model::fragNormal = vec3(0.0f, 1.0f, 0.0f); model::fragTangent = vec3(1.0f, 0.0f, 0.0f); model::fragBitangent = vec3(0.0f, 0.0f, 1.0f);
    for (int i = 0; i < n; i++) {
        model::fragPositionWC = vec3(positions_xyz[3 * i], positions_xyz[3 * i + 1], positions_xyz[3 * i + 2]);
        model::fragPosition = vec3(0.3f, -0.2f, -4.0f);                      // view-space position: feeds shading and fog only
        probe_value = -1.0f;
        model::main();
        out_density[i] = probe_value;
    }
    if (fetches) { fetches[0] = e.n2d; fetches[1] = e.n3d; }
    return 0;
}

}
