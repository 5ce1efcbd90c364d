// sky_camera.cpp -- host-side value producers for the uniform blocks (CPU only).
//
// The cloud pass is driven by three uniform blocks that the reference fills every frame in
// VulkanApplication.cpp:351-386 from its SkyManager and Camera classes.  An engine that already
// owns those classes just memcpy's its structs into mm_set_uniforms.  For callers without the
// engine (tests, bench, headless rendering) SkyManager and Camera below produce the same values:
//   SkyManager  <- SkyManager.cpp:15-70 (sun direction/basis/colour/intensity, Rayleigh/Mie betas)
//   Camera      <- camera.cpp:27-39 (view), camera.cpp:179-195 (yaw/pitch frame), camera.h:72
// Promotion rules follow g++ on the reference sources: unqualified cos/sin/exp bind to the double
// overloads, std::cos/std::sin/std::tan on floats stay float, PI is the float 3.14159265f.
// `turbidity` is never initialised in the reference (SkyManager.cpp:42); callers pass it (10 is
// the value of the Three.js sky this code derives from; it scales betaV by ~1e-17 either way).
#include <cmath>
#include <cstring>

#include "../../include/marshmallow.h"
#include "SkyManager.h"
#include "Camera.h"

namespace marshmallow {

namespace {
struct V3 { float x, y, z; };
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 normalize(V3 a) { float inv = 1.0f / std::sqrt(dot(a, a)); return {a.x * inv, a.y * inv, a.z * inv}; }
inline float clampf(float t, float lo, float hi) { return std::fmax(lo, std::fmin(hi, t)); }
const float kPi = 3.14159265f;
}  // namespace

SkyManager::SkyManager() {
    std::memset(&sun, 0, sizeof sun);
    std::memset(&sky, 0, sizeof sky);
    for (int i = 0; i < 4; i++) sun.directionBasis[5 * i] = 1.0f;
    sky.wind[0] = 1.0f; sky.wind[1] = 0.05f; sky.wind[2] = 1.0f; sky.wind[3] = 0.0f;
    sun.color[0] = sun.color[1] = sun.color[2] = 1.0f; sun.color[3] = 0.0f;
    elevation = kPi / 4.f;
    azimuth = kPi / 8.f;
    turbidity = 10.0f;
    mie = 0.005f;
    rayleigh = 2.f;
    sky.mie_directional = 0.8f;
    rebuildSkyFromNewSun(elevation, azimuth);
}

void SkyManager::calcSunPosition() {
    float theta = (float)(2.0 * kPi * (elevation - 0.5));
    float phi = (float)(2.0 * kPi * (azimuth - 0.5));
    V3 dir = {(float)std::cos((double)phi), (float)(std::sin((double)phi) * std::sin((double)theta)),
              (float)(std::sin((double)phi) * std::cos((double)theta))};
    sun.direction[0] = dir.x; sun.direction[1] = dir.y; sun.direction[2] = dir.z; sun.direction[3] = 0.0f;
    const float dist = 400000.0f;
    float sgn = dir.y < 0.0f ? -1.0f : 1.0f;
    sun.location[0] = sgn * (dist * dir.x); sun.location[1] = sgn * (dist * dir.y);
    sun.location[2] = sgn * (dist * dir.z); sun.location[3] = sgn * 1.0f;
    V3 right = (std::fabs(dir.y) < 0.001f) ? ((dir.z < 0) ? V3{0, 1, 0} : V3{0, -1, 0}) : V3{0, 0, 1};
    V3 n = normalize(cross(dir, right));
    V3 b = normalize(cross(dir, n));
    float m[16] = {n.x, n.y, n.z, 0.0f, dir.x, dir.y, dir.z, 0.0f, b.x, b.y, b.z, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f};
    for (int i = 0; i < 16; i++) sun.directionBasis[i] = sgn * m[i];
}

void SkyManager::calcSunIntensity() {
    float c = clampf(sun.direction[1], -1.f, 1.f);
    sun.intensity = 1000.0f * std::fmax(0.f, 1.f - powf(2.718281828459f, -((1.6110731557f - acosf(c)) / 1.5f)));
    if (sun.direction[1] < 0.0f) sun.intensity = 2.0f;
}

void SkyManager::calcSunColor() {
    if (sun.direction[1] < 0.0f) {
        sun.color[0] = 0.8f; sun.color[1] = 0.9f; sun.color[2] = 1.0f;
    } else {
        const float sunset[3] = {2.f, (float)0.33922, 0.0431f};
        float t = clampf(sun.direction[1] * 13.f, 0.f, 1.f);
        for (int i = 0; i < 3; i++) sun.color[i] = (1.f - t) * sunset[i] + t * 1.f;
    }
}

void SkyManager::calcSkyBetaR() {
    float sunFade = (float)(1.0f - (double)clampf((float)(1.0f - std::exp(sun.location[1] / 450000.0)), 0.0f, 1.0f));
    const float total[3] = {(float)5.804542996261093E-6, (float)1.3562911419845635E-5, (float)3.0265902468824876E-5};
    float k = rayleigh - 1.f + sunFade;
    for (int i = 0; i < 3; i++) sky.betaR[i] = total[i] * k;
    sky.betaR[3] = 0.0f;
}

void SkyManager::calcSkyBetaV() {
    float c = (0.2f * turbidity) * 10E-18f;
    const float mieConst[3] = {1.839991851443397f, 2.779802391966052f, 4.079047954386109f};
    for (int i = 0; i < 3; i++) sky.betaV[i] = ((0.434f * c) * mieConst[i]) * mie;
    sky.betaV[3] = 0.0f;
}

void SkyManager::rebuildSkyFromNewSun(float e, float a) {
    elevation = e; azimuth = a;
    calcSunPosition();
    calcSunIntensity();
    calcSunColor();
    calcSkyBetaR();
    calcSkyBetaV();
}

void SkyManager::rebuildSkyFromScattering(float turb, float m, float md) {
    turbidity = turb; mie = m; sky.mie_directional = md;
    calcSkyBetaR();
    calcSkyBetaV();
}

Camera::Camera(const float pos[3], float yaw, float pitch, float fovDeg, float aspect)
    : m_yaw(yaw), m_pitch(pitch), m_fov(fovDeg), m_aspect(aspect) {
    m_position[0] = pos[0]; m_position[1] = pos[1]; m_position[2] = pos[2];
    updateFrame();
}

void Camera::updateFrame() {
    V3 f = {std::cos(m_yaw) * std::cos(m_pitch), std::sin(m_pitch), std::sin(m_yaw) * std::cos(m_pitch)};
    V3 wUp = (1.0f - std::abs(dot(f, V3{0, 1, 0})) < 0.00001f) ? V3{0, 0, 1} : V3{0, 1, 0};
    V3 r = normalize(cross(f, wUp));
    V3 u = normalize(cross(r, f));
    m_forward[0] = f.x; m_forward[1] = f.y; m_forward[2] = f.z;
    m_right[0] = r.x; m_right[1] = r.y; m_right[2] = r.z;
    m_up[0] = u.x; m_up[1] = u.y; m_up[2] = u.z;
}

void Camera::getView(float view[16]) const {
    // camera.cpp:27-39: view = R * T as a 4x4 product in glm's order (column j of the result is
    // ((R0*T[j][0] + R1*T[j][1]) + R2*T[j][2]) + R3*T[j][3]); rows of R are right / up / forward, the camera looks along -forward.
    // Written out as the product, not in closed form, so that the zeros come out with the signs the reference's product gives them.
    float R[16] = {0}, T[16] = {0};
    for (int c = 0; c < 3; c++) { R[4 * c + 0] = m_right[c]; R[4 * c + 1] = m_up[c]; R[4 * c + 2] = m_forward[c]; }
    R[15] = 1.0f;
    T[0] = T[5] = T[10] = T[15] = 1.0f;
    T[12] = -m_position[0]; T[13] = -m_position[1]; T[14] = -m_position[2];
    for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++) {
            volatile float a = R[0 + i] * T[4 * j + 0], b = R[4 + i] * T[4 * j + 1], c = R[8 + i] * T[4 * j + 2], d = R[12 + i] * T[4 * j + 3];
            volatile float s = a + b;
            s = s + c;
            s = s + d;
            view[4 * j + i] = s;
        }
}

float Camera::getHTanFov() const { return std::tan(0.5f * 0.01745f * m_fov); }

void Camera::fillUniform(UniformCameraObject &uco) const {
    std::memset(&uco, 0, sizeof uco);
    getView(uco.view);
    // proj is irrelevant to the cloud pass (never read by compute-clouds.comp); a GL-style perspective
    // with the y flip of VulkanApplication.cpp:364 is filled in for completeness.
    float t = getHTanFov(), n = 0.1f, fa = 100.0f;
    uco.proj[0] = 1.0f / (m_aspect * t); uco.proj[5] = -1.0f / t;
    uco.proj[10] = -(fa + n) / (fa - n); uco.proj[11] = -1.0f; uco.proj[14] = -(2.0f * fa * n) / (fa - n);
    uco.cameraPosition[0] = m_position[0]; uco.cameraPosition[1] = m_position[1]; uco.cameraPosition[2] = m_position[2];
    uco.cameraPosition[3] = 1.0f;
    uco.cameraParams[0] = m_aspect;
    uco.cameraParams[1] = t;
}

}  // namespace marshmallow

extern "C" {

int mm_host_sky(float elevation, float azimuth, float turbidity, float rayleigh, float mie, float mie_directional,
                const float wind_xyz[3], float time, int pixel_phase, void *sun116_out, void *sky52_out) {
    if (!sun116_out || !sky52_out || !wind_xyz) return MM_ERR_ARG;
    marshmallow::SkyManager sm;
    sm.setRayleigh(rayleigh);
    sm.rebuildSkyFromScattering(turbidity, mie, mie_directional);
    sm.rebuildSkyFromNewSun(elevation, azimuth);
    sm.setWindDirection(wind_xyz);
    sm.setTime(time);
    sm.getSun().color[3] = (float)(pixel_phase % 16);
    std::memcpy(sun116_out, &sm.getSun(), 116);
    marshmallow::UniformSkyObject sky = sm.getSky();
    std::memcpy(sky52_out, &sky, 52);
    return MM_OK;
}

int mm_host_camera(const float position[3], float yaw, float pitch, float fov_deg, float aspect, void *camera160_out) {
    if (!position || !camera160_out) return MM_ERR_ARG;
    marshmallow::Camera cam(position, yaw, pitch, fov_deg, aspect);
    marshmallow::UniformCameraObject uco;
    cam.fillUniform(uco);
    std::memcpy(camera160_out, &uco, 160);
    return MM_OK;
}

}  // extern "C"
