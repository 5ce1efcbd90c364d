// ref_host_shim.cpp -- the reference's own SkyManager (SkyManager.cpp, compiled verbatim from /root/reference) producing the sun
// and sky uniform blocks (TEST INFRASTRUCTURE, oracle/_ref/libref_host.so).  tests/test_host_values.py compares mm_host_sky with it.
#include <cstring>
#include <new>

#include "SkyManager.h"

extern "C" int ref_host_sky(float elevation, float azimuth, float turbidity, const float wind_xyz[3], float time, int pixel_phase,
                            void *sun116_out, void *sky52_out) {
    static_assert(sizeof(UniformSunObject) == 116 && sizeof(UniformSkyObject) == 52, "uniform block layout");
    // SkyManager never initialises its `turbidity` member (SkyManager.cpp:75-101; rebuildSkyFromScattering ignores its argument,
    // :118-123) yet reads it in calcSkyBetaV (:42-45).  The object is therefore constructed on storage that already holds the
    // intended value in every float slot, so the reference code runs unmodified with a defined turbidity.
    alignas(SkyManager) static unsigned char storage[sizeof(SkyManager)];
    float *slots = reinterpret_cast<float *>(storage);
    for (size_t i = 0; i < sizeof(SkyManager) / sizeof(float); i++) slots[i] = turbidity;
    SkyManager *sm = new (storage) SkyManager();
    sm->rebuildSkyFromNewSun(elevation, azimuth);                                 // VulkanApplication.cpp:373
    sm->setWindDirection(glm::vec3(wind_xyz[0], wind_xyz[1], wind_xyz[2]));
    sm->setTime(time);                                                            // VulkanApplication.cpp:374
    UniformSkyObject sky = sm->getSky();
    UniformSunObject &sun = sm->getSun();
    sun.color.a = (float)pixel_phase;                                             // VulkanApplication.cpp:382
    memcpy(sun116_out, &sun, 116);
    memcpy(sky52_out, &sky, 52);
    sm->~SkyManager();
    return 0;
}
