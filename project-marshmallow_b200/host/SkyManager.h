// SkyManager.h -- host mirror of the reference's SkyManager (SkyManager.h:51-77): same method
// names and argument meaning, producing the sun / sky uniform blocks of the cloud pass.
#pragma once
#include "uniform_blocks.h"

namespace marshmallow {

class MM_CXX_API SkyManager {
public:
    SkyManager();
    void rebuildSkyFromNewSun(float elevation, float azimuth);
    void rebuildSkyFromScattering(float turbidity, float mie, float mie_directional);
    void setWindDirection(const float dir[3]) { sky.wind[0] = dir[0]; sky.wind[1] = dir[1]; sky.wind[2] = dir[2]; }
    void setTime(float t) { sky.wind[3] = t; }
    void setRayleigh(float r) { rayleigh = r; }
    UniformSunObject &getSun() { return sun; }
    UniformSkyObject getSky() const { return sky; }

private:
    float elevation, azimuth, turbidity, rayleigh, mie;
    UniformSkyObject sky;
    UniformSunObject sun;
    void calcSunPosition();
    void calcSunIntensity();
    void calcSunColor();
    void calcSkyBetaR();
    void calcSkyBetaV();
};

}  // namespace marshmallow
