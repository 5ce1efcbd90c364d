// uniform_blocks.h -- byte layouts of the three uniform blocks the cloud pass consumes.
// They mirror the reference structs field for field so that an engine can memcpy its own
// UniformCameraObject / UniformSunObject / UniformSkyObject straight into mm_set_uniforms:
//   UniformCameraObject  Shader.h:24-29      (160 bytes)
//   UniformSunObject     SkyManager.h:8-14   (116 bytes)
//   UniformSkyObject     SkyManager.h:28-36  ( 52 bytes)
// Matrices are column-major (glm): element [c][r] lives at float index 4*c + r.
#pragma once
#include <cstddef>

#ifndef MM_CXX_API
#define MM_CXX_API __attribute__((visibility("default")))
#endif

namespace marshmallow {

struct UniformCameraObject {
    float view[16];
    float proj[16];
    float cameraPosition[4];
    float cameraParams[4];   // x = aspect, y = tan(fov/2)
};

struct UniformSunObject {
    float location[4];
    float direction[4];
    float color[4];          // .a carries the 0..15 pixel phase (VulkanApplication.cpp:384)
    float directionBasis[16];
    float intensity;
};

struct UniformSkyObject {
    float betaR[4];
    float betaV[4];
    float wind[4];           // .w = time
    float mie_directional;
};

static_assert(sizeof(UniformCameraObject) == 160, "camera block must be 160 bytes");
static_assert(sizeof(UniformSunObject) == 116, "sun block must be 116 bytes");
static_assert(sizeof(UniformSkyObject) == 52, "sky block must be 52 bytes");
static_assert(offsetof(UniformCameraObject, cameraPosition) == 128, "");
static_assert(offsetof(UniformCameraObject, cameraParams) == 144, "");
static_assert(offsetof(UniformSunObject, directionBasis) == 48, "");
static_assert(offsetof(UniformSunObject, intensity) == 112, "");
static_assert(offsetof(UniformSkyObject, wind) == 32, "");
static_assert(offsetof(UniformSkyObject, mie_directional) == 48, "");

}  // namespace marshmallow
