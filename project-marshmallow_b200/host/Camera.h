// Camera.h -- host mirror of the parts of the reference Camera the cloud pass depends on
// (camera.cpp:27-39 getView, camera.cpp:179-195 yaw/pitch frame, camera.h:71-72 aspect / tan(fov/2)).
#pragma once
#include "uniform_blocks.h"

namespace marshmallow {

class MM_CXX_API Camera {
public:
    Camera(const float position[3], float yaw, float pitch, float fovDeg = 45.0f, float aspect = 1920.0f / 1080.0f);
    void getView(float view16[16]) const;
    float getAspect() const { return m_aspect; }
    float getHTanFov() const;
    void fillUniform(UniformCameraObject &uco) const;   // VulkanApplication.cpp:362-369

private:
    float m_position[3], m_forward[3], m_right[3], m_up[3];
    float m_yaw, m_pitch, m_fov, m_aspect;
    void updateFrame();
};

}  // namespace marshmallow
