// ref_camera_shim.cpp -- the reference's own Camera (camera.h / camera.cpp from /root/reference; only the backslashes of two #include
// lines are rewritten, oracle/Makefile) producing the camera uniform block (TEST INFRASTRUCTURE, oracle/_ref/libref_host.so).
// The block is assembled from the class exactly as VulkanApplication.cpp:362-368 does (that file needs Vulkan and cannot compile).
#include <cstring>
#define private public                      // test access to m_position / m_yaw / m_pitch: the app sets them through mouse and keys
#include "camera.h"
#undef private

extern "C" int ref_host_camera(const float position[3], float yaw, float pitch, float fov_deg, float aspect_w, float aspect_h,
                               void *camera160_out) {
    Camera cam;
    cam.m_position = glm::vec3(position[0], position[1], position[2]);
    cam.setAspect(aspect_w, aspect_h);
    cam.setFOV(fov_deg);
    cam.m_yaw = yaw; cam.m_pitch = pitch;
    cam.addYaw(0.0f);                                                             // camera.cpp:147-155: forward / right / up from yaw, pitch
    struct { glm::mat4 view, proj; glm::vec4 cameraPosition, cameraParams; } uco = {};
    uco.proj = cam.getProj();                                                     // VulkanApplication.cpp:362-368
    uco.proj[1][1] *= -1;
    uco.view = cam.getView();
    uco.cameraPosition = glm::vec4(cam.getPosition(), 1.0f);
    uco.cameraParams.x = cam.getAspect();
    uco.cameraParams.y = cam.getHTanFov();
    static_assert(sizeof(uco) == 160, "UniformCameraObject");
    memcpy(camera160_out, &uco, 160);
    return 0;
}
