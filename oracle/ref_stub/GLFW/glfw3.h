// Stand-in for <GLFW/glfw3.h> + <vulkan/vulkan.h> (TEST INFRASTRUCTURE): the few Vulkan names the reference's SkyManager.h mentions in
// inline descriptor-layout helpers that the host-value comparison never calls.  Lets the reference's SkyManager.cpp compile verbatim
// on a machine without Vulkan or GLFW headers (oracle/Makefile -> oracle/_ref/libref_host.so).
#pragma once
#include <cstdint>
typedef uint32_t VkFlags;
typedef VkFlags VkShaderStageFlags;
typedef struct VkSampler_T *VkSampler;
typedef enum VkDescriptorType { VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER = 6 } VkDescriptorType;
enum { VK_SHADER_STAGE_VERTEX_BIT = 0x1, VK_SHADER_STAGE_FRAGMENT_BIT = 0x10, VK_SHADER_STAGE_COMPUTE_BIT = 0x20 };
typedef struct VkDescriptorSetLayoutBinding {
    uint32_t binding;
    VkDescriptorType descriptorType;
    uint32_t descriptorCount;
    VkShaderStageFlags stageFlags;
    const VkSampler *pImmutableSamplers;
} VkDescriptorSetLayoutBinding;
