// vk_stub.h — TEST-ONLY stand-in for the Vulkan / VK:: layer under the reference's frame orchestration, so
// that the reference's OWN text of RVPT::update() (src/rvpt/rvpt.cpp:264-290), RVPT::record_compute_command_buffer()
// (:1096-1143) and RVPT::generate_probe_rays() (:1145-1224) can be compiled where it lies and run against
// libddgi_b200.so (oracle/ref_glsl/build_shim.py).  This is the drop-in of SURVEY.md 8f-3 made concrete: the
// reference's per-frame uploads and its two dispatches are the calls below, and each lands on one C-ABI
// entry point of include/ddgi.h with the reference's own PODs passed through reinterpret_cast unchanged.
//
//   VK::Fence::wait                     (rvpt.cpp:277)   ->  ddgi_sync
//   settings_uniform.copy_to            (rvpt.cpp:282)   ->  ddgi_set_render_settings   (RVPT::RenderSettings, 32 B)
//   camera_uniform.copy_to              (rvpt.cpp:283)   ->  ddgi_set_camera            (std::vector<glm::vec4>, 80 B)
//   probe_buffer.copy_to                (rvpt.cpp:285)   ->  ddgi_set_probe_rays        (std::vector<ProbeRay>, 48 B each)
//   irradiance_field_uniform.copy_to    (rvpt.cpp:287)   ->  ddgi_set_irradiance_field  (RVPT::IrradianceField, 48 B)
//   vkCmdDispatch #1 (rvpt.cpp:1128)                     ->  ddgi_probe_update   (group counts checked against the texture)
//   vkCmdDispatch #2 (rvpt.cpp:1139)                     ->  ddgi_render_frame   (group counts checked against the frame)
// Everything else Vulkan (barriers, pipeline / descriptor binds, command-buffer begin / end) is a no-op:
// a CUDA stream orders the two kernels.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <optional>
#include <vector>

#include <ddgi.h>

struct ShimState {
    ddgi_ctx* ctx = nullptr;
    int dispatches = 0;  // vkCmdDispatch calls since command_buffer.begin()
    int failures = 0;
};
inline ShimState& shim()
{
    static ShimState s;
    return s;
}
inline void shim_check(int rc, const char* what)
{
    if (rc == DDGI_OK) return;
    std::fprintf(stderr, "shim: %s: %s\n", what, shim().ctx ? ddgi_last_error(shim().ctx) : "no context");  // VK_CHECK_RESULT prints
    shim().failures++;
}

// ---- the handful of Vulkan names record_compute_command_buffer() mentions
typedef struct VkCommandBuffer_T* VkCommandBuffer;
typedef void* VkImage;
typedef void* VkPipeline;
typedef void* VkPipelineLayout;
typedef void* VkDescriptorSet;
enum { VK_STRUCTURE_TYPE_IMAGE_MEMORY_BARRIER = 45, VK_IMAGE_LAYOUT_GENERAL = 1, VK_IMAGE_ASPECT_COLOR_BIT = 1,
       VK_ACCESS_SHADER_READ_BIT = 0x20, VK_ACCESS_SHADER_WRITE_BIT = 0x40, VK_PIPELINE_STAGE_COMPUTE_SHADER_BIT = 0x800,
       VK_PIPELINE_BIND_POINT_COMPUTE = 1 };
struct VkImageSubresourceRange {
    uint32_t aspectMask, baseMipLevel, levelCount, baseArrayLayer, layerCount;
};
struct VkImageMemoryBarrier {
    int sType;
    const void* pNext;
    uint32_t srcAccessMask, dstAccessMask;
    int oldLayout, newLayout;
    uint32_t srcQueueFamilyIndex, dstQueueFamilyIndex;
    VkImage image;
    VkImageSubresourceRange subresourceRange;
};
inline void vkCmdPipelineBarrier(VkCommandBuffer, int, int, int, uint32_t, const void*, uint32_t, const void*, uint32_t,
                                 const VkImageMemoryBarrier*)
{
}
inline void vkCmdBindPipeline(VkCommandBuffer, int, VkPipeline) {}
inline void vkCmdBindDescriptorSets(VkCommandBuffer, int, VkPipelineLayout, uint32_t, uint32_t, const VkDescriptorSet*, uint32_t, const uint32_t*) {}
// The two dispatches.  The reference passes ceil(W/16) x ceil(H/16) groups for the probe pass and
// floor(w/16) x floor(h/16) for the pixel pass; the engine derives the same shapes itself, the stub checks.
inline void vkCmdDispatch(VkCommandBuffer, uint32_t gx, uint32_t gy, uint32_t gz)
{
    ShimState& s = shim();
    if (s.dispatches == 0) {
        int32_t w = 0, h = 0;
        ddgi_probe_texture_size(s.ctx, &w, &h);
        if (gx != (uint32_t)std::ceil(w / 16.0f) || gy != (uint32_t)std::ceil(h / 16.0f) || gz != 1) {
            std::fprintf(stderr, "shim: probe dispatch %u x %u for a %d x %d texture\n", gx, gy, w, h);
            s.failures++;
        }
        shim_check(ddgi_probe_update(s.ctx, nullptr), "ddgi_probe_update");
    } else if (s.dispatches == 1) {
        shim_check(ddgi_render_frame(s.ctx, nullptr), "ddgi_render_frame");
    } else {
        std::fprintf(stderr, "shim: unexpected third dispatch\n");
        s.failures++;
    }
    s.dispatches++;
}

namespace VK
{
constexpr int FLAGS_NONE = 0;
struct Fence {
    void wait() { shim_check(ddgi_sync(shim().ctx), "ddgi_sync"); }
    void reset() {}
};
struct Queue {
    uint32_t get_family() const { return 0; }
};
struct CommandBuffer {
    void begin() { shim().dispatches = 0; }
    void end() {}
    VkCommandBuffer get() { return nullptr; }
};
struct DescriptorSet {
    VkDescriptorSet set = nullptr;
};
struct ImageHandle {
    VkImage handle = nullptr;
};
struct Image {
    ImageHandle image;
    uint32_t width = 0, height = 0;
};
struct PipelineBuilder {
    VkPipeline get_pipeline(int) { return nullptr; }
};
// A uniform / storage buffer of the reference: copy_to memcpys the host object into it
// (src/rvpt/vk_util.cpp:1141-1146); here the bytes go to the C-ABI entry point of the buffer's role.
enum class Role { settings, camera, probe_rays, irradiance_field };
struct Buffer {
    Role role;
    explicit Buffer(Role r) : role(r) {}
    void upload(const void* data, size_t bytes)
    {
        ddgi_ctx* c = shim().ctx;
        switch (role) {
            case Role::settings:
                if (bytes != sizeof(ddgi_render_settings)) break;
                return shim_check(ddgi_set_render_settings(c, reinterpret_cast<const ddgi_render_settings*>(data)), "ddgi_set_render_settings");
            case Role::camera:
                if (bytes != 80) break;
                return shim_check(ddgi_set_camera(c, reinterpret_cast<const float*>(data)), "ddgi_set_camera");
            case Role::irradiance_field:
                if (bytes != sizeof(ddgi_irradiance_field)) break;
                return shim_check(ddgi_set_irradiance_field(c, reinterpret_cast<const ddgi_irradiance_field*>(data)), "ddgi_set_irradiance_field");
            case Role::probe_rays:
                if (bytes % sizeof(ddgi_probe_ray) != 0) break;
                return shim_check(ddgi_set_probe_rays(c, reinterpret_cast<const ddgi_probe_ray*>(data), bytes / sizeof(ddgi_probe_ray)), "ddgi_set_probe_rays");
        }
        std::fprintf(stderr, "shim: copy_to of %zu bytes does not fit the buffer's role\n", bytes);
        shim().failures++;
    }
    template <typename T>
    void copy_to(const T& v)
    {
        upload(&v, sizeof(T));
    }
    template <typename T>
    void copy_to(const std::vector<T>& v)
    {
        upload(v.data(), v.size() * sizeof(T));
    }
};
}  // namespace VK
