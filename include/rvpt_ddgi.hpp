// rvpt_ddgi.hpp — C++ host side above the C-ABI (include/ddgi.h): the probe-field part of the
// reference's `class RVPT` (src/rvpt/rvpt.h:33-92) and its `Camera` (src/rvpt/camera.{h,cpp}) with
// the same member names, defaults and call order, backed by libddgi_b200.so instead of the two
// Vulkan compute dispatches.  The reference's host is C++17 / glm / Vulkan; glm, GLFW and the
// Vulkan SDK are not in this image, so the few glm operations the path needs (translate, rotate,
// radians: glm 0.9.9.8, pinned at external/CMakeLists.txt:21-25) are restated here from their
// published definitions.  Header-only; link with -lddgi_b200.
//
//   ddgi::Window-less main loop, as src/rvpt/main.cpp:37-96:
//     ddgi::RVPT rvpt(1600, 900);
//     rvpt.generate_probe_rays();
//     if (!rvpt.initialize()) return 1;
//     while (...) { rvpt.update(); rvpt.draw(); }
//     rvpt.shutdown();
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ddgi.h"

namespace ddgi {

struct vec3 {
    float x = 0, y = 0, z = 0;
};
struct vec4 {
    float x = 0, y = 0, z = 0, w = 0;
};
struct mat4 {
    vec4 c[4];  // columns, as glm::mat4
};

namespace detail {
inline vec4 mul(vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4 add(vec4 a, vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline mat4 identity()
{
    mat4 m;
    m.c[0].x = m.c[1].y = m.c[2].z = m.c[3].w = 1.f;
    return m;
}
// glm::radians: degrees * pi / 180 with the constant folded in fp32
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
// glm::translate(m, v): Result[3] = m[0]*v.x + m[1]*v.y + m[2]*v.z + m[3]
inline mat4 translate(mat4 m, vec3 v)
{
    mat4 r = m;
    r.c[3] = add(add(add(mul(m.c[0], v.x), mul(m.c[1], v.y)), mul(m.c[2], v.z)), m.c[3]);
    return r;
}
// glm::rotate(m, angle, axis) for a unit axis (matrix_transform.inl)
inline mat4 rotate(mat4 m, float angle, vec3 axis)
{
    float c = std::cos(angle), s = std::sin(angle);
    vec3 t = {(1.f - c) * axis.x, (1.f - c) * axis.y, (1.f - c) * axis.z};
    float r00 = c + t.x * axis.x, r01 = t.x * axis.y + s * axis.z, r02 = t.x * axis.z - s * axis.y;
    float r10 = t.y * axis.x - s * axis.z, r11 = c + t.y * axis.y, r12 = t.y * axis.z + s * axis.x;
    float r20 = t.z * axis.x + s * axis.y, r21 = t.z * axis.y - s * axis.x, r22 = c + t.z * axis.z;
    mat4 r;
    r.c[0] = add(add(mul(m.c[0], r00), mul(m.c[1], r01)), mul(m.c[2], r02));
    r.c[1] = add(add(mul(m.c[0], r10), mul(m.c[1], r11)), mul(m.c[2], r12));
    r.c[2] = add(add(mul(m.c[0], r20), mul(m.c[1], r21)), mul(m.c[2], r22));
    r.c[3] = m.c[3];
    return r;
}
}  // namespace detail

// construct_camera_matrix, src/rvpt/camera.cpp:18-26: T * R_y(rot.x) * R_x(rot.y) * R_z(rot.z)
inline mat4 construct_camera_matrix(vec3 translation, vec3 rotation)
{
    mat4 m = detail::identity();
    m = detail::translate(m, translation);
    m = detail::rotate(m, detail::radians(rotation.x), {0, 1, 0});
    m = detail::rotate(m, detail::radians(rotation.y), {1, 0, 0});
    m = detail::rotate(m, detail::radians(rotation.z), {0, 0, 1});
    return m;
}

// src/rvpt/camera.h: only what the path consumes (get_data) plus the setters that feed it.
class Camera {
public:
    explicit Camera(float aspect, vec3 origin = {1.5f, 2.f, -2.f}, vec3 rotation = {-38.f, 36.f, 0.f})
        : aspect(aspect), translation(origin), rotation(rotation)
    {
    }
    void set_fov(float in_fov) { fov = in_fov; }
    float get_fov() const { return fov; }
    void set_scale(float in_scale) { scale = in_scale; }
    float get_scale() const { return scale; }
    // Camera::get_data, src/rvpt/camera.cpp:100-111: 4 matrix columns + (aspect, radians(fov), scale, 0)
    std::vector<vec4> get_data() const
    {
        mat4 m = construct_camera_matrix(translation, rotation);
        return {m.c[0], m.c[1], m.c[2], m.c[3], vec4{aspect, detail::radians(fov), scale, 0.f}};
    }
    float aspect;
    vec3 translation, rotation;

private:
    float fov = 75.f, scale = 4.f;
};

// The probe-field part of class RVPT (src/rvpt/rvpt.h:33-92).
class RVPT {
public:
    RVPT(int width, int height, int device = 0)
        : scene_camera(float(width) / float(height)), device_(device)
    {
        render_settings.screen_width = width;
        render_settings.screen_height = height;
    }
    ~RVPT() { shutdown(); }
    RVPT(RVPT const&) = delete;
    RVPT& operator=(RVPT const&) = delete;

    // RVPT::initialize (rvpt.cpp:231-264): one-time set-up; false (with last_error()) on failure,
    // as the reference's returns false when Vulkan cannot be brought up.
    bool initialize()
    {
        if (ctx_) return true;
        if (ddgi_create(&ctx_, device_) != DDGI_OK) {
            error_ = "ddgi_create failed: no sm_100 CUDA device (the engine has no CPU fallback)";
            return false;
        }
        if (!ok(ddgi_set_irradiance_field(ctx_, &ir))) return false;
        // the scene is compiled into the reference's shaders; here it is baked once per `scene`
        if (!bake_scene()) return false;
        if (need_generate_probe_rays) generate_probe_rays();
        return error_.empty();
    }

    // Bakes render_settings.scene over a box that holds the whole scene (Cornell: the 32^3 box of
    // SURVEY.md 8d; cave / house: [-64,64)^3).
    bool bake_scene()
    {
        int32_t dims[3] = {128, 128, 128}, org[3] = {-64, -64, -64};
        if (render_settings.scene == 1) {
            dims[0] = dims[1] = dims[2] = 32;
            org[0] = org[1] = -15;
            org[2] = 0;
        }
        baked_scene_ = render_settings.scene;
        return ok(ddgi_bake_scene(ctx_, render_settings.scene, dims, org));
    }

    // RVPT::generate_probe_rays (rvpt.h:63, rvpt.cpp:1177-1224): the same libc rand() sample set;
    // the device keeps only the per-probe direction table (probe_rays() rebuilds the 48 B/ray list).
    void generate_probe_rays()
    {
        need_generate_probe_rays = true;
        if (!ctx_) return;  // before initialize(), as main.cpp:47 calls it: done there
        if (ok(ddgi_set_irradiance_field(ctx_, &ir)) && ok(ddgi_generate_probe_rays(ctx_, /*reseed*/ 0)))
            need_generate_probe_rays = false;
    }
    std::vector<ddgi_probe_ray> probe_rays()
    {
        std::vector<ddgi_probe_ray> out(ctx_ ? ddgi_num_probe_rays(ctx_) : 0);
        if (!out.empty()) ok(ddgi_get_probe_rays(ctx_, out.data(), out.size()));
        return out;
    }

    // RVPT::update (rvpt.cpp:265-290): wait for the previous frame, time += 2, upload the uniforms.
    bool update()
    {
        if (!ctx_) return false;
        if (!ok(ddgi_sync(ctx_))) return false;  // raytrace_work_fence.wait()
        render_settings.time += 2;
        if (render_settings.scene != baked_scene_ && !bake_scene()) return false;
        if (!ok(ddgi_set_render_settings(ctx_, &render_settings))) return false;
        std::vector<vec4> cam = scene_camera.get_data();
        static_assert(sizeof(vec4) == 16, "vec4 is 4 floats");
        if (!ok(ddgi_set_camera(ctx_, &cam[0].x))) return false;
        if (!ok(ddgi_set_irradiance_field(ctx_, &ir))) return false;  // re-creates the textures on a shape change
        if (need_generate_probe_rays) generate_probe_rays();
        ddgi_light l[DDGI_MAX_LIGHTS], moved[DDGI_MAX_LIGHTS];
        int32_t n = 0;
        if (!ok(ddgi_default_lights(render_settings.scene, l, &n))) return false;
        if (animate_lights) {  // the update_lights() call both shaders have commented out
            if (!ok(ddgi_update_lights(render_settings.scene, render_settings.time, l, n, moved))) return false;
            return ok(ddgi_set_lights(ctx_, n, moved));
        }
        return ok(ddgi_set_lights(ctx_, n, l));
    }

    enum class draw_return { success, swapchain_out_of_date, error };
    // RVPT::draw -> record_compute_command_buffer (rvpt.cpp:372, :1096-1143): the two dispatches.
    draw_return draw()
    {
        if (!ctx_) return draw_return::error;
        if (!ok(ddgi_probe_update(ctx_, stream)) || !ok(ddgi_render_frame(ctx_, stream))) return draw_return::error;
        return draw_return::success;
    }

    // what output_image / the probe textures held (RGBA8)
    std::vector<uint32_t> read_frame()
    {
        std::vector<uint32_t> out(size_t(render_settings.screen_width) * render_settings.screen_height);
        if (ctx_ && !out.empty()) ok(ddgi_read_frame(ctx_, DDGI_FMT_RGBA8, out.data(), out.size() * 4));
        return out;
    }
    std::vector<uint32_t> read_probe_texture(int which = 0)
    {
        int32_t w = 0, h = 0;
        if (!ctx_ || ddgi_probe_texture_size(ctx_, &w, &h) != DDGI_OK) return {};
        std::vector<uint32_t> out(size_t(w) * h);
        if (!out.empty()) ok(ddgi_read_probe_texture(ctx_, which, DDGI_FMT_RGBA8, out.data(), out.size() * 4));
        return out;
    }

    void shutdown()
    {
        if (ctx_) ddgi_destroy(ctx_);
        ctx_ = nullptr;
    }
    const std::string& last_error() const { return error_; }
    ddgi_ctx* context() { return ctx_; }

    Camera scene_camera;
    // RVPT::RenderSettings / RVPT::IrradianceField with the reference's defaults (rvpt.h:70-90)
    ddgi_render_settings render_settings = {1600, 900, 8, 0, 0, 0, 0.f, 0};
    ddgi_irradiance_field ir = {{9, 7, 9}, 11, 0.9f, 20, {0, 0}, {1.4f, 0.f, 1.f}, 1};
    bool animate_lights = false;
    void* stream = nullptr;  // cudaStream_t; nullptr = the default stream

private:
    bool ok(int rc)
    {
        if (rc == DDGI_OK) return true;
        error_ = ctx_ ? ddgi_last_error(ctx_) : "ddgi error";
        std::fprintf(stderr, "ddgi: %s\n", error_.c_str());  // VK_CHECK_RESULT prints; this never aborts
        return false;
    }
    ddgi_ctx* ctx_ = nullptr;
    int device_ = 0;
    int baked_scene_ = -1;
    bool need_generate_probe_rays = true;
    std::string error_;
};

static_assert(sizeof(ddgi_render_settings) == 32, "RVPT::RenderSettings is 32 bytes (rvpt.h:70-80)");
static_assert(sizeof(ddgi_irradiance_field) == 48, "RVPT::IrradianceField is 48 bytes in std140 (rvpt.h:82-90)");
static_assert(sizeof(ddgi_probe_ray) == 48, "ProbeRay is 48 bytes (probe.h:5-20)");

}  // namespace ddgi
