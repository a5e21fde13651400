// host_main.cpp — TEST program for the C++ host mirror (include/rvpt_ddgi.hpp): drives the C-ABI
// from compiled code the way the reference's main loop drives class RVPT (src/rvpt/main.cpp:37-96).
//   host_main camera <aspect> <ox oy oz> <rx ry rz>      prints Camera::get_data() as 20 hex words (no GPU)
//   host_main frame <scene> <X Y Z> <side> <s> <fx fy fz> <w h> <ox oy oz> <rx ry rz> <frames> <out.bin>
//        generate_probe_rays / initialize / (update, draw) x frames; writes W, H, w, h (int32), the
//        albedo probe texture and the frame (RGBA8)
#include <cstdlib>
#include <cstring>

#include "rvpt_ddgi.hpp"

static uint32_t bits(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}

int main(int argc, char** argv)
{
    if (argc >= 9 && !std::strcmp(argv[1], "camera")) {
        ddgi::Camera cam((float)atof(argv[2]), {(float)atof(argv[3]), (float)atof(argv[4]), (float)atof(argv[5])},
                         {(float)atof(argv[6]), (float)atof(argv[7]), (float)atof(argv[8])});
        for (const ddgi::vec4& v : cam.get_data()) std::printf("%08x %08x %08x %08x\n", bits(v.x), bits(v.y), bits(v.z), bits(v.w));
        return 0;
    }
    if (argc >= 21 && !std::strcmp(argv[1], "frame")) {
        int a = 2;
        int scene = atoi(argv[a++]);
        int X = atoi(argv[a++]), Y = atoi(argv[a++]), Z = atoi(argv[a++]);
        int side = atoi(argv[a++]), s = atoi(argv[a++]);
        float fx = (float)atof(argv[a++]), fy = (float)atof(argv[a++]), fz = (float)atof(argv[a++]);
        int w = atoi(argv[a++]), h = atoi(argv[a++]);
        ddgi::vec3 o = {(float)atof(argv[a]), (float)atof(argv[a + 1]), (float)atof(argv[a + 2])};
        a += 3;
        ddgi::vec3 r = {(float)atof(argv[a]), (float)atof(argv[a + 1]), (float)atof(argv[a + 2])};
        a += 3;
        int frames = atoi(argv[a++]);
        const char* out = argv[a++];

        ddgi::RVPT rvpt(w, h);
        rvpt.scene_camera = ddgi::Camera(float(w) / float(h), o, r);
        rvpt.render_settings.scene = scene;
        rvpt.ir.probe_count[0] = X;
        rvpt.ir.probe_count[1] = Y;
        rvpt.ir.probe_count[2] = Z;
        rvpt.ir.side_length = side;
        rvpt.ir.sqrt_rays_per_probe = s;
        rvpt.ir.field_origin[0] = fx;
        rvpt.ir.field_origin[1] = fy;
        rvpt.ir.field_origin[2] = fz;
        rvpt.generate_probe_rays();  // main.cpp:47: before initialize()
        if (!rvpt.initialize()) {
            std::fprintf(stderr, "initialize failed: %s\n", rvpt.last_error().c_str());
            return 2;
        }
        for (int f = 0; f < frames; f++) {
            if (!rvpt.update() || rvpt.draw() != ddgi::RVPT::draw_return::success) {
                std::fprintf(stderr, "frame %d failed: %s\n", f, rvpt.last_error().c_str());
                return 3;
            }
        }
        ddgi_sync(rvpt.context());
        int32_t W = 0, H = 0;
        ddgi_probe_texture_size(rvpt.context(), &W, &H);
        std::vector<uint32_t> tex = rvpt.read_probe_texture(0), frame = rvpt.read_frame();
        FILE* fp = std::fopen(out, "wb");
        if (!fp) return 4;
        int32_t hdr[4] = {W, H, w, h};
        std::fwrite(hdr, 4, 4, fp);
        std::fwrite(tex.data(), 4, tex.size(), fp);
        std::fwrite(frame.data(), 4, frame.size(), fp);
        std::fclose(fp);
        std::printf("ok %d probe rays, time %.1f, %llu kernel launches\n", (int)ddgi_num_probe_rays(rvpt.context()),
                    rvpt.render_settings.time, (unsigned long long)ddgi_launch_count(rvpt.context()));
        rvpt.shutdown();
        return 0;
    }
    std::fprintf(stderr, "usage: host_main camera ... | frame ...\n");
    return 1;
}
