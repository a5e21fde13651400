// host_main.cpp — TEST program for the C++ host mirror (include/rvpt_ddgi.hpp): drives the C-ABI
// from compiled code the way the reference's main loop drives class RVPT (src/rvpt/main.cpp:37-96).
//   host_main camera <aspect> <ox oy oz> <rx ry rz>      prints Camera::get_data() as 20 hex words (no GPU)
//   host_main frame <scene> <X Y Z> <side> <s> <fx fy fz> <w h> <ox oy oz> <rx ry rz> <frames> <out.bin>
//        generate_probe_rays / initialize / (update, draw) x frames; writes W, H, w, h (int32), the
//        albedo probe texture and the frame (RGBA8)
//   host_main shard <rank> <world> <device> <idfile> <out.bin>
//        one process per GPU, Cornell with 2x4x2 probes: rank 0 writes an ncclUniqueId to <idfile>, every rank
//        joins (ddgi_comm_init), updates its slab of probe rows, exchanges the texture in place over NCCL
//        (ddgi_exchange_allgather) and writes the whole albedo texture; also round-trips a checkpoint file
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

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
        if (const char* fl = std::getenv("DDGI_FRAMES_IN_FLIGHT")) {
            // the reference keeps MAX_FRAMES_IN_FLIGHT = 2 (src/rvpt/rvpt.h:23): updates on the engine's own two streams
            if (ddgi_set_double_buffer(rvpt.context(), 1) != DDGI_OK || ddgi_set_frames_in_flight(rvpt.context(), atoi(fl)) != DDGI_OK) {
                std::fprintf(stderr, "frames in flight: %s\n", ddgi_last_error(rvpt.context()));
                return 2;
            }
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
    if (argc >= 7 && !std::strcmp(argv[1], "shard")) {
        int rank = atoi(argv[2]), world = atoi(argv[3]), device = atoi(argv[4]);
        std::string idfile = argv[5];
        const char* out = argv[6];
        ddgi::RVPT rvpt(64, 64, device);
        rvpt.scene_camera = ddgi::Camera(1.f, {0, 0, -5}, {0, 0, 0});
        rvpt.render_settings.scene = 1;
        rvpt.ir.probe_count[0] = 2;
        rvpt.ir.probe_count[1] = 4;
        rvpt.ir.probe_count[2] = 2;
        rvpt.ir.side_length = 7;
        rvpt.ir.sqrt_rays_per_probe = 8;
        rvpt.ir.field_origin[0] = 0;
        rvpt.ir.field_origin[1] = 0;
        rvpt.ir.field_origin[2] = 15;
        srand(1);
        rvpt.generate_probe_rays();
        if (!rvpt.initialize()) {
            std::fprintf(stderr, "initialize failed: %s\n", rvpt.last_error().c_str());
            return 2;
        }
        ddgi_ctx* ctx = rvpt.context();
        unsigned char id[128];
        if (rank == 0) {
            if (ddgi_comm_unique_id(id) != DDGI_OK) {
                std::fprintf(stderr, "ddgi_comm_unique_id failed (no libnccl.so.2?)\n");
                return 6;
            }
            FILE* fp = std::fopen((idfile + ".tmp").c_str(), "wb");
            if (!fp) return 4;
            std::fwrite(id, 1, 128, fp);
            std::fclose(fp);
            std::rename((idfile + ".tmp").c_str(), idfile.c_str());
        } else {
            FILE* fp = nullptr;
            for (int i = 0; i < 600 && !(fp = std::fopen(idfile.c_str(), "rb")); i++) std::this_thread::sleep_for(std::chrono::milliseconds(100));
            if (!fp || std::fread(id, 1, 128, fp) != 128) return 4;
            std::fclose(fp);
        }
        if (ddgi_comm_init(ctx, id, rank, world) != DDGI_OK) {
            std::fprintf(stderr, "ddgi_comm_init: %s\n", ddgi_last_error(ctx));
            return 6;
        }
        int Y = rvpt.ir.probe_count[1];
        if (ddgi_set_probe_rows(ctx, rank * Y / world, (rank + 1) * Y / world) != DDGI_OK) return 3;
        if (!rvpt.update()) return 3;
        if (ddgi_probe_update(ctx, nullptr) != DDGI_OK || ddgi_exchange_allgather(ctx, nullptr) != DDGI_OK) {
            std::fprintf(stderr, "update / exchange: %s\n", ddgi_last_error(ctx));
            return 3;
        }
        ddgi_sync(ctx);
        std::vector<uint32_t> tex = rvpt.read_probe_texture(0);
        // checkpoint round trip through the C entry points
        std::string ck = std::string(out) + ".ckpt";
        float t = -1.f;
        if (ddgi_save_checkpoint(ctx, ck.c_str(), rvpt.render_settings.time) != DDGI_OK) return 7;
        std::vector<uint32_t> zero(tex.size(), 0u);
        ddgi_write_probe_texture(ctx, 0, zero.data(), zero.size() * 4);
        if (ddgi_load_checkpoint(ctx, ck.c_str(), &t) != DDGI_OK || t != rvpt.render_settings.time) return 7;
        if (rvpt.read_probe_texture(0) != tex) return 7;
        std::remove(ck.c_str());
        int32_t W = 0, H = 0;
        ddgi_probe_texture_size(ctx, &W, &H);
        FILE* fp = std::fopen(out, "wb");
        if (!fp) return 4;
        int32_t hdr[4] = {W, H, rank, world};
        std::fwrite(hdr, 4, 4, fp);
        std::fwrite(tex.data(), 4, tex.size(), fp);
        std::fclose(fp);
        std::printf("ok rank %d of %d, %d probe rays\n", rank, world, (int)ddgi_num_probe_rays(ctx));
        ddgi_comm_destroy(ctx);
        rvpt.shutdown();
        return 0;
    }
    std::fprintf(stderr, "usage: host_main camera ... | frame ... | shard ...\n");
    return 1;
}
