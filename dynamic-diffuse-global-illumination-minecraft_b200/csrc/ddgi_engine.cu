// ddgi_engine.cu — host side of libddgi_b200.so: the context that owns the device
// buffers (what class RVPT's per-frame UBOs / SSBO / storage images were,
// src/rvpt/rvpt.h:150-201) and the C-ABI of include/ddgi.h.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/ddgi.h"
#include "ddgi_internal.h"

using namespace ddgi;

struct ddgi_ctx {
    int device = 0;
    char err[512] = {0};
    uint64_t launches = 0;
    int debug = 0;
    int variant = 2;
    int color_mode = 0;  // 0 flat palette, 1 the reference's procedural colours
    int blend_mode = 0;  // 1: hysteresis blend into the previous texel (field.hysteresis)
    int layout = 0, oct = 8;  // probe-texture layout: 0 the reference's ray tile, 1 octahedral oct x oct tiles
    float4* d_ray_out = nullptr;     // octahedral layout: per-ray (radiance, first-hit t)
    size_t ray_out_cap = 0;
    uint32_t* d_owned = nullptr;     // the owned probes, ascending (octahedral blend, packed all-gather)
    uint32_t* d_pack = nullptr;      // gather buffer of the packed all-gather: comm_world chunks of tiles
    size_t pack_cap = 0;
    size_t owned_cap = 0;
    std::vector<uint32_t> owned;
    int weight_mode = 0;    // 1: Chebyshev visibility weight restored in the cage sample
    int distance_mode = 0;  // 1: the probe pass stores first-hit distance moments
    float distance_scale = 1.0f;
    int march_min = 16;  // wavefront kernel: keep stepping while >= march_min/32 of the live lanes march
    int grid_limit = 0;  // wavefront kernel: cap on resident blocks per SM (0 = what the occupancy allows)
    int slot_pref = 32;  // schedule granularity in rays (0 = a whole probe): one warp's fetch measured best
    unsigned long long* d_warp_times = nullptr;  // debug level 2
    size_t warp_times_cap = 0, warp_times_n = 0;
    uint32_t* d_counter = nullptr;   // two ray counters: one per frame in flight
    // Frames in flight (ddgi_set_frames_in_flight): with 2, the update (and exchange) of frame i runs on
    // the engine's own stream frame_stream[i & 1], so the first blocks of update i+1 start while the
    // persistent kernel of update i drains - the reference keeps two frames in flight too
    // (MAX_FRAMES_IN_FLIGHT, src/rvpt/rvpt.h:23).  ev_frame[b]: the frame in texture allocation b is complete.
    int in_flight = 1;
    cudaStream_t frame_stream[2] = {nullptr, nullptr};
    cudaEvent_t ev_frame[2] = {nullptr, nullptr};
    cudaEvent_t ev_in = nullptr;                        // the caller's stream at the time of a dispatch
    cudaEvent_t ev_k0[2] = {nullptr, nullptr}, ev_k1[2] = {nullptr, nullptr};  // around the update kernel (timing)

    ddgi_render_settings rs{};
    ddgi_irradiance_field field{};
    bool have_field = false;
    int rx = 0, ry = 0;
    float cam[20] = {0};
    bool have_cam = false;
    int n_lights = 0;
    Light lights[kMaxLights];

    // voxel field
    int vdim[3] = {0, 0, 0}, vorg[3] = {0, 0, 0}, borg[3] = {0, 0, 0}, nb[3] = {0, 0, 0};
    uint8_t* d_types = nullptr;
    uint32_t* d_occ = nullptr;  // one word per brick of 32 cells (ddgi_scene.cuh: kBrickL*)
    uint8_t* d_edit = nullptr;  // staging buffer of ddgi_edit_voxels
    size_t edit_cap = 0;
    float* d_palette = nullptr;

    // rays
    std::vector<float> samples;  // rx*ry raw sphere samples (xyz)
    // normalised directions, generated mode: a ring of three tables, so that a new sample set (one per
    // frame in bench.py's e2e loop) is uploaded - on the engine's own small upload stream - while the
    // update launched last still reads its own; `ev_dirs[b]` = the last update that read table b
    float* d_dirs = nullptr;     // = d_dirs_ring[dirs_cur]
    float* d_dirs_ring[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_dirs[3] = {nullptr, nullptr, nullptr};
    int dirs_cur = 0;
    size_t dirs_cap = 0;
    cudaStream_t up_stream = nullptr;
    cudaEvent_t ev_update = nullptr;  // recorded after every probe update: what a new ray table must wait for
    float4* d_rays = nullptr;    // literal storage-buffer mode
    size_t n_rays_ssbo = 0;
    int ray_mode = 0;  // 0 none, 1 generated, 2 storage buffer
    bool sample_y_first = false;  // order of the two rand() draws of a sample (ddgi_set_sample_order)
    cudaStream_t last_stream = nullptr;  // stream of the last dispatch: what ddgi_sync waits for

    // probe textures: one allocation, albedo then distance
    int tex_w = 0, tex_h = 0;
    uint32_t* d_tex = nullptr;  // the current texture allocation (= d_tex_pair[cur_tex] under double buffering)
    // double buffering (ddgi_set_double_buffer): probe updates alternate between two allocations so
    // that frame i can be copied out (ddgi_read_probe_texture_async) while frame i+1 is traced
    // The reference stores vec4(0) into the distance image every frame (probe_pass.comp:276,302).
    // While nothing else has written the plane it already holds those zeros, in every replica, and
    // the stores (one local + one per peer and ray) are skipped.  Set by anything that may leave
    // other bytes there; never cleared (a re-created texture starts clean again).
    bool distance_dirty[2] = {false, false};
    bool double_buffer = false;
    uint32_t* d_tex_pair[2] = {nullptr, nullptr};
    int cur_tex = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr};  // last asynchronous read of each buffer
    cudaEvent_t ev_ready = nullptr;                 // what an asynchronous read waits for on the dispatch stream
    float4* d_tex_f32 = nullptr;
    uint32_t* d_ray_lookups = nullptr;
    size_t ray_lookups_cap = 0;  // rays d_ray_lookups was allocated for
    int row0 = 0, row1 = 0;  // probe rows owned by this context (contiguous ownership)
    int cyc_world = 0, cyc_rank = 0, cyc_block = 1;  // block-cyclic ownership when cyc_world > 0
    int cyc_unit = 0;                                // 0: blocks of probe rows, 1: blocks of probes

    // schedule: the owned slots (groups of slot_rays consecutive rays of a probe), most expensive
    // first once calibrated
    std::vector<uint32_t> order;
    uint32_t* d_order = nullptr;
    uint32_t* d_slot_cost = nullptr;
    size_t order_cap = 0;
    bool order_dirty = true;   // ownership / field changed: rebuild the list
    bool calibrated = false;   // per-slot costs of the current scene + rays are in `cost`
    int auto_schedule = 1;
    std::vector<uint32_t> cost;

    // frame
    int frame_w = 0, frame_h = 0;
    uint32_t* d_frame = nullptr;
    float4* d_frame_f32 = nullptr;
    uint32_t* d_px_lookups = nullptr;

    int band_rank = 0, band_world = 1;  // pixel pass: this context renders band `rank` of `world` row bands

    // fused exchange
    uint32_t epoch = 0;              // completion barriers issued since the textures were created
    uint32_t epoch_pre = 0;          // pre-update barriers (single-buffered fused exchange)
    uint32_t* d_barrier_error = nullptr;
    int n_peers = 0, self_index = 0;
    // the peers' texture allocations: [0] the only one, or [0] / [1] the two of a double-buffered
    // context (what ddgi_export_texture_handles listed, in that order); [b][self] is the local one
    void* peer_base[2][kMaxPeers] = {{nullptr}};
    bool peer_opened[2][kMaxPeers] = {{false}};
    int peer_buffers = 0;  // allocations mapped per rank: 1 or 2

    // NCCL exchange (ddgi_comm_init / ddgi_exchange_allgather)
    void* nccl_comm = nullptr;
    int comm_rank = 0, comm_world = 0;
};

// ------------------------------------------------------------------ helpers
static int fail(ddgi_ctx* c, int code, const char* fmt, ...)
{
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(ctx, DDGI_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
#define NEED(cond, ...)                                        \
    do {                                                       \
        if (!(cond)) return fail(ctx, DDGI_E_INVALID, __VA_ARGS__); \
    } while (0)

template <typename T>
static void dfree(T*& p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

static size_t tex_texels(const ddgi_ctx* c) { return (size_t)c->tex_w * c->tex_h; }
static int tile_w(const ddgi_ctx* c) { return c->layout == 1 ? c->oct : c->rx; }
static int tile_h(const ddgi_ctx* c) { return c->layout == 1 ? c->oct : c->ry; }
static size_t num_rays(const ddgi_ctx* c)
{
    if (!c->have_field) return 0;
    return (size_t)c->field.probe_count[0] * c->field.probe_count[1] * c->field.probe_count[2] * c->rx * c->ry;
}

static size_t num_probes(const ddgi_ctx* c)
{
    return (size_t)c->field.probe_count[0] * c->field.probe_count[1] * c->field.probe_count[2];
}

static bool owns_probe(const ddgi_ctx* c, int p)
{
    int per_row = c->field.probe_count[0] * c->field.probe_count[2];
    int y = p / per_row;
    if (c->cyc_world == 0) return y >= c->row0 && y < c->row1;
    int unit = c->cyc_unit ? p : y;
    return (unit / c->cyc_block) % c->cyc_world == c->cyc_rank;
}

// Rays per scheduling slot: ctx->slot_pref when it divides rays/probe, else the whole probe.
static uint32_t slot_rays(const ddgi_ctx* c)
{
    uint32_t rpp = (uint32_t)(c->rx * c->ry);
    uint32_t want = (uint32_t)c->slot_pref;
    return want >= 1u && want <= rpp && rpp % want == 0 ? want : rpp;
}
static size_t num_slots(const ddgi_ctx* c) { return num_probes(c) * ((size_t)(c->rx * c->ry) / slot_rays(c)); }

// Builds the list of owned slots (by the ownership mode: a slab of probe rows, block-cyclic
// rows or block-cyclic probes) and uploads it.  With measured costs the list is sorted most
// expensive first (cost = the largest voxel-lookup count of the slot's rays; ties by index): a
// probe ray's bounces and marches are one long dependent chain, so the longest rays must start
// early or they ARE the kernel's tail, and the last slots taken must hold nothing but short rays.
static int schedule(ddgi_ctx* ctx)
{
    if (!ctx->order_dirty && ctx->d_order) return DDGI_OK;
    for (int b = 0; b < 2; b++)  // (an update in flight on the engine's own streams still reads the old list)
        if (ctx->frame_stream[b]) CU(cudaStreamSynchronize(ctx->frame_stream[b]));
    size_t np = num_probes(ctx), ns = num_slots(ctx);
    uint32_t spp = (uint32_t)(ns / np);
    ctx->order.clear();
    ctx->owned.clear();
    for (size_t p = 0; p < np; p++)
        if (owns_probe(ctx, (int)p)) {
            ctx->owned.push_back((uint32_t)p);
            for (uint32_t j = 0; j < spp; j++) ctx->order.push_back((uint32_t)p * spp + j);
        }
    if (ctx->owned.size() > ctx->owned_cap) {
        dfree(ctx->d_owned);
    dfree(ctx->d_pack);
        CU(cudaMalloc(&ctx->d_owned, ctx->owned.size() * sizeof(uint32_t)));
        ctx->owned_cap = ctx->owned.size();
    }
    if (!ctx->owned.empty())
        CU(cudaMemcpy(ctx->d_owned, ctx->owned.data(), ctx->owned.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (ctx->auto_schedule && ctx->calibrated && ctx->cost.size() == ns) {
        const std::vector<uint32_t>& c = ctx->cost;
        bool measured = true;
        for (uint32_t q : ctx->order) measured = measured && c[q] != 0xffffffffu;
        if (measured)
            std::stable_sort(ctx->order.begin(), ctx->order.end(), [&c](uint32_t a, uint32_t b) { return c[a] > c[b]; });
    }
    if (ns > ctx->order_cap) {
        dfree(ctx->d_order);
        dfree(ctx->d_slot_cost);
        CU(cudaMalloc(&ctx->d_order, ns * sizeof(uint32_t)));
        CU(cudaMalloc(&ctx->d_slot_cost, ns * sizeof(uint32_t)));
        ctx->order_cap = ns;
    }
    if (!ctx->order.empty())
        CU(cudaMemcpy(ctx->d_order, ctx->order.data(), ctx->order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    ctx->order_dirty = false;
    return DDGI_OK;
}

// Per-ray debug buffers are sized by the ray count, which changes with the ray tile even when the
// texture size does not (octahedral layout): they are re-created on demand.
static void free_ray_buffers(ddgi_ctx* ctx)
{
    dfree(ctx->d_ray_lookups);
    ctx->ray_lookups_cap = 0;
}

// (Re)creates the probe textures for the current field, as recreate_probe_textures
// does when probe counts or rays/probe change (src/rvpt/rvpt.cpp:661-755).  The new allocations are
// made before the old ones are released: a failure leaves the context as it was.
static int resize_textures(ddgi_ctx* ctx)
{
    int w = ctx->field.probe_count[0] * ctx->field.probe_count[2] * tile_w(ctx);
    int h = ctx->field.probe_count[1] * tile_h(ctx);
    if (w == ctx->tex_w && h == ctx->tex_h && ctx->d_tex) return DDGI_OK;
    if (ctx->n_peers) return fail(ctx, DDGI_E_STATE, "close peers before resizing the probe textures");
    if (ctx->copy_stream) CU(cudaStreamSynchronize(ctx->copy_stream));
    size_t n = (size_t)w * h;
    // both planes, then the epoch flags of the fused exchange (ddgi_exchange_barrier)
    uint32_t* fresh[2] = {nullptr, nullptr};
    for (int b = 0; b < (ctx->double_buffer ? 2 : 1); b++) {
        cudaError_t e = cudaMalloc(&fresh[b], (2 * n + kFlagWords) * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMemset(fresh[b], 0, (2 * n + kFlagWords) * sizeof(uint32_t));
        if (e != cudaSuccess) {
            dfree(fresh[0]);
            dfree(fresh[1]);
            return fail(ctx, DDGI_E_CUDA, "probe textures %d x %d: %s", w, h, cudaGetErrorString(e));
        }
    }
    dfree(ctx->d_tex_pair[0]);
    dfree(ctx->d_tex_pair[1]);
    dfree(ctx->d_tex_f32);
    free_ray_buffers(ctx);
    ctx->d_tex_pair[0] = fresh[0];
    ctx->d_tex_pair[1] = fresh[1];
    ctx->tex_w = w;
    ctx->tex_h = h;
    ctx->cur_tex = 0;
    ctx->d_tex = ctx->d_tex_pair[0];
    ctx->epoch = ctx->epoch_pre = 0;
    ctx->distance_dirty[0] = ctx->distance_dirty[1] = false;
    return DDGI_OK;
}

static int ensure_debug_buffers(ddgi_ctx* ctx)
{
    if (!ctx->debug) return DDGI_OK;
    size_t n = tex_texels(ctx);
    if (n && !ctx->d_tex_f32) {
        CU(cudaMalloc(&ctx->d_tex_f32, n * sizeof(float4)));
        CU(cudaMemset(ctx->d_tex_f32, 0, n * sizeof(float4)));
    }
    if (n && (!ctx->d_ray_lookups || ctx->ray_lookups_cap != num_rays(ctx))) {
        free_ray_buffers(ctx);
        CU(cudaMalloc(&ctx->d_ray_lookups, num_rays(ctx) * sizeof(uint32_t)));
        CU(cudaMemset(ctx->d_ray_lookups, 0, num_rays(ctx) * sizeof(uint32_t)));
        ctx->ray_lookups_cap = num_rays(ctx);
    }
    size_t px = (size_t)ctx->frame_w * ctx->frame_h;
    if (px && !ctx->d_frame_f32) {
        CU(cudaMalloc(&ctx->d_frame_f32, px * sizeof(float4)));
        CU(cudaMemset(ctx->d_frame_f32, 0, px * sizeof(float4)));
    }
    if (px && !ctx->d_px_lookups) {
        CU(cudaMalloc(&ctx->d_px_lookups, px * sizeof(uint32_t)));
        CU(cudaMemset(ctx->d_px_lookups, 0, px * sizeof(uint32_t)));
    }
    return DDGI_OK;
}

static int resize_frame(ddgi_ctx* ctx)
{
    int w = ctx->rs.screen_width, h = ctx->rs.screen_height;
    if (w == ctx->frame_w && h == ctx->frame_h && ctx->d_frame) return DDGI_OK;
    dfree(ctx->d_frame);
    dfree(ctx->d_frame_f32);
    dfree(ctx->d_px_lookups);
    ctx->frame_w = w;
    ctx->frame_h = h;
    if ((size_t)w * h == 0) return DDGI_OK;
    CU(cudaMalloc(&ctx->d_frame, (size_t)w * h * sizeof(uint32_t)));
    CU(cudaMemset(ctx->d_frame, 0, (size_t)w * h * sizeof(uint32_t)));
    return DDGI_OK;
}

static void fill_params(const ddgi_ctx* c, FrameParams* P)
{
    memset(P, 0, sizeof(*P));
    P->scene.occ = c->d_occ;
    P->scene.types = c->d_types;
    P->scene.palette = c->d_palette;
    P->scene.color_mode = c->color_mode;
    for (int a = 0; a < 3; a++) {
        P->scene.vorg[a] = c->vorg[a];
        P->scene.vdim[a] = c->vdim[a];
        P->scene.nb[a] = c->nb[a];
        P->scene.borg[a] = c->borg[a];
        P->scene.kneg[a] = -(kCellBias + c->borg[a]);
        P->scene.lo[a] = (float)c->vorg[a];
        P->scene.hi[a] = (float)(c->vorg[a] + c->vdim[a] - 1);
        P->probe_count[a] = c->field.probe_count[a];
        P->field_origin[a] = c->field.field_origin[a];
    }
    P->n_lights = c->n_lights;
    for (int i = 0; i < c->n_lights; i++) P->lights[i] = c->lights[i];
    light_bounds(*P);
    P->side_length = c->field.side_length;
    P->rx = c->rx;
    P->ry = c->ry;
    P->max_bounces = c->rs.max_bounces;
    P->screen_w = c->rs.screen_width;
    P->screen_h = c->rs.screen_height;
    memcpy(P->cam, c->cam, sizeof(P->cam));
    // camera.glsl:37  w = 1.0/tan(0.5*hfov): a per-frame uniform, evaluated once here
    P->cam_w = 1.0f / (float)tan((double)(0.5f * c->cam[17]));
    P->render_mode = c->rs.render_mode;
    P->visualize_probes = c->rs.visualize_probes != 0;
    P->weight_mode = c->weight_mode;
    P->distance_scale = c->distance_scale;
    P->layout = c->layout;
    P->oct = c->oct;
    P->tile_w = tile_w(c);
    P->tile_h = tile_h(c);
    P->early_out = c->variant == 2;
}

// The flat-colour table of the reference's block types: 2-5 are getColorAt's own flat
// colours (intersection.glsl:908-919); the procedural types take the base colour of
// their branch (README.md:266 "flat colors" variant).
static void default_palette(float* pal)
{
    static const float t[14][3] = {{0.f, 0.f, 0.f},       {0.99f, 0.3f, 0.3f},   {.95f, 0.f, 0.f},
                                   {0.f, .95f, 0.f},      {0.f, 0.f, .95f},      {0.95f, 0.95f, 0.95f},
                                   {1.f, 0.2f, 0.f},      {1.f, 0.f, 0.011f},    {1.f, 0.5f, 0.f},
                                   {1.f, 0.5f, 0.f},      {1.f, 0.5f, 0.f},      {1.f, 0.f, 0.f},
                                   {0.619f, 1.f, 0.278f}, {0.356f, 1.f, 0.101f}};
    memset(pal, 0, 256 * 3 * sizeof(float));
    memcpy(pal, t, sizeof(t));
}

static int alloc_voxels(ddgi_ctx* ctx, const int32_t dims[3], const int32_t origin[3], const float* palette)
{
    NEED(dims && origin, "null dims/origin");
    NEED(dims[0] > 0 && dims[1] > 0 && dims[2] > 0, "voxel dims must be positive");
    NEED((size_t)dims[0] * dims[1] * dims[2] <= ((size_t)1 << 34), "voxel field too large");
    for (int a = 0; a < 3; a++)
        NEED(origin[a] > -(1 << 22) && origin[a] + dims[a] < (1 << 22), "voxel ids must stay below 2^22");
    dfree(ctx->d_types);
    dfree(ctx->d_occ);
    for (int a = 0; a < 3; a++) {
        ctx->vdim[a] = dims[a];
        ctx->vorg[a] = origin[a];
        ctx->borg[a] = origin[a] & ~(kBrickAlign - 1);  // two's complement: rounds toward -inf to a multiple of 8
        int cells = 1 << (a == 0 ? kBrickLx : a == 1 ? kBrickLy : kBrickLz);
        ctx->nb[a] = (origin[a] + dims[a] - ctx->borg[a] + cells - 1) / cells;
    }
    size_t n = (size_t)dims[0] * dims[1] * dims[2];
    size_t nbk = (size_t)ctx->nb[0] * ctx->nb[1] * ctx->nb[2];
    CU(cudaMalloc(&ctx->d_types, n));
    CU(cudaMalloc(&ctx->d_occ, nbk * sizeof(uint32_t)));
    if (!ctx->d_palette) CU(cudaMalloc(&ctx->d_palette, 256 * 3 * sizeof(float)));
    float pal[256 * 3];
    if (palette) memcpy(pal, palette, sizeof(pal));
    else default_palette(pal);
    CU(cudaMemcpy(ctx->d_palette, pal, sizeof(pal), cudaMemcpyHostToDevice));
    return DDGI_OK;
}

static int finish_voxels(ddgi_ctx* ctx)
{
    int l = 0;
    int shift[3] = {ctx->vorg[0] - ctx->borg[0], ctx->vorg[1] - ctx->borg[1], ctx->vorg[2] - ctx->borg[2]};
    int b0[3] = {0, 0, 0};
    CU(launch_build_occupancy(ctx->vdim, shift, ctx->nb, b0, ctx->nb, ctx->d_types, ctx->d_occ, 0, &l));
    ctx->launches += l;
    CU(cudaDeviceSynchronize());
    ctx->calibrated = false;  // a new scene: measure the per-slot costs again
    return DDGI_OK;
}

// generate_samples, src/rvpt/rvpt.cpp:1147-1173, generalised to an rx x ry tile: libc
// rand() jitter, PI = 3.1415926 as a double, libm cosf/sinf/sqrtf.
static void host_generate_samples(int rx, int ry, bool y_first, std::vector<float>& out)
{
    const double host_pi = 3.1415926;
    float inv_x = 1.f / float(rx), inv_y = 1.f / float(ry);
    out.resize((size_t)rx * ry * 3);
    size_t i = 0;
    for (int y = 0; y < ry; y++)
        for (int x = 0; x < rx; x++) {
            // the two rand() calls are constructor arguments in the reference (rvpt.cpp:1161-1162): C++
            // leaves their order open.  x first is SURVEY.md 8c-5's pin (default); g++, the reference's
            // Linux toolchain, evaluates them right to left: y first (ddgi_set_sample_order)
            float jx, jy;
            if (y_first) {
                jy = float(rand()) / float(RAND_MAX);
                jx = float(rand()) / float(RAND_MAX);
            } else {
                jx = float(rand()) / float(RAND_MAX);
                jy = float(rand()) / float(RAND_MAX);
            }
            float su = (x + jx) * inv_x;
            float sv = (y + jy) * inv_y;
            float z = 1 - (2 * su);
            float ang = (float)(2.0f * host_pi * sv);
            float ring = sqrtf(1 - (z * z));
            out[i++] = cosf(ang) * ring;
            out[i++] = sinf(ang) * ring;
            out[i++] = z;
        }
}

static int upload_dirs(ddgi_ctx* ctx)
{
    size_t n = (size_t)ctx->rx * ctx->ry;
    std::vector<float> dirs(n * 3);
    for (size_t i = 0; i < n; i++) {
        v3 d = normalize(V3(ctx->samples[3 * i], ctx->samples[3 * i + 1], ctx->samples[3 * i + 2]));
        dirs[3 * i] = d.x;
        dirs[3 * i + 1] = d.y;
        dirs[3 * i + 2] = d.z;
    }
    if (n > ctx->dirs_cap) {
        // (growing: every update that may read an old table must be over)
        for (int b = 0; b < 3; b++) {
            if (ctx->ev_dirs[b]) CU(cudaEventSynchronize(ctx->ev_dirs[b]));
            dfree(ctx->d_dirs_ring[b]);
            CU(cudaMalloc(&ctx->d_dirs_ring[b], n * 3 * sizeof(float)));
        }
        ctx->dirs_cap = n;
    }
    // the next table of the ring: the update that read it last was launched at least two uploads ago
    // (waited for, normally long over); the updates in flight keep reading theirs
    if (!ctx->up_stream) CU(cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
    int b = (ctx->dirs_cur + 1) % 3;
    if (ctx->ev_dirs[b]) CU(cudaEventSynchronize(ctx->ev_dirs[b]));
    CU(cudaMemcpyAsync(ctx->d_dirs_ring[b], dirs.data(), n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->up_stream));
    CU(cudaStreamSynchronize(ctx->up_stream));  // (`dirs` is a local; 3 KB)
    ctx->dirs_cur = b;
    ctx->d_dirs = ctx->d_dirs_ring[b];
    ctx->ray_mode = 1;
    return DDGI_OK;
}

// The stream the work of the frame in texture allocation `b` runs on: the caller's, or with two frames
// in flight the engine's own frame_stream[b].
static cudaStream_t work_stream(ddgi_ctx* ctx, int b, cudaStream_t caller) { return ctx->in_flight == 2 ? ctx->frame_stream[b] : caller; }

// Makes the work stream of frame b follow everything the caller's stream has been given so far (the
// caller's uploads, edits and - what matters under an exchange - its readers of older frames).
static int follow_caller(ddgi_ctx* ctx, int b, cudaStream_t caller)
{
    if (ctx->in_flight != 2) return DDGI_OK;
    CU(cudaEventRecord(ctx->ev_in, caller));
    CU(cudaStreamWaitEvent(ctx->frame_stream[b], ctx->ev_in, 0));
    return DDGI_OK;
}

// Waits (host) for the engine's own streams.
static int drain_frames(ddgi_ctx* ctx)
{
    for (int b = 0; b < 2; b++)
        if (ctx->frame_stream[b]) CU(cudaStreamSynchronize(ctx->frame_stream[b]));
    return DDGI_OK;
}

// ------------------------------------------------------------------ C ABI
extern "C" {

const char* ddgi_version(void) { return "0.1 sm_100a"; }

int ddgi_create(ddgi_ctx** out, int device)
{
    if (!out) return DDGI_E_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return DDGI_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return DDGI_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DDGI_E_CUDA;
    if (prop.major != 10) return DDGI_E_CUDA;  // sm_100a code only
    ddgi_ctx* ctx = new ddgi_ctx();
    ctx->device = device;
    if (cudaMalloc(&ctx->d_counter, 2 * sizeof(uint32_t)) != cudaSuccess) {
        delete ctx;
        return DDGI_E_CUDA;
    }
    *out = ctx;
    return DDGI_OK;
}

void ddgi_destroy(ddgi_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    ddgi_close_peers(ctx);
    ddgi_comm_destroy(ctx);
    dfree(ctx->d_counter);
    dfree(ctx->d_types);
    dfree(ctx->d_occ);
    dfree(ctx->d_edit);
    dfree(ctx->d_palette);
    for (int b = 0; b < 3; b++) {
        dfree(ctx->d_dirs_ring[b]);
        if (ctx->ev_dirs[b]) cudaEventDestroy(ctx->ev_dirs[b]);
    }
    if (ctx->up_stream) cudaStreamDestroy(ctx->up_stream);
    for (int b = 0; b < 2; b++) {
        if (ctx->frame_stream[b]) cudaStreamDestroy(ctx->frame_stream[b]);
        if (ctx->ev_frame[b]) cudaEventDestroy(ctx->ev_frame[b]);
        if (ctx->ev_k0[b]) cudaEventDestroy(ctx->ev_k0[b]);
        if (ctx->ev_k1[b]) cudaEventDestroy(ctx->ev_k1[b]);
    }
    if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
    dfree(ctx->d_rays);
    dfree(ctx->d_tex_pair[0]);
    dfree(ctx->d_tex_pair[1]);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int b = 0; b < 2; b++)
        if (ctx->ev_copied[b]) cudaEventDestroy(ctx->ev_copied[b]);
    if (ctx->ev_ready) cudaEventDestroy(ctx->ev_ready);
    dfree(ctx->d_tex_f32);
    dfree(ctx->d_ray_lookups);
    dfree(ctx->d_frame);
    dfree(ctx->d_frame_f32);
    dfree(ctx->d_px_lookups);
    dfree(ctx->d_order);
    dfree(ctx->d_slot_cost);
    dfree(ctx->d_warp_times);
    dfree(ctx->d_barrier_error);
    dfree(ctx->d_ray_out);
    dfree(ctx->d_owned);
    if (ctx->ev_update) cudaEventDestroy(ctx->ev_update);
    delete ctx;
}

const char* ddgi_last_error(const ddgi_ctx* ctx) { return ctx ? ctx->err : "null context"; }
uint64_t ddgi_launch_count(const ddgi_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ddgi_set_render_settings(ddgi_ctx* ctx, const ddgi_render_settings* rs)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(rs, "null settings");
    NEED(rs->screen_width >= 0 && rs->screen_height >= 0 && rs->screen_width <= 16384 && rs->screen_height <= 16384,
         "bad screen size");
    NEED(rs->max_bounces >= 0 && rs->max_bounces <= 64, "max_bounces out of range");
    NEED(rs->camera_mode == 0, "only the pinhole camera (camera_mode 0) is supported");
    // render_mode: any value is legal, as in eval_integrator's switch (compute_pass.comp:58-87):
    // 1-5 are the debug views, everything else the DDGI integrator
    CU(cudaSetDevice(ctx->device));
    ctx->rs = *rs;
    return resize_frame(ctx);
}

int ddgi_set_irradiance_field(ddgi_ctx* ctx, const ddgi_irradiance_field* f)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(f, "null field");
    for (int a = 0; a < 3; a++) NEED(f->probe_count[a] >= 1 && f->probe_count[a] <= 1024, "probe_count out of range");
    NEED(f->side_length >= 1, "side_length must be >= 1");
    NEED(f->sqrt_rays_per_probe >= 1 && f->sqrt_rays_per_probe <= 64, "sqrt_rays_per_probe out of range");
    size_t probes = (size_t)f->probe_count[0] * f->probe_count[1] * f->probe_count[2];
    NEED(probes < ((size_t)1 << 24), "probe index must be exact in fp32 (< 2^24), src/rvpt/probe.h:13");
    NEED(probes * f->sqrt_rays_per_probe * f->sqrt_rays_per_probe < ((size_t)1 << 32), "too many rays");
    CU(cudaSetDevice(ctx->device));
    bool shape_changed = !ctx->have_field || memcmp(ctx->field.probe_count, f->probe_count, sizeof(f->probe_count)) ||
                         ctx->field.sqrt_rays_per_probe != f->sqrt_rays_per_probe;
    if (!shape_changed) {
        ctx->field = *f;
        return DDGI_OK;
    }
    // a new shape re-creates the textures, which can fail (peers open, out of memory): commit the
    // new field only when it has not, so that the context never holds a shape its textures do not have
    const ddgi_irradiance_field old_field = ctx->field;
    const bool old_have = ctx->have_field;
    const int old_rx = ctx->rx, old_ry = ctx->ry;
    ctx->field = *f;
    ctx->have_field = true;
    ctx->rx = ctx->ry = f->sqrt_rays_per_probe;
    int rc = resize_textures(ctx);
    if (rc != DDGI_OK) {
        ctx->field = old_field;
        ctx->have_field = old_have;
        ctx->rx = old_rx;
        ctx->ry = old_ry;
        return rc;
    }
    ctx->ray_mode = 0;
    ctx->row0 = 0;
    ctx->row1 = f->probe_count[1];
    ctx->order_dirty = true;
    ctx->calibrated = false;
    return DDGI_OK;
}

int ddgi_set_ray_tile(ddgi_ctx* ctx, int32_t rx, int32_t ry)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->have_field, "set the irradiance field first");
    NEED(rx >= 1 && ry >= 1 && rx <= 64 && ry <= 64, "tile out of range");
    CU(cudaSetDevice(ctx->device));
    if (rx != ctx->rx || ry != ctx->ry) {
        const int old_rx = ctx->rx, old_ry = ctx->ry;
        ctx->rx = rx;
        ctx->ry = ry;
        int rc = resize_textures(ctx);
        if (rc != DDGI_OK) {  // (see ddgi_set_irradiance_field)
            ctx->rx = old_rx;
            ctx->ry = old_ry;
            return rc;
        }
        ctx->ray_mode = 0;
        ctx->calibrated = false;
        ctx->order_dirty = true;  // the slot count changed
        free_ray_buffers(ctx);    // sized by rays per probe
    }
    return DDGI_OK;
}

int ddgi_set_camera(ddgi_ctx* ctx, const float cam[20])
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(cam, "null camera");
    memcpy(ctx->cam, cam, sizeof(ctx->cam));
    ctx->have_cam = true;
    return DDGI_OK;
}

int ddgi_set_lights(ddgi_ctx* ctx, int32_t n, const ddgi_light* lights)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(n >= 0 && n <= DDGI_MAX_LIGHTS, "at most %d lights", DDGI_MAX_LIGHTS);
    NEED(n == 0 || lights, "null lights");
    ctx->n_lights = n;
    for (int i = 0; i < n; i++) {
        ctx->lights[i].intensity = lights[i].intensity;
        for (int a = 0; a < 3; a++) {
            ctx->lights[i].col[a] = lights[i].col[a];
            ctx->lights[i].pos[a] = lights[i].pos[a];
        }
    }
    return DDGI_OK;
}

int ddgi_default_lights(int32_t scene, ddgi_light* out, int32_t* n)
{
    if (!out || !n) return DDGI_E_INVALID;
    static const ddgi_light cave = {100.f, {1.f, 1.f, 1.f}, {4.f, 17.5f, 8.5f}};
    static const ddgi_light cornell = {15.f, {1.f, 1.f, 1.f}, {0.f, 8.f, 13.f}};
    static const ddgi_light house[2] = {{1.f, {1.f, 1.f, 1.f}, {5.f, 9.3f, 36.5f}}, {1.f, {1.f, 1.f, 1.f}, {0.f, 0.f, 0.f}}};
    switch (scene) {
        case 0: out[0] = cave; *n = 1; return DDGI_OK;
        case 1: out[0] = cornell; *n = 1; return DDGI_OK;
        case 2: out[0] = house[0]; out[1] = house[1]; *n = 2; return DDGI_OK;
    }
    *n = 0;
    return DDGI_E_INVALID;
}

// update_lights, probe_pass.comp:217-250 (== compute_pass.comp:126-160)
int ddgi_update_lights(int32_t scene, float time, const ddgi_light* base, int32_t n, ddgi_light* out)
{
    if (!base || !out || n < 0 || n > DDGI_MAX_LIGHTS || scene < 0 || scene > 2) return DDGI_E_INVALID;
    for (int i = 0; i < n; i++) {
        ddgi_light l = base[i];
        if (scene == 0) {
            float t = 0.05f * time;
            if (i == 0) {
                l.pos[2] = base[i].pos[2] + 10.0f * pin_cos(t * 0.1f);
            } else {
                float sn = pin_sin(t * 0.5f), cs = pin_cos(t * 0.5f);
                l.pos[0] = base[i].pos[0] + (float)((i + 1) * 2) * sn;
                l.pos[1] = base[i].pos[1] + (float)((i / 2) * 4) * sn;
                l.pos[2] = base[i].pos[2] + (float)((i + 1) * 2) * cs;
            }
        } else if (scene == 1) {
            float t = 0.005f * time;
            float sn = pin_sin(t), cs = pin_cos(t);
            l.pos[0] = base[i].pos[0] + (float)(i + 1) * sn;
            l.pos[1] = base[i].pos[1] + (float)((i / 2) * 4) * sn;
            l.pos[2] = base[i].pos[2] + (float)(i + 1) * cs;
        } else {
            float d = 0.00005f * time;
            for (int a = 0; a < 3; a++) l.pos[a] = base[i].pos[a] + d;
        }
        out[i] = l;
    }
    return DDGI_OK;
}

int ddgi_cave_lights4(float time, ddgi_light* out)
{
    if (!out) return DDGI_E_INVALID;
    static const ddgi_light base[4] = {{20.f, {1.f, 1.f, 1.f}, {4.f, 17.5f, 8.5f}},
                                       {10.f, {1.f, 0.5f, 0.1f}, {0.f, 2.f, 0.f}},
                                       {10.f, {0.1f, 1.1f, 1.f}, {5.f, 0.f, 0.f}},
                                       {10.f, {1.1f, 0.f, 1.1f}, {0.f, 5.f, 0.f}}};
    float t = 0.05f * time;
    for (int i = 0; i < 4; i++) {
        out[i] = base[i];
        if (i == 0) {
            out[i].pos[2] = base[i].pos[2] + 10.0f * pin_cos(t * 0.1f);
            continue;
        }
        float s = pin_sin(t * 0.5f), c = pin_cos(t * 0.5f);
        out[i].pos[0] = base[i].pos[0] + (float)((i + 1) * 2) * s;
        out[i].pos[1] = base[i].pos[1] + (float)((i / 2) * 4) * s;
        out[i].pos[2] = base[i].pos[2] + (float)((i + 1) * 2) * c;
    }
    return DDGI_OK;
}

int ddgi_upload_voxels(ddgi_ctx* ctx, const int32_t dims[3], const int32_t origin[3], const uint8_t* types,
                       const float* palette)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(types, "null voxel types");
    CU(cudaSetDevice(ctx->device));
    int rc = alloc_voxels(ctx, dims, origin, palette);
    if (rc) return rc;
    CU(cudaMemcpy(ctx->d_types, types, (size_t)dims[0] * dims[1] * dims[2], cudaMemcpyHostToDevice));
    return finish_voxels(ctx);
}

int ddgi_bake_scene(ddgi_ctx* ctx, int32_t scene, const int32_t dims[3], const int32_t origin[3])
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(scene >= 0 && scene <= 2, "scene must be 0, 1 or 2");
    CU(cudaSetDevice(ctx->device));
    int rc = alloc_voxels(ctx, dims, origin, nullptr);
    if (rc) return rc;
    int l = 0;
    CU(launch_bake_scene(scene, ctx->vdim, ctx->vorg, ctx->d_types, 0, &l));
    ctx->launches += l;
    return finish_voxels(ctx);
}

int ddgi_bake_synthetic(ddgi_ctx* ctx, const int32_t dims[3], const int32_t origin[3], int32_t solid_permille,
                        uint32_t seed)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(solid_permille >= 0 && solid_permille <= 1000, "solid_permille in [0,1000]");
    CU(cudaSetDevice(ctx->device));
    int rc = alloc_voxels(ctx, dims, origin, nullptr);
    if (rc) return rc;
    int l = 0;
    CU(launch_bake_synthetic(ctx->vdim, ctx->vorg, solid_permille, seed, ctx->d_types, 0, &l));
    ctx->launches += l;
    return finish_voxels(ctx);
}

int ddgi_edit_voxels(ddgi_ctx* ctx, const int32_t origin[3], const int32_t dims[3], const uint8_t* types, void* stream)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->d_types && ctx->d_occ, "no voxel field");
    NEED(origin && dims && types, "null origin / dims / types");
    int at[3], ext[3];
    for (int a = 0; a < 3; a++) {
        at[a] = origin[a] - ctx->vorg[a];
        ext[a] = dims[a];
        NEED(ext[a] > 0 && at[a] >= 0 && at[a] + ext[a] <= ctx->vdim[a], "edit box must lie inside the voxel field");
    }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    {   // updates in flight on the engine's own streams still read the field: the edit follows them
        int rc = ddgi_frame_fence(ctx, stream);
        if (rc) return rc;
    }
    size_t n = (size_t)ext[0] * ext[1] * ext[2];
    if (n > ctx->edit_cap) {
        CU(cudaStreamSynchronize(s));  // an earlier edit may still read the old staging buffer
        dfree(ctx->d_edit);
        CU(cudaMalloc(&ctx->d_edit, n));
        ctx->edit_cap = n;
    }
    CU(cudaMemcpyAsync(ctx->d_edit, types, n, cudaMemcpyHostToDevice, s));
    int l = 0;
    CU(launch_edit_voxels(ctx->vdim, at, ext, ctx->d_edit, ctx->d_types, s, &l));
    // the bricks the box touches (aligned to borg)
    int shift[3], b0[3], bn[3];
    for (int a = 0; a < 3; a++) {
        int cells = 1 << (a == 0 ? kBrickLx : a == 1 ? kBrickLy : kBrickLz);
        shift[a] = ctx->vorg[a] - ctx->borg[a];
        b0[a] = (at[a] + shift[a]) / cells;
        bn[a] = (at[a] + ext[a] - 1 + shift[a]) / cells - b0[a] + 1;
    }
    CU(launch_build_occupancy(ctx->vdim, shift, ctx->nb, b0, bn, ctx->d_types, ctx->d_occ, s, &l));
    ctx->launches += l;
    ctx->last_stream = s;
    // `types` may be reused on return: a copy from pageable memory is staged by the runtime before
    // cudaMemcpyAsync returns; a PINNED source must stay untouched until the stream has passed the
    // copy (ddgi_sync, or any later synchronisation of `stream`) - documented in ddgi.h
    return DDGI_OK;
}

int ddgi_read_voxels(ddgi_ctx* ctx, uint8_t* dst, size_t bytes)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->d_types, "no voxel field");
    NEED(dst && bytes == (size_t)ctx->vdim[0] * ctx->vdim[1] * ctx->vdim[2], "size mismatch");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->last_stream));  // the dispatches may run on a non-blocking stream
    CU(cudaMemcpy(dst, ctx->d_types, bytes, cudaMemcpyDeviceToHost));
    return DDGI_OK;
}

int ddgi_generate_probe_rays(ddgi_ctx* ctx, int32_t reseed)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->have_field, "set the irradiance field first");
    CU(cudaSetDevice(ctx->device));
    if (reseed) srand(1);
    host_generate_samples(ctx->rx, ctx->ry, ctx->sample_y_first, ctx->samples);
    return upload_dirs(ctx);
}

// Spherical Fibonacci set (the north star's ray generator; the reference's is the stratified
// libc-rand() set above): direction i of n has cos(theta) = 1 - (2i + 1)/n and azimuth
// 2 pi frac(i (phi - 1)), phi the golden ratio, evaluated in fp64 and rounded once to fp32.
int ddgi_generate_fibonacci_rays(ddgi_ctx* ctx)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->have_field, "set the irradiance field first");
    CU(cudaSetDevice(ctx->device));
    const int n = ctx->rx * ctx->ry;
    const double golden = 0.61803398874989484820;  // phi - 1
    ctx->samples.resize((size_t)n * 3);
    for (int i = 0; i < n; i++) {
        double f = (double)i * golden;
        double az = 6.283185307179586476925 * (f - floor(f));
        double z = 1.0 - (2.0 * (double)i + 1.0) / (double)n;
        double ring = sqrt(1.0 - z * z);
        ctx->samples[3 * (size_t)i] = (float)(cos(az) * ring);
        ctx->samples[3 * (size_t)i + 1] = (float)(sin(az) * ring);
        ctx->samples[3 * (size_t)i + 2] = (float)z;
    }
    return upload_dirs(ctx);
}

int ddgi_set_layout(ddgi_ctx* ctx, int32_t layout, int32_t oct)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->have_field, "set the irradiance field first");
    NEED(layout == DDGI_LAYOUT_RAY_TILE || layout == DDGI_LAYOUT_OCTAHEDRAL, "layout must be DDGI_LAYOUT_RAY_TILE or DDGI_LAYOUT_OCTAHEDRAL");
    if (layout == DDGI_LAYOUT_OCTAHEDRAL) {
        NEED(oct >= 2 && oct <= 64, "octahedral tile size in [2,64]");
        NEED(ctx->rx * ctx->ry <= 4096, "the octahedral layout stages a probe's rays in shared memory: at most 4096 rays/probe");
    }
    CU(cudaSetDevice(ctx->device));
    ctx->layout = layout;
    if (layout == DDGI_LAYOUT_OCTAHEDRAL) ctx->oct = oct;
    return resize_textures(ctx);
}

int ddgi_set_ray_samples(ddgi_ctx* ctx, const float* samples, size_t count)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->have_field, "set the irradiance field first");
    NEED(samples && count == (size_t)ctx->rx * ctx->ry, "expected rx*ry samples");
    CU(cudaSetDevice(ctx->device));
    ctx->samples.assign(samples, samples + 3 * count);
    return upload_dirs(ctx);
}

int ddgi_get_ray_samples(ddgi_ctx* ctx, float* dst, size_t count)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->ray_mode == 1, "no generated ray set");
    NEED(dst && count * 3 == ctx->samples.size(), "expected rx*ry samples");
    memcpy(dst, ctx->samples.data(), ctx->samples.size() * sizeof(float));
    return DDGI_OK;
}

int ddgi_set_probe_rays(ddgi_ctx* ctx, const ddgi_probe_ray* rays, size_t count)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->have_field, "set the irradiance field first");
    NEED(rays && count == num_rays(ctx), "expected probes*rx*ry rays");
    CU(cudaSetDevice(ctx->device));
    if (ctx->ev_update) CU(cudaEventSynchronize(ctx->ev_update));  // the last update may still read the old list
    if (ctx->n_rays_ssbo != count) {
        dfree(ctx->d_rays);
        CU(cudaMalloc(&ctx->d_rays, count * sizeof(ddgi_probe_ray)));
        ctx->n_rays_ssbo = count;
    }
    CU(cudaMemcpy(ctx->d_rays, rays, count * sizeof(ddgi_probe_ray), cudaMemcpyHostToDevice));
    ctx->ray_mode = 2;
    return DDGI_OK;
}

size_t ddgi_num_probe_rays(const ddgi_ctx* ctx) { return ctx ? num_rays(ctx) : 0; }

// RVPT::generate_probe_rays, src/rvpt/rvpt.cpp:1177-1224
int ddgi_get_probe_rays(ddgi_ctx* ctx, ddgi_probe_ray* dst, size_t count)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->ray_mode != 0, "no ray set");
    NEED(dst && count == num_rays(ctx), "expected probes*rx*ry rays");
    if (ctx->ray_mode == 2) {
        CU(cudaSetDevice(ctx->device));
        CU(cudaMemcpy(dst, ctx->d_rays, count * sizeof(ddgi_probe_ray), cudaMemcpyDeviceToHost));
        return DDGI_OK;
    }
    FrameParams P;
    fill_params(ctx, &P);
    int n = ctx->rx * ctx->ry;
    int probes = (int)(count / n);
    size_t o = 0;
    for (int p = 0; p < probes; p++) {
        v3 org = probe_origin(P, p);
        for (int i = 0; i < n; i++) {
            v3 d = normalize(V3(ctx->samples[3 * i], ctx->samples[3 * i + 1], ctx->samples[3 * i + 2]));
            ddgi_probe_ray* r = &dst[o++];
            memset(r, 0, sizeof(*r));
            r->origin[0] = org.x; r->origin[1] = org.y; r->origin[2] = org.z;
            r->direction[0] = d.x; r->direction[1] = d.y; r->direction[2] = d.z;
            r->probe_info[0] = (float)p;
            r->probe_info[1] = (float)(i % ctx->rx);
            r->probe_info[2] = (float)(i / ctx->rx);
        }
    }
    return DDGI_OK;
}

int ddgi_set_probe_rows(ddgi_ctx* ctx, int32_t y0, int32_t y1)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->have_field, "set the irradiance field first");
    NEED(0 <= y0 && y0 <= y1 && y1 <= ctx->field.probe_count[1], "rows out of range");
    ctx->row0 = y0;
    ctx->row1 = y1;
    ctx->cyc_world = 0;
    ctx->order_dirty = true;
    ctx->calibrated = false;
    return DDGI_OK;
}

int ddgi_set_probe_rows_cyclic(ddgi_ctx* ctx, int32_t rank, int32_t world, int32_t block)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->have_field, "set the irradiance field first");
    NEED(world >= 1 && rank >= 0 && rank < world && block >= 1, "bad rank / world / block");
    ctx->cyc_world = world;
    ctx->cyc_rank = rank;
    ctx->cyc_block = block;
    ctx->cyc_unit = 0;
    ctx->order_dirty = true;
    ctx->calibrated = false;
    return DDGI_OK;
}

int ddgi_set_probes_cyclic(ddgi_ctx* ctx, int32_t rank, int32_t world, int32_t block)
{
    int rc = ddgi_set_probe_rows_cyclic(ctx, rank, world, block);
    if (rc == DDGI_OK) ctx->cyc_unit = 1;  // order_dirty already set
    return rc;
}

int ddgi_probe_texture_device_ptr(ddgi_ctx* ctx, int32_t which, void** ptr, size_t* bytes)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->d_tex, "no probe texture");
    NEED(ptr && bytes && (which == 0 || which == 1), "bad arguments");
    *ptr = ctx->d_tex + (which ? tex_texels(ctx) : 0);
    *bytes = tex_texels(ctx) * sizeof(uint32_t);
    return DDGI_OK;
}

int ddgi_export_texture_handle(ddgi_ctx* ctx, void* handle64)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->d_tex && handle64, "no probe texture");
    NEED(!ctx->double_buffer, "a double-buffered context has two allocations: use ddgi_export_texture_handles");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->d_tex));
    memcpy(handle64, &h, 64);
    return DDGI_OK;
}

int ddgi_export_texture_handles(ddgi_ctx* ctx, void* handles, int32_t* count)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->d_tex && handles && count, "no probe texture");
    CU(cudaSetDevice(ctx->device));
    *count = ctx->double_buffer ? 2 : 1;
    for (int b = 0; b < *count; b++) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, ctx->double_buffer ? ctx->d_tex_pair[b] : ctx->d_tex));
        memcpy((char*)handles + 64 * b, &h, 64);
    }
    return DDGI_OK;
}

int ddgi_open_peers(ddgi_ctx* ctx, int32_t n_peers, const void* handles64, int32_t self_index)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->d_tex, "no probe texture");
    NEED(n_peers >= 1 && n_peers <= kMaxPeers && handles64 && self_index >= 0 && self_index < n_peers, "bad peers");
    CU(cudaSetDevice(ctx->device));
    ddgi_close_peers(ctx);
    // per rank as many handles as this context has allocations (every rank must be configured alike)
    const int nb = ctx->double_buffer ? 2 : 1;
    for (int g = 0; g < n_peers; g++)
        for (int b = 0; b < nb; b++) {
            if (g == self_index) {
                ctx->peer_base[b][g] = ctx->double_buffer ? ctx->d_tex_pair[b] : ctx->d_tex;
                continue;
            }
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)handles64 + 64 * ((size_t)g * nb + b), 64);
            cudaError_t e = cudaIpcOpenMemHandle(&ctx->peer_base[b][g], h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                ctx->peer_base[b][g] = nullptr;
                ddgi_close_peers(ctx);
                return fail(ctx, DDGI_E_CUDA, "cudaIpcOpenMemHandle (rank %d, buffer %d): %s", g, b, cudaGetErrorString(e));
            }
            ctx->peer_opened[b][g] = true;
        }
    ctx->n_peers = n_peers;
    ctx->self_index = self_index;
    ctx->peer_buffers = nb;
    return DDGI_OK;
}

int ddgi_close_peers(ddgi_ctx* ctx)
{
    if (!ctx) return DDGI_E_INVALID;
    for (int b = 0; b < 2; b++)
        for (int g = 0; g < kMaxPeers; g++) {
            if (ctx->peer_opened[b][g]) cudaIpcCloseMemHandle(ctx->peer_base[b][g]);
            ctx->peer_opened[b][g] = false;
            ctx->peer_base[b][g] = nullptr;
        }
    ctx->n_peers = 0;
    ctx->peer_buffers = 0;
    return DDGI_OK;
}

// One epoch barrier over the peers' flag words (which live behind the planes of allocation [0] of
// every rank): `phase` 0 = completion barrier after an update (flag slots 0..7), 1 = the barrier in
// front of an update of a single-buffered context (slots 32..39).
static int issue_barrier(ddgi_ctx* ctx, int phase, cudaStream_t stream)
{
    if (!ctx->d_barrier_error) {
        CU(cudaMalloc(&ctx->d_barrier_error, sizeof(uint32_t)));
        CU(cudaMemset(ctx->d_barrier_error, 0, sizeof(uint32_t)));
    }
    PeerBarrier B;
    memset(&B, 0, sizeof(B));
    B.n_ranks = ctx->n_peers;
    B.self = ctx->self_index;
    B.epoch = phase == 0 ? ++ctx->epoch : ++ctx->epoch_pre;
    B.timeout_ns = 5000000000ull;
    size_t flags_at = 2 * tex_texels(ctx) + (phase ? kFlagWords / 2 : 0);
    B.local_flags = (uint32_t*)ctx->peer_base[0][ctx->self_index] + flags_at;
    for (int g = 0; g < ctx->n_peers; g++) B.peer_flags[g] = (uint32_t*)ctx->peer_base[0][g] + flags_at;
    B.error = ctx->d_barrier_error;
    int l = 0;
    CU(launch_peer_barrier(B, stream, &l));
    ctx->launches += l;
    return DDGI_OK;
}

int ddgi_exchange_barrier(ddgi_ctx* ctx, void* stream)
{
    if (!ctx) return DDGI_E_INVALID;
    if (ctx->n_peers < 1 || !ctx->d_tex) return fail(ctx, DDGI_E_STATE, "no peers: call ddgi_open_peers first");
    CU(cudaSetDevice(ctx->device));
    // The peers' NEXT update stores into the allocation this context's next update writes too.  An
    // asynchronous read of that allocation (ddgi_read_probe_texture_async, on the copy stream) must
    // be over before any peer may start: this rank only arrives at the barrier once it is.
    int next = ctx->double_buffer ? ctx->cur_tex ^ 1 : ctx->cur_tex;
    cudaStream_t w = work_stream(ctx, ctx->cur_tex, (cudaStream_t)stream);  // behind the update it completes
    if (ctx->ev_copied[next]) CU(cudaStreamWaitEvent(w, ctx->ev_copied[next], 0));
    ctx->last_stream = w;
    int rc = issue_barrier(ctx, 0, w);
    if (rc) return rc;
    if (ctx->ev_frame[ctx->cur_tex]) CU(cudaEventRecord(ctx->ev_frame[ctx->cur_tex], w));
    return DDGI_OK;
}

int ddgi_exchange_status(ddgi_ctx* ctx)
{
    if (!ctx) return DDGI_E_INVALID;
    if (!ctx->d_barrier_error) return DDGI_OK;
    CU(cudaSetDevice(ctx->device));
    uint32_t e = 0;
    CU(cudaMemcpy(&e, ctx->d_barrier_error, sizeof(e), cudaMemcpyDeviceToHost));
    if (e) return fail(ctx, DDGI_E_STATE, "a peer did not reach the exchange barrier within 5 s");
    return DDGI_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ NCCL exchange
// The collective library is bound at run time (dlopen of libnccl.so.2: the copy the process
// already holds, e.g. torch's, or the system's), so libddgi_b200.so has no link-time dependency
// on it and a single-GPU host needs no NCCL at all.
namespace {
struct NcclId {
    char internal[128];  // ncclUniqueId
};
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    const char* why = nullptr;
};
constexpr int kNcclUint8 = 1;  // ncclUint8

NcclApi& nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) {
        api.why = "libnccl.so.2 not found";
        return api;
    }
    bool ok = true;
    auto sym = [&](const char* n) {
        void* p = dlsym(api.lib, n);
        ok = ok && p != nullptr;
        return p;
    };
    api.GetUniqueId = (int (*)(NcclId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(void**, int, NcclId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))sym("ncclAllGather");
    api.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclBroadcast");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!ok) {
        api.why = "libnccl.so.2 lacks an expected symbol";
        api.lib = nullptr;
    }
    return api;
}
}  // namespace
#define NCCLCHK(call)                                                                                     \
    do {                                                                                                  \
        int r_ = (call);                                                                                  \
        if (r_ != 0) return fail(ctx, DDGI_E_CUDA, "%s: %s", #call, nccl().GetErrorString ? nccl().GetErrorString(r_) : "nccl error"); \
    } while (0)

extern "C" {

int ddgi_comm_unique_id(void* id128)
{
    if (!id128) return DDGI_E_INVALID;
    NcclApi& N = nccl();
    if (!N.lib) return DDGI_E_STATE;
    NcclId id;
    if (N.GetUniqueId(&id) != 0) return DDGI_E_CUDA;
    memcpy(id128, &id, sizeof(id));
    return DDGI_OK;
}

int ddgi_comm_init(ddgi_ctx* ctx, const void* id128, int32_t rank, int32_t world)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(id128 && world >= 1 && rank >= 0 && rank < world, "bad communicator arguments");
    NcclApi& N = nccl();
    if (!N.lib) return fail(ctx, DDGI_E_STATE, "NCCL is not available: %s", N.why ? N.why : "?");
    CU(cudaSetDevice(ctx->device));
    if (ctx->nccl_comm) {
        N.CommDestroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    NCCLCHK(N.CommInitRank(&ctx->nccl_comm, world, id, rank));
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    return DDGI_OK;
}

int ddgi_comm_destroy(ddgi_ctx* ctx)
{
    if (!ctx) return DDGI_E_INVALID;
    if (ctx->nccl_comm) {
        cudaSetDevice(ctx->device);
        nccl().CommDestroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->comm_world = 0;
    return DDGI_OK;
}

// In-place exchange of the texture planes over NCCL after ddgi_probe_update: probe row
// y is texture rows [y*th, (y+1)*th), a contiguous byte range of each plane.
//  * contiguous slabs of Y/world probe rows (ddgi_set_probe_rows(rank*Y/world, (rank+1)*Y/world)):
//    ONE ncclAllGather per plane whose send buffer is the rank's own slab inside the receive buffer;
//  * block-cyclic rows (ddgi_set_probe_rows_cyclic(rank, world, block)): every block is broadcast in
//    place from its owner, all of them inside one ncclGroupStart / ncclGroupEnd.
// The distance plane is skipped while it only holds the reference's zeros (every replica has them).
int ddgi_exchange_allgather(ddgi_ctx* ctx, void* stream)
{
    if (!ctx) return DDGI_E_INVALID;
    if (!ctx->nccl_comm) return fail(ctx, DDGI_E_STATE, "no communicator: call ddgi_comm_init first");
    if (!ctx->have_field || !ctx->d_tex) return fail(ctx, DDGI_E_STATE, "no irradiance field");
    NcclApi& N = nccl();
    CU(cudaSetDevice(ctx->device));
    const int Y = ctx->field.probe_count[1], G = ctx->comm_world, r = ctx->comm_rank;
    const size_t row_bytes = (size_t)ctx->tex_w * 4 * tile_h(ctx);  // one probe row of one plane
    const size_t plane = tex_texels(ctx) * 4;
    const int planes = ctx->distance_dirty[ctx->cur_tex] ? 2 : 1;
    cudaStream_t s = work_stream(ctx, ctx->cur_tex, (cudaStream_t)stream);  // behind the update it completes
    ctx->last_stream = s;
    if (ctx->cyc_world == 0) {
        if (Y % G != 0 || ctx->row0 != r * (Y / G) || ctx->row1 != (r + 1) * (Y / G))
            return fail(ctx, DDGI_E_STATE, "ddgi_exchange_allgather needs equal slabs: ddgi_set_probe_rows(rank*Y/world, (rank+1)*Y/world) "
                                          "with Y %% world == 0, or block-cyclic rows (ddgi_set_probe_rows_cyclic)");
        const size_t chunk = (size_t)(Y / G) * row_bytes;
        for (int p = 0; p < planes; p++) {
            char* base = (char*)ctx->d_tex + p * plane;
            NCCLCHK(N.AllGather(base + r * chunk, base, chunk, kNcclUint8, ctx->nccl_comm, s));
        }
        if (ctx->ev_frame[ctx->cur_tex]) CU(cudaEventRecord(ctx->ev_frame[ctx->cur_tex], s));
        return DDGI_OK;
    }
    if (ctx->cyc_unit != 0) {
        // Probe-cyclic ownership (the balanced one): a rank's tiles are scattered over the texture, so they are
        // packed into one chunk per rank, gathered with ONE ncclAllGather and scattered again (ddgi_kernels.cu:
        // pack_tiles / unpack_tiles; both planes travel when the distance plane is in use).
        if (ctx->cyc_world != G || ctx->cyc_rank != r)
            return fail(ctx, DDGI_E_STATE, "probe ownership (rank %d of %d) does not match the communicator (rank %d of %d)", ctx->cyc_rank,
                        ctx->cyc_world, r, G);
        int rc = schedule(ctx);  // (the owned-probe list)
        if (rc) return rc;
        const int B = ctx->cyc_block, n_probes = (int)num_probes(ctx);
        const int blocks = (n_probes + B - 1) / B;
        // rank 0 owns the most: ceil(blocks / G) blocks, the last of which may be short only if it is the field's last
        const int most_blocks = (blocks + G - 1) / G;
        int max_owned = most_blocks * B;
        if (max_owned > n_probes) max_owned = n_probes;
        TilePack T;
        memset(&T, 0, sizeof(T));
        T.tw = tile_w(ctx);
        T.th = tile_h(ctx);
        T.planes = planes;
        T.plane = tex_texels(ctx);
        T.max_owned = max_owned;
        T.chunk_texels = (size_t)max_owned * T.tw * T.th * planes;
        const size_t need = T.chunk_texels * G;
        if (need > ctx->pack_cap) {
            CU(cudaStreamSynchronize(s));
            dfree(ctx->d_pack);
            ctx->pack_cap = 0;
            CU(cudaMalloc(&ctx->d_pack, need * sizeof(uint32_t)));
            ctx->pack_cap = need;
        }
        T.tex = ctx->d_tex;
        T.pack = ctx->d_pack;
        T.owned = ctx->d_owned;
        T.n_owned = (int)ctx->owned.size();
        T.tiles_x = ctx->field.probe_count[0] * ctx->field.probe_count[2];
        T.tex_w = ctx->tex_w;
        T.G = G;
        T.self = r;
        T.B = B;
        T.n_probes = n_probes;
        int l = 0;
        CU(launch_pack_tiles(T, false, s, &l));
        NCCLCHK(N.AllGather((char*)ctx->d_pack + (size_t)r * T.chunk_texels * 4, ctx->d_pack, T.chunk_texels * 4, kNcclUint8, ctx->nccl_comm, s));
        CU(launch_pack_tiles(T, true, s, &l));
        ctx->launches += l;
        if (ctx->ev_frame[ctx->cur_tex]) CU(cudaEventRecord(ctx->ev_frame[ctx->cur_tex], s));
        return DDGI_OK;
    }
    if (ctx->cyc_world != G || ctx->cyc_rank != r)
        return fail(ctx, DDGI_E_STATE, "probe-row ownership (rank %d of %d) does not match the communicator (rank %d of %d)", ctx->cyc_rank,
                    ctx->cyc_world, r, G);
    NCCLCHK(N.GroupStart());
    for (int p = 0; p < planes; p++) {
        char* base = (char*)ctx->d_tex + p * plane;
        for (int y = 0, b = 0; y < Y; y += ctx->cyc_block, b++) {
            int y1 = y + ctx->cyc_block < Y ? y + ctx->cyc_block : Y;
            char* at = base + (size_t)y * row_bytes;
            int rc = N.Broadcast(at, at, (size_t)(y1 - y) * row_bytes, kNcclUint8, b % G, ctx->nccl_comm, s);
            if (rc != 0) {
                N.GroupEnd();
                return fail(ctx, DDGI_E_CUDA, "ncclBroadcast: %s", N.GetErrorString(rc));
            }
        }
    }
    NCCLCHK(N.GroupEnd());
    if (ctx->ev_frame[ctx->cur_tex]) CU(cudaEventRecord(ctx->ev_frame[ctx->cur_tex], s));
    return DDGI_OK;
}

// ------------------------------------------------------------------ on-disk formats
// Voxel file: "DDGIVOX1", int32 dims[3] (x, y, z), int32 origin[3], then dims product block types, x fastest.
// Checkpoint: "DDGIPTX1", int32 W, int32 H, float time, albedo plane, distance plane (RGBA8 rows).
// All little-endian (the only byte order the engine runs on).
int ddgi_save_voxels(ddgi_ctx* ctx, const char* path)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(path, "null path");
    NEED(ctx->d_types, "no voxel field");
    size_t n = (size_t)ctx->vdim[0] * ctx->vdim[1] * ctx->vdim[2];
    std::vector<uint8_t> vox(n);
    int rc = ddgi_read_voxels(ctx, vox.data(), n);
    if (rc) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(ctx, DDGI_E_INVALID, "%s: cannot open for writing", path);
    int32_t hdr[6] = {ctx->vdim[0], ctx->vdim[1], ctx->vdim[2], ctx->vorg[0], ctx->vorg[1], ctx->vorg[2]};
    bool ok = fwrite("DDGIVOX1", 1, 8, f) == 8 && fwrite(hdr, sizeof(hdr), 1, f) == 1 && fwrite(vox.data(), 1, n, f) == n;
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(ctx, DDGI_E_INVALID, "%s: write failed", path);
    return DDGI_OK;
}

int ddgi_load_voxels(ddgi_ctx* ctx, const char* path, const float* palette, int32_t dims_out[3], int32_t origin_out[3])
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(path, "null path");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(ctx, DDGI_E_INVALID, "%s: cannot open", path);
    char magic[8];
    int32_t hdr[6];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "DDGIVOX1", 8) != 0 || fread(hdr, sizeof(hdr), 1, f) != 1) {
        fclose(f);
        return fail(ctx, DDGI_E_INVALID, "%s: not a voxel file", path);
    }
    if (hdr[0] <= 0 || hdr[1] <= 0 || hdr[2] <= 0 || (size_t)hdr[0] * hdr[1] * hdr[2] > ((size_t)1 << 34)) {
        fclose(f);
        return fail(ctx, DDGI_E_INVALID, "%s: bad voxel dimensions", path);
    }
    size_t n = (size_t)hdr[0] * hdr[1] * hdr[2];
    std::vector<uint8_t> vox(n);
    size_t got = fread(vox.data(), 1, n, f);
    fclose(f);
    if (got != n) return fail(ctx, DDGI_E_INVALID, "%s: truncated voxel file", path);
    if (dims_out) memcpy(dims_out, hdr, 3 * sizeof(int32_t));
    if (origin_out) memcpy(origin_out, hdr + 3, 3 * sizeof(int32_t));
    return ddgi_upload_voxels(ctx, hdr, hdr + 3, vox.data(), palette);
}

int ddgi_save_checkpoint(ddgi_ctx* ctx, const char* path, float time)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(path, "null path");
    NEED(ctx->d_tex, "no probe texture");
    size_t n = tex_texels(ctx);
    std::vector<uint32_t> planes(2 * n);
    for (int which = 0; which < 2; which++) {
        int rc = ddgi_read_probe_texture(ctx, which, DDGI_FMT_RGBA8, planes.data() + which * n, n * 4);
        if (rc) return rc;
    }
    FILE* f = fopen(path, "wb");
    if (!f) return fail(ctx, DDGI_E_INVALID, "%s: cannot open for writing", path);
    int32_t wh[2] = {ctx->tex_w, ctx->tex_h};
    bool ok = fwrite("DDGIPTX1", 1, 8, f) == 8 && fwrite(wh, sizeof(wh), 1, f) == 1 && fwrite(&time, 4, 1, f) == 1 &&
              fwrite(planes.data(), 4, 2 * n, f) == 2 * n;
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(ctx, DDGI_E_INVALID, "%s: write failed", path);
    return DDGI_OK;
}

int ddgi_load_checkpoint(ddgi_ctx* ctx, const char* path, float* time_out)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(path, "null path");
    NEED(ctx->d_tex, "no probe texture: set the irradiance field of the dumped shape first");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(ctx, DDGI_E_INVALID, "%s: cannot open", path);
    char magic[8];
    int32_t wh[2];
    float time = 0.0f;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "DDGIPTX1", 8) != 0 || fread(wh, sizeof(wh), 1, f) != 1 || fread(&time, 4, 1, f) != 1) {
        fclose(f);
        return fail(ctx, DDGI_E_INVALID, "%s: not a probe-texture checkpoint", path);
    }
    if (wh[0] != ctx->tex_w || wh[1] != ctx->tex_h) {
        fclose(f);
        return fail(ctx, DDGI_E_INVALID, "%s: checkpoint is %dx%d, the field's texture is %dx%d", path, wh[0], wh[1], ctx->tex_w, ctx->tex_h);
    }
    size_t n = tex_texels(ctx);
    std::vector<uint32_t> planes(2 * n);
    size_t got = fread(planes.data(), 4, 2 * n, f);
    fclose(f);
    if (got != 2 * n) return fail(ctx, DDGI_E_INVALID, "%s: truncated checkpoint", path);
    for (int which = 0; which < 2; which++) {
        int rc = ddgi_write_probe_texture(ctx, which, planes.data() + which * n, n * 4);
        if (rc) return rc;
    }
    if (time_out) *time_out = time;
    return DDGI_OK;
}

int ddgi_probe_update(ddgi_ctx* ctx, void* stream)
{
    if (!ctx) return DDGI_E_INVALID;
    if (!ctx->have_field || !ctx->d_tex) return fail(ctx, DDGI_E_STATE, "no irradiance field");
    if (!ctx->d_occ) return fail(ctx, DDGI_E_STATE, "no voxel field");
    if (ctx->ray_mode == 0) return fail(ctx, DDGI_E_STATE, "no probe rays: call ddgi_generate_probe_rays or ddgi_set_probe_rays");
    if (ctx->layout == DDGI_LAYOUT_OCTAHEDRAL && ctx->ray_mode != 1)
        return fail(ctx, DDGI_E_STATE, "the octahedral layout needs a generated ray set (ddgi_generate_probe_rays / _fibonacci_rays / ddgi_set_ray_samples)");
    if (ctx->tex_w != ctx->field.probe_count[0] * ctx->field.probe_count[2] * tile_w(ctx) ||
        ctx->tex_h != ctx->field.probe_count[1] * tile_h(ctx))
        return fail(ctx, DDGI_E_STATE, "the probe textures do not have the field's shape");
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_debug_buffers(ctx);
    if (rc) return rc;
    FrameParams P;
    fill_params(ctx, &P);
    // the allocation this update writes (the other one under double buffering) and the stream it runs on
    const int target = ctx->double_buffer ? ctx->cur_tex ^ 1 : ctx->cur_tex;
    const cudaStream_t caller = (cudaStream_t)stream;
    stream = (void*)work_stream(ctx, target, caller);
    rc = follow_caller(ctx, target, caller);
    if (rc) return rc;
    if (ctx->in_flight == 2 && (ctx->blend_mode || ctx->debug || ctx->layout == DDGI_LAYOUT_OCTAHEDRAL) && ctx->ev_frame[target ^ 1])
        // the blend reads the previous frame, the debug and ray-result buffers are shared: no overlap then
        CU(cudaStreamWaitEvent((cudaStream_t)stream, ctx->ev_frame[target ^ 1], 0));
    ProbeJob J;
    memset(&J, 0, sizeof(J));
    J.rays = ctx->ray_mode == 2 ? ctx->d_rays : nullptr;
    J.dirs = ctx->d_dirs;
    rc = schedule(ctx);
    if (rc) return rc;
    bool calibrate = ctx->auto_schedule && !ctx->calibrated;
    J.order = ctx->d_order;
    J.n_owned = (uint32_t)ctx->order.size();
    J.slot_rays = slot_rays(ctx);
    if (calibrate) {
        CU(cudaMemsetAsync(ctx->d_slot_cost, 0, num_slots(ctx) * sizeof(uint32_t), (cudaStream_t)stream));
        J.slot_cost = ctx->d_slot_cost;
    }
    J.blend = ctx->blend_mode;
    J.hysteresis = ctx->field.hysteresis;
    J.distance_mode = ctx->distance_mode;
    J.distance_scale = ctx->distance_scale;
    J.tex_w = ctx->tex_w;
    J.tex_h = ctx->tex_h;
    const uint32_t* old_tex = ctx->d_tex;
    {
        // the last asynchronous read of the allocation must have finished before the kernel may overwrite it
        if (ctx->ev_copied[target]) CU(cudaStreamWaitEvent((cudaStream_t)stream, ctx->ev_copied[target], 0));
        ctx->cur_tex = target;
        ctx->d_tex = ctx->d_tex_pair[target];
    }
    J.albedo = ctx->d_tex;
    J.distance = ctx->d_tex + tex_texels(ctx);
    if (ctx->distance_mode == DDGI_DISTANCE_MOMENTS || ctx->layout == DDGI_LAYOUT_OCTAHEDRAL) ctx->distance_dirty[ctx->cur_tex] = true;
    // peers run the same sequence of modes, so their replicas are clean exactly when this one is;
    // an uploaded plane (ddgi_write_probe_texture) or a double-buffer copy marks it dirty below
    const bool skip_distance = !ctx->distance_dirty[ctx->cur_tex];
    J.albedo_old = old_tex;
    J.albedo_f32 = ctx->debug ? ctx->d_tex_f32 : nullptr;
    J.lookups = ctx->debug ? ctx->d_ray_lookups : nullptr;
    // fused exchange: every texel also goes into the peers' replicas - of the same allocation this
    // update writes locally (the other one of a double-buffered context holds the frame the peers
    // may still be rendering or reading)
    const int peer_buf = ctx->peer_buffers == 2 ? ctx->cur_tex : 0;
    for (int g = 0; g < ctx->n_peers; g++) {
        if (g == ctx->self_index) continue;
        J.peer_albedo[J.n_peers] = (uint32_t*)ctx->peer_base[peer_buf][g];
        J.peer_distance[J.n_peers] = (uint32_t*)ctx->peer_base[peer_buf][g] + tex_texels(ctx);
        J.n_peers++;
    }
    if (ctx->n_peers > 1 && (ctx->peer_buffers == 1 || ctx->in_flight == 2)) {
        // Single-buffered replicas: a faster rank's update i+1 would store into this rank's texture
        // while it still renders or reads frame i.  A second epoch barrier in front of every update
        // closes that window: nobody starts update i+1 before everybody has issued - in stream order,
        // behind its readers of frame i - its own.  (A double-buffered context does not need it: update
        // i+1 writes the other allocation, and update i+2 cannot start before this rank has passed
        // the completion barrier of i+1, which its readers of frame i precede in stream order.  With two
        // frames in flight that chain is gone - update i+2 runs on another stream than barrier i+1 - and
        // this barrier, issued behind the caller's stream (follow_caller), restores it without tying
        // update i+2 to the END of update i+1: a rank arrives here when its readers of frame i are done.)
#ifndef DDGI_TEST_NO_PRE_BARRIER  // (tests/test_multi_gpu_fused.py builds the library once without it: the test must then fail)
        rc = issue_barrier(ctx, 1, (cudaStream_t)stream);
        if (rc) return rc;
#endif
    }
    if (ctx->layout == 1) {
        if (num_rays(ctx) > ctx->ray_out_cap) {
            dfree(ctx->d_ray_out);
            CU(cudaMalloc(&ctx->d_ray_out, num_rays(ctx) * sizeof(float4)));
            ctx->ray_out_cap = num_rays(ctx);
        }
        J.ray_out = ctx->d_ray_out;
    }
    ctx->warp_times_n = 0;
    if (ctx->debug >= 2 && ctx->variant != 0 && P.max_bounces > 0) {
        size_t warps = wavefront_warps(J.n_owned * J.slot_rays, ctx->grid_limit);
        if (warps > ctx->warp_times_cap) {
            dfree(ctx->d_warp_times);
            CU(cudaMalloc(&ctx->d_warp_times, warps * 3 * sizeof(unsigned long long)));
            ctx->warp_times_cap = warps;
        }
        J.warp_times = ctx->d_warp_times;
        ctx->warp_times_n = warps;
    }
    if (skip_distance && ctx->layout != DDGI_LAYOUT_OCTAHEDRAL) J.distance = nullptr;
    int l = 0;
    for (cudaEvent_t* e : {&ctx->ev_k0[target], &ctx->ev_k1[target]})
        if (!*e) CU(cudaEventCreate(e));
    CU(cudaEventRecord(ctx->ev_k0[target], (cudaStream_t)stream));
    CU(launch_probe_update(P, J, ctx->variant, ctx->d_counter + target, ctx->march_min, ctx->grid_limit, (cudaStream_t)stream, &l));
    CU(cudaEventRecord(ctx->ev_k1[target], (cudaStream_t)stream));
    if (!J.distance) J.distance = ctx->d_tex + tex_texels(ctx);
    if (ctx->layout == 1) {
        OctJob O;
        memset(&O, 0, sizeof(O));
        O.probes = ctx->d_owned;
        O.n_probes = (uint32_t)ctx->owned.size();
        O.n_rays = ctx->rx * ctx->ry;
        O.dirs = ctx->d_dirs;
        O.ray_out = ctx->d_ray_out;
        O.tex_w = ctx->tex_w;
        O.albedo = J.albedo;
        O.distance = J.distance;
        O.albedo_old = old_tex;
        O.distance_old = old_tex + tex_texels(ctx);
        O.blend = ctx->blend_mode;
        O.hysteresis = ctx->field.hysteresis;
        O.distance_scale = ctx->distance_scale;
        O.n_peers = J.n_peers;
        for (int g = 0; g < J.n_peers; g++) {
            O.peer_albedo[g] = J.peer_albedo[g];
            O.peer_distance[g] = J.peer_distance[g];
        }
        CU(launch_probe_blend_octahedral(P, O, (cudaStream_t)stream, &l));
    }
    ctx->launches += l;
    ctx->last_stream = (cudaStream_t)stream;
    if (!ctx->ev_update) CU(cudaEventCreateWithFlags(&ctx->ev_update, cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->ev_update, (cudaStream_t)stream));
    if (ctx->ray_mode == 1) {
        if (!ctx->ev_dirs[ctx->dirs_cur]) CU(cudaEventCreateWithFlags(&ctx->ev_dirs[ctx->dirs_cur], cudaEventDisableTiming));
        CU(cudaEventRecord(ctx->ev_dirs[ctx->dirs_cur], (cudaStream_t)stream));
    }
    if (!ctx->ev_frame[target]) CU(cudaEventCreateWithFlags(&ctx->ev_frame[target], cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->ev_frame[target], (cudaStream_t)stream));  // (an exchange records it again behind itself)
    if (calibrate) {
        // First update after the scene / rays / field changed: this launch also recorded the largest
        // voxel-lookup count per slot.  Read them once (the only synchronising probe update) and
        // list the owned slots most expensive first from now on.
        size_t ns = num_slots(ctx);
        std::vector<uint32_t> c(ns);
        CU(cudaMemcpyAsync(c.data(), ctx->d_slot_cost, ns * sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        CU(cudaStreamSynchronize((cudaStream_t)stream));
        if (ctx->cost.size() != ns) ctx->cost.assign(ns, 0xffffffffu);
        for (uint32_t q : ctx->order) ctx->cost[q] = c[q];  // only the owned slots were measured
        ctx->calibrated = true;
        ctx->order_dirty = true;
    }
    return DDGI_OK;
}

int ddgi_render_frame(ddgi_ctx* ctx, void* stream)
{
    if (!ctx) return DDGI_E_INVALID;
    if (!ctx->have_field || !ctx->d_tex) return fail(ctx, DDGI_E_STATE, "no irradiance field");
    if (!ctx->d_occ) return fail(ctx, DDGI_E_STATE, "no voxel field");
    if (!ctx->have_cam) return fail(ctx, DDGI_E_STATE, "no camera");
    if (!ctx->d_frame) return fail(ctx, DDGI_E_STATE, "no frame: set render settings with a non-empty screen");
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_debug_buffers(ctx);
    if (rc) return rc;
    FrameParams P;
    fill_params(ctx, &P);
    PixelJob J;
    memset(&J, 0, sizeof(J));
    J.albedo = ctx->d_tex;
    J.distance = ctx->d_tex + tex_texels(ctx);
    J.tex_w = ctx->tex_w;
    // rows of 16x16 workgroups: floor(h/16) in all (rvpt.cpp:1139-1140), split evenly over the bands
    int groups = ctx->rs.screen_height / 16;
    J.group_row0 = (int)((long long)groups * ctx->band_rank / ctx->band_world);
    J.group_rows = (int)((long long)groups * (ctx->band_rank + 1) / ctx->band_world) - J.group_row0;
    J.frame = ctx->d_frame;
    J.frame_f32 = ctx->debug ? ctx->d_frame_f32 : nullptr;
    J.lookups = ctx->debug ? ctx->d_px_lookups : nullptr;
    int l = 0;
    if (ctx->in_flight == 2 && ctx->ev_frame[ctx->cur_tex]) CU(cudaStreamWaitEvent((cudaStream_t)stream, ctx->ev_frame[ctx->cur_tex], 0));
    CU(launch_render_frame(P, J, (cudaStream_t)stream, &l));
    ctx->launches += l;
    ctx->last_stream = (cudaStream_t)stream;
    return DDGI_OK;
}

int ddgi_set_frame_band(ddgi_ctx* ctx, int32_t rank, int32_t world)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(world >= 1 && rank >= 0 && rank < world, "bad band rank / world");
    ctx->band_rank = rank;
    ctx->band_world = world;
    return DDGI_OK;
}

int ddgi_frame_band_rows(const ddgi_ctx* ctx, int32_t* y0, int32_t* y1)
{
    if (!ctx || !y0 || !y1) return DDGI_E_INVALID;
    int groups = ctx->rs.screen_height / 16;
    *y0 = 16 * (int)((long long)groups * ctx->band_rank / ctx->band_world);
    *y1 = 16 * (int)((long long)groups * (ctx->band_rank + 1) / ctx->band_world);
    return DDGI_OK;
}

int ddgi_sync(ddgi_ctx* ctx)
{
    if (!ctx) return DDGI_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    // the stream of the last dispatch and the engine's own copy stream - not the device: the caller's
    // other streams keep running (raytrace_work_fence.wait, rvpt.cpp:277, waits for one submission too)
    CU(cudaStreamSynchronize(ctx->last_stream));
    int rc = drain_frames(ctx);
    if (rc) return rc;
    if (ctx->copy_stream) CU(cudaStreamSynchronize(ctx->copy_stream));
    return DDGI_OK;
}

int ddgi_probe_texture_size(const ddgi_ctx* ctx, int32_t* width, int32_t* height)
{
    if (!ctx || !width || !height) return DDGI_E_INVALID;
    *width = ctx->tex_w;
    *height = ctx->tex_h;
    return DDGI_OK;
}

int ddgi_read_probe_texture(ddgi_ctx* ctx, int32_t which, int32_t fmt, void* dst, size_t bytes)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->d_tex, "no probe texture");
    NEED(dst && (which == 0 || which == 1), "bad arguments");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->last_stream));  // the dispatches may run on a non-blocking stream
    size_t n = tex_texels(ctx);
    if (fmt == DDGI_FMT_RGBA8) {
        NEED(bytes == n * 4, "expected width*height*4 bytes");
        CU(cudaMemcpy(dst, ctx->d_tex + (which ? n : 0), bytes, cudaMemcpyDeviceToHost));
        return DDGI_OK;
    }
    if (fmt == DDGI_FMT_F32) {
        NEED(which == 0 && ctx->d_tex_f32, "fp32 copy exists only for the albedo texture in debug mode");
        NEED(bytes == n * 16, "expected width*height*16 bytes");
        CU(cudaMemcpy(dst, ctx->d_tex_f32, bytes, cudaMemcpyDeviceToHost));
        return DDGI_OK;
    }
    return fail(ctx, DDGI_E_INVALID, "unknown format");
}

int ddgi_set_double_buffer(ddgi_ctx* ctx, int32_t on)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->n_peers == 0, "close peers first: the peers have mapped this context's allocations");
    NEED(on || ctx->in_flight == 1, "two frames are in flight: ddgi_set_frames_in_flight(ctx, 1) first");
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    bool want = on != 0;
    if (want == ctx->double_buffer) return DDGI_OK;
    ctx->double_buffer = want;
    if (!ctx->d_tex) return DDGI_OK;  // no field yet: resize_textures allocates the pair
    size_t bytes = (2 * tex_texels(ctx) + kFlagWords) * sizeof(uint32_t);
    if (want) {
        int other = ctx->cur_tex ^ 1;
        CU(cudaMalloc(&ctx->d_tex_pair[other], bytes));
        CU(cudaMemcpy(ctx->d_tex_pair[other], ctx->d_tex, bytes, cudaMemcpyDeviceToDevice));
        ctx->distance_dirty[other] = ctx->distance_dirty[ctx->cur_tex];
    } else {
        dfree(ctx->d_tex_pair[ctx->cur_tex ^ 1]);
    }
    return DDGI_OK;
}

int ddgi_read_probe_texture_rows_async(ddgi_ctx* ctx, int32_t which, int32_t row0, int32_t row1, void* dst, size_t bytes)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->d_tex, "no probe texture");
    NEED(dst && (which == 0 || which == 1) && row0 >= 0 && row0 <= row1 && row1 <= ctx->tex_h, "bad rows");
    NEED(bytes == (size_t)(row1 - row0) * ctx->tex_w * 4, "expected (row1 - row0) * width * 4 bytes");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    int b = ctx->cur_tex;
    if (!ctx->ev_copied[b]) CU(cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming));
    // after everything the dispatch stream has been given so far (the update that produced this buffer and,
    // under an exchange, the barrier / collective that completed it), on the engine's own copy stream
    if (!ctx->ev_ready) CU(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->ev_ready, ctx->last_stream));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_ready, 0));
    const uint32_t* src = ctx->d_tex + (which ? tex_texels(ctx) : 0) + (size_t)row0 * ctx->tex_w;
    if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CU(cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
    return DDGI_OK;
}

int ddgi_read_probe_texture_async(ddgi_ctx* ctx, int32_t which, void* dst, size_t bytes)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(bytes == tex_texels(ctx) * 4, "expected width*height*4 bytes");
    return ddgi_read_probe_texture_rows_async(ctx, which, 0, ctx->tex_h, dst, bytes);
}

int ddgi_read_wait(ddgi_ctx* ctx)
{
    if (!ctx) return DDGI_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (ctx->copy_stream) CU(cudaStreamSynchronize(ctx->copy_stream));
    return DDGI_OK;
}

int ddgi_write_probe_texture(ddgi_ctx* ctx, int32_t which, const void* src, size_t bytes)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->d_tex, "no probe texture");
    NEED(src && (which == 0 || which == 1) && bytes == tex_texels(ctx) * 4, "expected width*height*4 bytes");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->last_stream));
    {
        int rc = drain_frames(ctx);
        if (rc) return rc;
    }
    CU(cudaMemcpy(ctx->d_tex + (which ? tex_texels(ctx) : 0), src, bytes, cudaMemcpyHostToDevice));
    if (which == 1) ctx->distance_dirty[ctx->cur_tex] = true;
    return DDGI_OK;
}

int ddgi_read_frame(ddgi_ctx* ctx, int32_t fmt, void* dst, size_t bytes)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ctx->d_frame && dst, "no frame");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->last_stream));  // the dispatches may run on a non-blocking stream
    size_t n = (size_t)ctx->frame_w * ctx->frame_h;
    if (fmt == DDGI_FMT_RGBA8) {
        NEED(bytes == n * 4, "expected w*h*4 bytes");
        CU(cudaMemcpy(dst, ctx->d_frame, bytes, cudaMemcpyDeviceToHost));
        return DDGI_OK;
    }
    if (fmt == DDGI_FMT_F32) {
        NEED(ctx->d_frame_f32, "fp32 frame exists only in debug mode");
        NEED(bytes == n * 16, "expected w*h*16 bytes");
        CU(cudaMemcpy(dst, ctx->d_frame_f32, bytes, cudaMemcpyDeviceToHost));
        return DDGI_OK;
    }
    return fail(ctx, DDGI_E_INVALID, "unknown format");
}

int ddgi_set_debug(ddgi_ctx* ctx, int32_t debug)
{
    if (!ctx) return DDGI_E_INVALID;
    ctx->debug = debug < 0 ? 0 : (debug > 2 ? 2 : debug);
    return DDGI_OK;
}

int ddgi_read_warp_times(ddgi_ctx* ctx, uint64_t* dst, size_t count, size_t* n_warps)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(n_warps, "null n_warps");
    *n_warps = ctx->warp_times_n;
    if (!dst) return DDGI_OK;
    NEED(count >= ctx->warp_times_n * 3, "expected 3 values per warp");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpy(dst, ctx->d_warp_times, ctx->warp_times_n * 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return DDGI_OK;
}

int ddgi_set_schedule_slot(ddgi_ctx* ctx, int32_t rays)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(rays >= 0 && rays <= 4096, "slot must be 0 (a whole probe) or a ray count that divides rays/probe");
    ctx->slot_pref = rays;
    ctx->order_dirty = true;
    ctx->calibrated = false;
    return DDGI_OK;
}

int ddgi_set_grid_limit(ddgi_ctx* ctx, int32_t blocks_per_sm)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(blocks_per_sm >= 0 && blocks_per_sm <= 32, "blocks_per_sm in [0,32]");
    ctx->grid_limit = blocks_per_sm;
    return DDGI_OK;
}

int ddgi_read_lookup_counts(ddgi_ctx* ctx, int32_t which, uint32_t* dst, size_t count)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(dst, "null dst");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->last_stream));  // the dispatches may run on a non-blocking stream
    if (which == 0) {
        NEED(ctx->d_ray_lookups && count == num_rays(ctx) && ctx->ray_lookups_cap == count, "no per-ray counts (debug mode off, or no update since the ray set changed)");
        CU(cudaMemcpy(dst, ctx->d_ray_lookups, count * 4, cudaMemcpyDeviceToHost));
        return DDGI_OK;
    }
    NEED(ctx->d_px_lookups && count == (size_t)ctx->frame_w * ctx->frame_h, "no per-pixel counts (debug mode off?)");
    CU(cudaMemcpy(dst, ctx->d_px_lookups, count * 4, cudaMemcpyDeviceToHost));
    return DDGI_OK;
}

int ddgi_set_kernel_variant(ddgi_ctx* ctx, int32_t variant)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(variant == 0 || variant == 1 || variant == 2, "variant must be 0, 1 or 2");
    ctx->variant = variant;
    return DDGI_OK;
}

int ddgi_set_tuning(ddgi_ctx* ctx, int32_t march_min)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(march_min >= 1 && march_min <= 32, "march_min in [1,32]");
    ctx->march_min = march_min;
    return DDGI_OK;
}

int ddgi_set_color_mode(ddgi_ctx* ctx, int32_t mode)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(mode == DDGI_COLOR_PALETTE || mode == DDGI_COLOR_LITERAL, "color mode must be DDGI_COLOR_PALETTE or DDGI_COLOR_LITERAL");
    ctx->color_mode = mode;
    return DDGI_OK;
}

int ddgi_set_blend_mode(ddgi_ctx* ctx, int32_t mode)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(mode == DDGI_BLEND_OVERWRITE || mode == DDGI_BLEND_HYSTERESIS, "blend mode must be DDGI_BLEND_OVERWRITE or DDGI_BLEND_HYSTERESIS");
    ctx->blend_mode = mode;
    return DDGI_OK;
}

int ddgi_set_weight_mode(ddgi_ctx* ctx, int32_t mode)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(mode == DDGI_WEIGHT_LITERAL || mode == DDGI_WEIGHT_CHEBYSHEV, "weight mode must be DDGI_WEIGHT_LITERAL or DDGI_WEIGHT_CHEBYSHEV");
    ctx->weight_mode = mode;
    return DDGI_OK;
}

int ddgi_set_distance_mode(ddgi_ctx* ctx, int32_t mode, float scale)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(mode == DDGI_DISTANCE_ZERO || mode == DDGI_DISTANCE_MOMENTS, "distance mode must be DDGI_DISTANCE_ZERO or DDGI_DISTANCE_MOMENTS");
    NEED(scale > 0.0f && scale < 1e30f, "distance scale must be positive and finite");
    ctx->distance_mode = mode;
    ctx->distance_scale = scale;
    return DDGI_OK;
}

int ddgi_set_frames_in_flight(ddgi_ctx* ctx, int32_t n)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(n == 1 || n == 2, "1 or 2 frames in flight");
    NEED(n == 1 || ctx->double_buffer, "two frames in flight need the double-buffered texture (ddgi_set_double_buffer)");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->last_stream));
    int rc = drain_frames(ctx);
    if (rc) return rc;
    if (n == 2) {
        for (int b = 0; b < 2; b++)
            if (!ctx->frame_stream[b]) CU(cudaStreamCreateWithFlags(&ctx->frame_stream[b], cudaStreamNonBlocking));
        if (!ctx->ev_in) CU(cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming));
    }
    ctx->in_flight = n;
    return DDGI_OK;
}

int ddgi_frame_fence(ddgi_ctx* ctx, void* stream)
{
    if (!ctx) return DDGI_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (ctx->in_flight == 2)
        for (int b = 0; b < 2; b++)
            if (ctx->ev_frame[b]) CU(cudaStreamWaitEvent((cudaStream_t)stream, ctx->ev_frame[b], 0));
    return DDGI_OK;
}

int ddgi_last_update_ms(ddgi_ctx* ctx, float* ms)
{
    if (!ctx) return DDGI_E_INVALID;
    NEED(ms, "null ms");
    NEED(ctx->ev_k0[ctx->cur_tex] && ctx->ev_k1[ctx->cur_tex], "no probe update yet");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->ev_k1[ctx->cur_tex]));
    CU(cudaEventElapsedTime(ms, ctx->ev_k0[ctx->cur_tex], ctx->ev_k1[ctx->cur_tex]));
    return DDGI_OK;
}

int ddgi_set_sample_order(ddgi_ctx* ctx, int32_t y_first)
{
    if (!ctx) return DDGI_E_INVALID;
    ctx->sample_y_first = y_first != 0;
    return DDGI_OK;
}

int ddgi_set_auto_schedule(ddgi_ctx* ctx, int32_t on)
{
    if (!ctx) return DDGI_E_INVALID;
    ctx->auto_schedule = on ? 1 : 0;
    ctx->order_dirty = true;
    ctx->calibrated = false;
    return DDGI_OK;
}

}  // extern "C"
