// hostsim.cpp — TEST-ONLY host build of the engine's per-ray device functions.
//
// The kernels' logic lives in host/device-clean headers (csrc/ddgi_*.cuh).  This file
// compiles those same headers with g++ so the CPU-only test tier (`-m "not gpu"`) can
// check them against the oracle without a GPU.  It is never linked into
// libddgi_b200.so and nothing in the product can reach it.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../dynamic-diffuse-global-illumination-minecraft_b200/csrc/ddgi_shade.cuh"
#include "../../dynamic-diffuse-global-illumination-minecraft_b200/csrc/ddgi_wavefront.cuh"

using namespace ddgi;

// same layout as oracle/ddgi_oracle.c OrcParams so the tests build one description
struct SimLight { float intensity; float col[3]; float pos[3]; };
struct SimParams {
    int32_t scene_mode, scene, color_mode, n_lights;
    SimLight lights[8];
    int32_t vdim[3];
    int32_t vorg[3];
    const uint8_t* vox;
    const float* palette;
    int32_t probe_count[3];
    int32_t side_length;
    int32_t rx, ry;
    float field_origin[3];
    int32_t max_bounces;
    int32_t screen_width, screen_height;
    int32_t blend_mode;
    float hysteresis;
    int32_t render_mode, visualize_probes, weight_mode, distance_mode;
    float distance_scale;
    int32_t layout, oct;
};

struct Built {
    FrameParams P;
    std::vector<uint32_t> occ;  // one word per brick
};

static void build(const SimParams* S, const float* cam, Built* B)
{
    FrameParams& P = B->P;
    memset(&P, 0, sizeof(P));
    int nb[3], sh[3];
    for (int a = 0; a < 3; a++) {
        int borg = S->vorg[a] & ~(kBrickAlign - 1);
        sh[a] = S->vorg[a] - borg;
        int cells = 1 << (a == 0 ? kBrickLx : a == 1 ? kBrickLy : kBrickLz);
        nb[a] = (S->vorg[a] + S->vdim[a] - borg + cells - 1) / cells;
        P.scene.vorg[a] = S->vorg[a];
        P.scene.vdim[a] = S->vdim[a];
        P.scene.borg[a] = borg;
        P.scene.kneg[a] = -(kCellBias + borg);
        P.scene.nb[a] = nb[a];
        P.scene.lo[a] = (float)S->vorg[a];
        P.scene.hi[a] = (float)(S->vorg[a] + S->vdim[a] - 1);
        P.probe_count[a] = S->probe_count[a];
        P.field_origin[a] = S->field_origin[a];
    }
    B->occ.assign((size_t)nb[0] * nb[1] * nb[2], 0u);
    for (int z = 0; z < S->vdim[2]; z++)
        for (int y = 0; y < S->vdim[1]; y++)
            for (int x = 0; x < S->vdim[0]; x++)
                if (S->vox[((size_t)z * S->vdim[1] + y) * S->vdim[0] + x]) {
                    int bx = x + sh[0], by = y + sh[1], bz = z + sh[2];
                    B->occ[((size_t)(bz >> kBrickLz) * nb[1] + (by >> kBrickLy)) * nb[0] + (bx >> kBrickLx)] |= occ_mask(occ_shift(bx, by, bz));
                }
    P.scene.occ = B->occ.data();
    P.scene.types = S->vox;
    P.scene.palette = S->palette;
    P.scene.color_mode = S->color_mode == 0 ? 1 : 0;  // OrcParams: 0 = literal colours, 1 = palette
    P.n_lights = S->n_lights;
    for (int i = 0; i < S->n_lights; i++) {
        P.lights[i].intensity = S->lights[i].intensity;
        for (int a = 0; a < 3; a++) {
            P.lights[i].col[a] = S->lights[i].col[a];
            P.lights[i].pos[a] = S->lights[i].pos[a];
        }
    }
    light_bounds(P);
    P.side_length = S->side_length;
    P.rx = S->rx;
    P.ry = S->ry;
    P.max_bounces = S->max_bounces;
    P.screen_w = S->screen_width;
    P.screen_h = S->screen_height;
    P.render_mode = S->render_mode;
    P.visualize_probes = S->visualize_probes;
    P.weight_mode = S->weight_mode;
    P.distance_scale = S->distance_scale;
    P.layout = S->layout;
    P.oct = S->oct;
    P.tile_w = S->layout == 1 ? S->oct : S->rx;
    P.tile_h = S->layout == 1 ? S->oct : S->ry;
    if (cam) {
        memcpy(P.cam, cam, sizeof(P.cam));
        P.cam_w = 1.0f / (float)tan((double)(0.5f * cam[17]));
    }
}

extern "C" {

// variant 0: trace_probe_ray (reference loop order); variant 1: the wavefront
// state machine stepped one lane at a time; variant 2: the same with its result-preserving
// early-outs (FrameParams::early_out: same texels, fewer voxel lookups).
void sim_probe_update(const SimParams* S, const float* rays /* R x 12 */, uint32_t k0, uint32_t k1,
                      int variant, uint32_t* albedo, float* f32, uint32_t* lookups, uint32_t* distance)
{
    Built B;
    build(S, nullptr, &B);
    B.P.early_out = variant == 2;
    const FrameParams& P = B.P;
    int W = P.probe_count[0] * P.probe_count[2] * P.rx;
    int tiles_x = P.probe_count[0] * P.probe_count[2];
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = k0; k < (int64_t)k1; k++) {
        const float* r = rays + 12 * k;
        v3 o = V3(r[0], r[1], r[2]), d = V3(r[4], r[5], r[6]);
        int p = f2i(r[8]);
        int yp = p / tiles_x, xp = p - yp * tiles_x;
        int tx = xp * P.rx + f2i(r[9]), ty = yp * P.ry + f2i(r[10]);
        uint32_t n = 0;
        float first_t = 0.0f;
        v3 c = variant == 0 ? trace_probe_ray(P, o, d, (uint32_t)k, n, &first_t)
                            : wavefront_trace_scalar(P, o, d, (uint32_t)k, n, &first_t);
        size_t t = (size_t)ty * W + tx;
        if (S->blend_mode) c = blend_hysteresis(albedo[t], c, S->hysteresis);
        albedo[t] = pack_rgba8(c.x, c.y, c.z, 1.0f);
        if (distance) {
            uint32_t moments = 0u;
            if (S->distance_mode == 1) {
                float dd = first_t / S->distance_scale;
                moments = pack_rgba8(dd, dd * dd, 0.0f, 0.0f);
            }
            distance[t] = moments;
        }
        if (f32) { f32[4 * t] = c.x; f32[4 * t + 1] = c.y; f32[4 * t + 2] = c.z; f32[4 * t + 3] = 1.0f; }
        if (lookups) lookups[k] = n;
    }
}

void sim_render_frame(const SimParams* S, const float* cam, const uint32_t* tex, const uint32_t* dist_tex, uint32_t* frame,
                      float* f32, uint32_t* lookups)
{
    Built B;
    build(S, cam, &B);
    const FrameParams& P = B.P;
    int W = P.probe_count[0] * P.probe_count[2] * P.tile_w;
    int w = P.screen_w, h = P.screen_h;
#pragma omp parallel for schedule(dynamic, 4)
    for (int gy = 0; gy < (h / 16) * 16; gy++)
        for (int gx = 0; gx < (w / 16) * 16; gx++) {
            float cx = (float)gx / (float)w, cy = (float)gy / (float)h;
            cy = 1.0f - cy;
            v3 o, d;
            pinhole_ray(P, cx, cy, &o, &d);
            uint32_t n = 0;
            bool ext = (P.render_mode >= 1 && P.render_mode <= 5) || P.visualize_probes != 0 || P.weight_mode != 0 || P.layout != 0;
            v3 s = ext ? shade_pixel<true>(P, tex, dist_tex, W, o, d, n) : shade_pixel<false>(P, tex, dist_tex, W, o, d, n);
            s = V3(0, 0, 0) + s;
            size_t at = (size_t)gy * w + gx;
            frame[at] = pack_rgba8(s.x, s.y, s.z, 1.0f);
            if (f32) { f32[4 * at] = s.x; f32[4 * at + 1] = s.y; f32[4 * at + 2] = s.z; f32[4 * at + 3] = 1.0f; }
            if (lookups) lookups[at] = n;
        }
}

// Octahedral layout: the trace (either variant) into a ray buffer, then the blend with the warp's
// 32 lanes and its xor-butterfly emulated in order (ddgi_kernels.cu: probe_blend_octahedral).
void sim_probe_update_oct(const SimParams* S, const float* rays /* R x 12 */, int variant, uint32_t* albedo, uint32_t* distance,
                          uint32_t* lookups)
{
    Built B;
    build(S, nullptr, &B);
    const FrameParams& P = B.P;
    const int n = P.rx * P.ry, oct = P.oct;
    const int W = P.probe_count[0] * P.probe_count[2] * oct;
    const int64_t probes = (int64_t)P.probe_count[0] * P.probe_count[1] * P.probe_count[2];
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t p = 0; p < probes; p++) {
        std::vector<float> dirs(3 * (size_t)n), rad(4 * (size_t)n);
        for (int i = 0; i < n; i++) {
            int64_t k = p * n + i;
            const float* r = rays + 12 * k;
            v3 o = V3(r[0], r[1], r[2]), d = V3(r[4], r[5], r[6]);
            uint32_t cnt = 0;
            float first_t = 0.0f;
            v3 c = variant == 0 ? trace_probe_ray(P, o, d, (uint32_t)k, cnt, &first_t)
                                : wavefront_trace_scalar(P, o, d, (uint32_t)k, cnt, &first_t);
            dirs[3 * i] = d.x; dirs[3 * i + 1] = d.y; dirs[3 * i + 2] = d.z;
            rad[4 * i] = c.x; rad[4 * i + 1] = c.y; rad[4 * i + 2] = c.z; rad[4 * i + 3] = first_t;
            if (lookups) lookups[k] = cnt;
        }
        int cx, cy;
        tile_origin(P, (int)p, &cx, &cy);
        for (int t = 0; t < oct * oct; t++) {
            int u = t % oct, v = t / oct;
            v3 dir_t = oct_texel_dir(u, v, oct);
            OctAcc lane[32], next[32];
            for (int l = 0; l < 32; l++) lane[l] = oct_lane_partial(dir_t, dirs.data(), rad.data(), n, l, S->distance_scale);
            for (int off = 16; off >= 1; off >>= 1) {
                for (int l = 0; l < 32; l++) next[l] = oct_add(lane[l], lane[l ^ off]);
                memcpy(lane, next, sizeof(lane));
            }
            size_t at = (size_t)(cy + v) * W + (cx + u);
            oct_finalize(lane[0], S->blend_mode, S->hysteresis, albedo[at], distance[at], &albedo[at], &distance[at]);
        }
    }
}

void sim_bake_scene(int scene, const int32_t* org, const int32_t* dim, uint8_t* out)
{
#pragma omp parallel for
    for (int z = 0; z < dim[2]; z++)
        for (int y = 0; y < dim[1]; y++)
            for (int x = 0; x < dim[0]; x++)
                out[((size_t)z * dim[1] + y) * dim[0] + x] =
                    (uint8_t)block_procedural(V3((float)(x + org[0]), (float)(y + org[1]), (float)(z + org[2])), scene);
}

// Range checks of the fast march step and the face-normal shortcut, for brute-force comparison with
// their float / literal definitions (tests/test_oracle_math.py).
void sim_regular_checks(const float* x, int n, uint8_t* component, uint8_t* origin, uint8_t* dir3, uint8_t* org3)
{
    for (int i = 0; i < n; i++) {
        component[i] = regular_component(x[i]);
        origin[i] = regular_origin(x[i]);
    }
    for (int i = 0; i + 2 < n; i += 3) {
        dir3[i / 3] = regular_direction(x[i], x[i + 1], x[i + 2]);
        org3[i / 3] = regular_origin3(x[i], x[i + 1], x[i + 2]);
    }
}
void sim_unorm8(float* out256)
{
    for (uint32_t b = 0; b < 256; b++) out256[b] = unorm8_to_float(b);
}
void sim_face_normals(const float* p, const float* cell, int n, float* fast, float* literal)
{
    for (int i = 0; i < n; i++) {
        v3 a = face_normal_axis(V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), V3(cell[3 * i], cell[3 * i + 1], cell[3 * i + 2]));
        v3 b = face_normal_unit(V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), V3(cell[3 * i], cell[3 * i + 1], cell[3 * i + 2]));
        fast[3 * i] = a.x; fast[3 * i + 1] = a.y; fast[3 * i + 2] = a.z;
        literal[3 * i] = b.x; literal[3 * i + 1] = b.y; literal[3 * i + 2] = b.z;
    }
}

void sim_pin_sincos(const float* x, int n, float* s, float* c)
{
    for (int i = 0; i < n; i++) pin_sincos(x[i], &s[i], &c[i]);
}
void sim_pin_acos(const float* x, int n, float* out)
{
    for (int i = 0; i < n; i++) out[i] = pin_acos(x[i]);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// Scheduling model of probe_update_wavefront (profiles/policy_sim.py): the kernel's warp loop
// re-enacted on the host with the same per-lane state functions, `n_warps` warps advanced in
// order of their accumulated cost (so the dynamic ray fetch behaves as on the device), counting
// how often each piece of code is issued and with how many lanes.  Not a test: a tool for
// comparing scheduling rules without a GPU.
//   split_hits = 0: the kernel's rule — march while >= march_min/32 of the live lanes can march,
//                   else the fullest of {HIT, FETCH, MARCH_SLOW}; bounce hits and feeler hits are
//                   one state whose common code (light test, aim, query) is issued once
//   split_hits = 1: round 1's rule — bounce hits and feeler hits are separate states, the fullest runs
//   rays_per_lane K > 1: every lane owns K rays and marches / resolves one of them per issue
//                   (static multiplexing; `cost[C_SWAP]` is charged when a lane changes the ray it marches)
// cost[]: warp instructions per issue of each code piece.
enum { C_MARCH, C_SLOW, C_LIGHT, C_BOUNCE, C_FEELER, C_AIM, C_SCATTER, C_QUERY, C_FETCH, C_ROUND, C_SWAP, C_COUNT };
struct PolicyOut {
    uint64_t issues[C_COUNT];  // issues of each code piece
    uint64_t lanes[C_COUNT];   // lanes active over those issues
    double makespan;           // largest accumulated warp cost
    double busy;               // sum of warp costs
};

extern "C" void sim_wavefront_policy(const SimParams* S, const float* rays, const uint32_t* order, uint32_t n_rays, int n_warps,
                                     int split_hits, int march_min, int rays_per_lane, const double* cost, PolicyOut* out)
{
    const int K = rays_per_lane < 1 ? 1 : (rays_per_lane > 4 ? 4 : rays_per_lane);
    Built B;
    build(S, nullptr, &B);
    const FrameParams& P = B.P;
    struct Lane {
        WfRay R[4];
        int cur = 0;  // the ray whose march state the lane holds in registers
    };
    struct Warp {
        Lane L[32];
        double t = 0;
        bool done = false;
    };
    std::vector<Warp> warps((size_t)n_warps);
    for (auto& w : warps)
        for (int l = 0; l < 32; l++)
            for (int k = 0; k < 4; k++) w.L[l].R[k].mode = k < K ? WF_FETCH : WF_IDLE;
    memset(out, 0, sizeof(*out));
    uint32_t next = 0;
    float stash[3];
    size_t live = (size_t)n_warps;
    auto issue = [&](Warp* w, int piece, int lanes) {
        if (lanes <= 0) return;
        out->issues[piece]++;
        out->lanes[piece] += (uint64_t)lanes;
        w->t += cost[piece];
    };
    // the ray of lane L in `state` (the current one first), or -1
    auto pick = [&](Lane& L, int state, bool feeler_only, bool bounce_only) {
        for (int j = 0; j < K; j++) {
            int k = (L.cur + j) % K;
            const WfRay& R = L.R[k];
            if (R.mode != state) continue;
            if (state == WF_HIT && feeler_only && R.phase == 0) continue;
            if (state == WF_HIT && bounce_only && R.phase != 0) continue;
            return k;
        }
        return -1;
    };
    while (live) {
        Warp* w = nullptr;
        for (auto& c : warps)
            if (!c.done && (!w || c.t < w->t)) w = &c;
        int n_live = 0, n_march = 0, n_hit = 0, n_hit_b = 0, n_hit_f = 0, n_fetch = 0, n_slow = 0;
        for (int l = 0; l < 32; l++) {
            Lane& L = w->L[l];
            bool any = false;
            for (int k = 0; k < K; k++) any = any || L.R[k].mode != WF_IDLE;
            n_live += any;
            n_march += pick(L, WF_MARCH, false, false) >= 0;
            n_hit += pick(L, WF_HIT, false, false) >= 0;
            n_hit_b += pick(L, WF_HIT, false, true) >= 0;
            n_hit_f += pick(L, WF_HIT, true, false) >= 0;
            n_fetch += pick(L, WF_FETCH, false, false) >= 0;
            n_slow += pick(L, WF_MARCH_SLOW, false, false) >= 0;
        }
        if (n_live == 0) {
            w->done = true;
            live--;
            continue;
        }
        int enough = (n_live * march_min + 31) >> 5;
        if (enough < 1) enough = 1;
        if (n_march >= enough || n_hit + n_fetch + n_slow == 0) {
            int swaps = 0;
            for (int l = 0; l < 32; l++) {
                Lane& L = w->L[l];
                int k = pick(L, WF_MARCH, false, false);
                if (k < 0) continue;
                if (k != L.cur) {
                    swaps++;
                    L.cur = k;
                }
                wf_step(P, L.R[k]);
                wf_end_march(L.R[k]);
            }
            issue(w, C_SWAP, swaps);
            issue(w, C_MARCH, n_march);
            continue;
        }
        issue(w, C_ROUND, 32);
        int n_scatter = 0, n_query = 0, n_aim = 0;
        auto after = [&](WfRay& R) {
            if (R.mode == WF_SCATTER) {
                wf_scatter(P, R);
                n_scatter++;
            }
            if (R.mode == WF_QUERY) {
                wf_begin_query(P, R);
                n_query++;
            }
        };
        auto run_hits = [&](bool feeler_only, bool bounce_only) {
            int nl = 0, nb = 0, nf = 0;
            for (int l = 0; l < 32; l++) {
                Lane& L = w->L[l];
                int k = pick(L, WF_HIT, feeler_only, bounce_only);
                if (k < 0) continue;
                WfRay& R = L.R[k];
                nl++;
                (R.phase == 0 ? nb : nf)++;
                wf_resolve_hit<false>(P, R, stash, 1);
                if (R.mode == WF_QUERY) n_aim++;
                after(R);
            }
            issue(w, C_LIGHT, nl);
            issue(w, C_BOUNCE, nb);
            issue(w, C_FEELER, nf);
        };
        if (n_slow > n_hit && n_slow > n_fetch) {
            for (int l = 0; l < 32; l++) {
                int k = pick(w->L[l], WF_MARCH_SLOW, false, false);
                if (k >= 0) {
                    wf_step_literal(P, w->L[l].R[k]);
                    wf_end_march(w->L[l].R[k]);
                }
            }
            issue(w, C_SLOW, n_slow);
        } else if (!split_hits ? n_hit > n_fetch : (n_hit_b > n_fetch || n_hit_f > n_fetch)) {
            if (!split_hits) run_hits(false, false);
            else if (n_hit_f >= n_hit_b) run_hits(true, false);
            else run_hits(false, true);
        } else {
            for (int l = 0; l < 32; l++) {
                Lane& L = w->L[l];
                int k = pick(L, WF_FETCH, false, false);
                if (k < 0) continue;
                WfRay& R = L.R[k];
                if (next < n_rays) {
                    uint32_t kk = order[next++];
                    const float* r = rays + 12 * (size_t)kk;
                    wf_init(R, V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), kk);
                    after(R);
                } else {
                    R.mode = WF_IDLE;
                }
            }
            issue(w, C_FETCH, n_fetch);
        }
        issue(w, C_AIM, n_aim);
        issue(w, C_SCATTER, n_scatter);
        issue(w, C_QUERY, n_query);
    }
    for (auto& c : warps) {
        out->busy += c.t;
        if (c.t > out->makespan) out->makespan = c.t;
    }
}

// Per-ray work profile for schedule experiments (profiles/policy_sim.py): counts[4k..4k+3] =
// voxel lookups, nearest-hit queries, bounce resolves, feeler resolves of ray k.
extern "C" void sim_ray_profile(const SimParams* S, const float* rays, uint32_t n_rays, uint32_t* counts)
{
    Built B;
    build(S, nullptr, &B);
    const FrameParams& P = B.P;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t k = 0; k < (int64_t)n_rays; k++) {
        const float* r = rays + 12 * k;
        WfRay R;
        float stash[3] = {0, 0, 0};
        wf_init(R, V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), (uint32_t)k);
        uint32_t q = 0, b = 0, f = 0;
        while (R.mode != WF_FETCH) {
            switch (R.mode) {
                case WF_MARCH: wf_step(P, R); break;
                case WF_MARCH_SLOW: wf_step_literal(P, R); break;
                case WF_LIMIT: wf_end_march(R); break;
                case WF_QUERY: wf_begin_query(P, R); q++; break;
                case WF_HIT: (R.phase == 0 ? b : f)++; wf_resolve_hit<false>(P, R, stash, 1); break;
                default: wf_scatter(P, R); break;
            }
        }
        counts[4 * k] = R.lookups;
        counts[4 * k + 1] = q;
        counts[4 * k + 2] = b;
        counts[4 * k + 3] = f;
    }
}
