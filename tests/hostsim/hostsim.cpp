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
#include "../../dynamic-diffuse-global-illumination-minecraft_b200/csrc/ddgi_pooled.cuh"
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
    std::vector<uint32_t> occ;  // one word per 4x4x2 brick
};

static void build(const SimParams* S, const float* cam, Built* B)
{
    FrameParams& P = B->P;
    memset(&P, 0, sizeof(P));
    int nb[3], sh[3];
    for (int a = 0; a < 3; a++) {
        int borg = S->vorg[a] & ~3;
        sh[a] = S->vorg[a] - borg;
        int cells = a == 2 ? 2 : 4;
        nb[a] = (S->vorg[a] + S->vdim[a] - borg + cells - 1) / cells;
        P.scene.vorg[a] = S->vorg[a];
        P.scene.vdim[a] = S->vdim[a];
        P.scene.borg[a] = borg;
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
                    B->occ[((size_t)(bz >> 1) * nb[1] + (by >> 2)) * nb[0] + (bx >> 2)] |=
                        1u << ((bx & 3) | ((by & 3) << 2) | ((bz & 1) << 4));
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

// Variant 2 (ddgi_pooled.cuh) for one ray: the state machine with the ray living in its packed
// pool record — after every state execution it is packed, and the next execution unpacks it into
// a WfRay whose every byte was poisoned first, the march through its own smaller record and in
// sessions of at most 3 steps.  A field missing from a record cannot go unnoticed.
static v3 pooled_trace_scalar(const FrameParams& P, v3 origin, v3 direction, uint32_t ray_index, uint32_t& lookups, float* first_t_out)
{
    if (P.max_bounces <= 0) return wavefront_trace_scalar(P, origin, direction, ray_index, lookups, first_t_out);  // the engine runs variant 0 then
    PoolVec rec[kPoolVecs];
    memset(rec, 0xFF, sizeof(rec));
    int mode;
    {
        WfRay R;
        memset(&R, 0xFF, sizeof(R));
        wf_init(R, origin, direction, ray_index);
        wf_begin_query(P, R);
        pool_pack(R, ray_index, 0.0f, rec);
        mode = R.mode;
    }
    float stash[3] = {0, 0, 0};
    for (;;) {
        WfRay R;
        memset(&R, 0xFF, sizeof(R));
        if (mode == WF_MARCH || mode == WF_MARCH_SLOW) {
            pool_unpack_march(rec, R);
            R.mode = mode;
            for (int i = 0; i < 3 && R.mode == mode; i++) {
                if (mode == WF_MARCH) wf_step(P, R);
                else wf_step_literal(P, R);
            }
            pool_pack_march(R, rec);
            mode = R.mode;
            continue;
        }
        uint32_t k;
        float first_t;
        pool_unpack(rec, R, k, first_t);
        R.mode = mode;
        if (mode == WF_FETCH) {
            lookups += R.lookups;
            if (first_t_out) *first_t_out = first_t;
            return R.color;
        }
        if (mode == WF_BOUNCE_HIT) {
            float nearest = 0.0f;
            bool first = R.bounce == 0;
            wf_resolve_bounce<false>(P, R, stash, 1, &nearest);
            if (first) first_t = nearest;
        } else {
            wf_resolve_feeler<false>(P, R, stash, 1);
        }
        if (R.mode == WF_SCATTER) wf_scatter(P, R);
        if (R.mode == WF_QUERY) wf_begin_query(P, R);
        pool_pack(R, k, first_t, rec);
        mode = R.mode;
    }
}

extern "C" {

// variant 0: trace_probe_ray (reference loop order); variant 1: the wavefront
// state machine stepped one lane at a time.
void sim_probe_update(const SimParams* S, const float* rays /* R x 12 */, uint32_t k0, uint32_t k1,
                      int variant, uint32_t* albedo, float* f32, uint32_t* lookups, uint32_t* distance)
{
    Built B;
    build(S, nullptr, &B);
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
        v3 c = variant == 0   ? trace_probe_ray(P, o, d, (uint32_t)k, n, &first_t)
               : variant == 1 ? wavefront_trace_scalar(P, o, d, (uint32_t)k, n, &first_t)
                              : pooled_trace_scalar(P, o, d, (uint32_t)k, n, &first_t);
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
// Scheduling-policy model of probe_update_wavefront (profiles/policy_sim.py): the kernel's
// warp loop re-enacted on the host with the same per-lane state functions, `n_warps` warps
// advanced in order of their accumulated cost (so the dynamic ray fetch behaves as on the
// device), counting how often each state's code runs and with how many lanes.  Not a test:
// a tool for comparing scheduling rules without a GPU.
//   policy 0: the kernel's rule (march while >= march_min/32 of the live lanes march, else the
//             fullest other state, ties to the later stage)
//   policy 1: always the fullest state, MARCH counted like any other (ties: march)
//   policy 2: as 0, but a non-march state runs only with >= min_other lanes unless no lane marches
//   policy 3: hysteresis: start marching at >= min_other/32 of the live lanes, keep marching down to march_min/32
//   policy 4: drain: below the march threshold run the other states until none of them holds
//             min_other or more lanes (re-ranked after each), only then look at the march count again
struct PolicyOut {
    uint64_t exec[8];    // executions of each state's code (warp passes)
    uint64_t lanes[8];   // lanes active over those executions
    uint64_t passes;     // outer-loop passes (scheduler rounds)
    double makespan;     // largest accumulated warp cost
    double busy;         // sum of warp costs
};
// `group` > 1 models ideal regrouping inside a block of `group` warps: the scheduling unit holds
// 32 * group rays and a state's code is issued ceil(count / 32) times (moving ray state between
// lanes is taken as free: an upper bound on what block-level compaction could give).
template <int kGroup>
static void wavefront_policy(const SimParams* S, const float* rays, const uint32_t* order, uint32_t n_rays, int n_warps, int policy,
                             int march_min, int min_other, const double* cost, PolicyOut* out);

extern "C" void sim_wavefront_policy(const SimParams* S, const float* rays, const uint32_t* order, uint32_t n_rays, int n_warps,
                                     int policy, int march_min, int min_other, const double* cost /* [8] + scheduler */,
                                     PolicyOut* out, int group)
{
    if (group == 2) wavefront_policy<2>(S, rays, order, n_rays, n_warps, policy, march_min, min_other, cost, out);
    else if (group == 4) wavefront_policy<4>(S, rays, order, n_rays, n_warps, policy, march_min, min_other, cost, out);
    else if (group == 8) wavefront_policy<8>(S, rays, order, n_rays, n_warps, policy, march_min, min_other, cost, out);
    else wavefront_policy<1>(S, rays, order, n_rays, n_warps, policy, march_min, min_other, cost, out);
}

template <int kGroup>
static void wavefront_policy(const SimParams* S, const float* rays, const uint32_t* order, uint32_t n_rays, int n_warps, int policy,
                             int march_min, int min_other, const double* cost, PolicyOut* out)
{
    constexpr int kLanes = 32 * kGroup;
    Built B;
    build(S, nullptr, &B);
    const FrameParams& P = B.P;
    struct Warp {
        WfRay R[kLanes];
        double t = 0;
        bool done = false;
        bool marching = false;  // policy 3
        bool draining = false;  // policy 4
    };
    std::vector<Warp> warps((size_t)n_warps);
    for (auto& w : warps)
        for (int l = 0; l < kLanes; l++) w.R[l].mode = WF_FETCH;
    memset(out, 0, sizeof(*out));
    uint32_t next = 0;
    float stash[3];
    size_t live = (size_t)n_warps;
    while (live) {
        // the warp that is furthest behind runs its next pass
        Warp* w = nullptr;
        for (auto& c : warps)
            if (!c.done && (!w || c.t < w->t)) w = &c;
        int count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int l = 0; l < kLanes; l++) count[w->R[l].mode]++;
        int n_live = kLanes - count[WF_IDLE];
        if (n_live == 0) {
            w->done = true;
            live--;
            continue;
        }
        out->passes++;
        w->t += cost[8];
        auto run = [&](int state) {
            const int issues = (count[state] + 31) / 32;
            out->exec[state] += (uint64_t)issues;
            out->lanes[state] += (uint64_t)count[state];
            w->t += cost[state] * issues;
            // regrouping is not free: moving a ray's state between the pool and a lane's registers
            // costs `cost[7]` instructions per issue (a quarter of it per march step: a warp keeps
            // its rays in registers over a run of steps)
            if (kGroup > 1) w->t += (state == WF_MARCH ? 0.25 : 1.0) * cost[7] * issues;
            int n_scatter = 0, n_query = 0;
            for (int l = 0; l < kLanes; l++) {
                WfRay& R = w->R[l];
                if (R.mode != state) continue;
                switch (state) {
                    case WF_MARCH: wf_step(P, R); break;
                    case WF_MARCH_SLOW: wf_step_literal(P, R); break;
                    case WF_BOUNCE_HIT: wf_resolve_bounce<false>(P, R, stash, 1); break;
                    case WF_FEELER_HIT: wf_resolve_feeler<false>(P, R, stash, 1); break;
                    case WF_FETCH:
                        if (next < n_rays) {
                            uint32_t k = order[next++];
                            const float* r = rays + 12 * (size_t)k;
                            wf_init(R, V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), k);
                        } else {
                            R.mode = WF_IDLE;
                        }
                        break;
                }
                // scatter and query are armed in the same pass, as in the kernel
                if (R.mode == WF_SCATTER) {
                    wf_scatter(P, R);
                    n_scatter++;
                }
                if (R.mode == WF_QUERY) {
                    wf_begin_query(P, R);
                    n_query++;
                }
            }
            if (n_scatter) {
                out->exec[WF_SCATTER] += (uint64_t)((n_scatter + 31) / 32);
                out->lanes[WF_SCATTER] += (uint64_t)n_scatter;
                w->t += cost[WF_SCATTER] * ((n_scatter + 31) / 32);
            }
            if (n_query) {
                out->exec[WF_QUERY] += (uint64_t)((n_query + 31) / 32);
                out->lanes[WF_QUERY] += (uint64_t)n_query;
                w->t += cost[WF_QUERY] * ((n_query + 31) / 32);
            }
        };
        auto fullest_other = [&]() {
            int best = -1, bc = 0;
            for (int s : {WF_BOUNCE_HIT, WF_FEELER_HIT, WF_FETCH, WF_MARCH_SLOW})
                if (count[s] > bc || (count[s] == bc && bc > 0 && s > best)) {
                    best = s;
                    bc = count[s];
                }
            return best;
        };
        if (policy == 1) {
            int o = fullest_other();
            if (count[WF_MARCH] > 0 && (o < 0 || count[WF_MARCH] >= count[o])) run(WF_MARCH);
            else if (o >= 0) run(o);
            continue;
        }
        if (policy == 3) {
            int lo = n_live * march_min > 32 ? n_live * march_min : 32, hi = n_live * min_other > 32 ? n_live * min_other : 32;
            int m32 = count[WF_MARCH] * 32;
            if (w->marching ? m32 >= lo : m32 >= hi) {
                w->marching = true;
                run(WF_MARCH);
                w->t -= cost[8];
                out->passes--;
                continue;
            }
            w->marching = false;
            int o = fullest_other();
            if (o >= 0) run(o);
            else if (count[WF_MARCH] > 0) run(WF_MARCH);
            continue;
        }
        if (policy == 4 && w->draining) {
            int o = fullest_other();
            if (o >= 0 && count[o] >= min_other) {
                run(o);
                continue;
            }
            w->draining = false;
        }
        const int enough = n_live * march_min > 32 ? n_live * march_min : 32;  // march_min/32 of the live lanes
        if (count[WF_MARCH] * 32 >= enough) {
            run(WF_MARCH);  // (one step per pass here; the kernel's inner loop re-checks the same condition)
            w->t -= cost[8];  // the inner march loop does not pay a scheduler round
            out->passes--;
            continue;
        }
        int o = fullest_other();
        if (o < 0 || (policy == 2 && count[o] < min_other && count[WF_MARCH] > 0)) {
            if (count[WF_MARCH] > 0) run(WF_MARCH);
            continue;
        }
        if (policy == 4) w->draining = true;
        run(o);
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
                case WF_QUERY: wf_begin_query(P, R); q++; break;
                case WF_BOUNCE_HIT: wf_resolve_bounce<false>(P, R, stash, 1); b++; break;
                case WF_FEELER_HIT: wf_resolve_feeler<false>(P, R, stash, 1); f++; break;
                default: wf_scatter(P, R); break;
            }
        }
        counts[4 * k] = R.lookups;
        counts[4 * k + 1] = q;
        counts[4 * k + 2] = b;
        counts[4 * k + 3] = f;
    }
}

// ---------------------------------------------------------------------------------------
// Host re-enactment of probe_update_pooled's BLOCK logic (ddgi_kernels.cu) — queues, claims,
// march sessions, fetch / retire accounting — with the block's warps taking turns.  It cannot show
// races, but a slot that is lost or a pool that never drains shows up here, not as a hung GPU.
// Writes the same texels as sim_probe_update; returns the number of loop passes (0 = did not end).
// stats (optional, 16 values): [q] issues per queue, [5 + q] lanes over those issues, [10] march iterations,
// [11] lanes over march iterations, [12] passes that found no work.
// pool_slots: ray slots per block (the kernel has 128 = one per thread); lockstep != 0: the block's
// warps all claim before any of them appends (every warp is always holding the rays it works on, as
// on the device) instead of taking turns.
extern "C" uint64_t sim_probe_update_pooled_stats(const SimParams* S, const float* rays, uint32_t n_rays, int n_blocks, int march_keep,
                                                  uint32_t* albedo, uint32_t* lookups_out, uint64_t* stats, int pool_slots, int lockstep);
extern "C" uint64_t sim_probe_update_pooled(const SimParams* S, const float* rays, uint32_t n_rays, int n_blocks, int march_keep,
                                            uint32_t* albedo, uint32_t* lookups_out)
{
    return sim_probe_update_pooled_stats(S, rays, n_rays, n_blocks, march_keep, albedo, lookups_out, nullptr, 128, 0);
}
extern "C" uint64_t sim_probe_update_pooled_stats(const SimParams* S, const float* rays, uint32_t n_rays, int n_blocks, int march_keep,
                                                  uint32_t* albedo, uint32_t* lookups_out, uint64_t* stats, int pool_slots, int lockstep)
{
    uint64_t local_stats[16] = {0};
    if (!stats) stats = local_stats;
    Built B;
    build(S, nullptr, &B);
    const FrameParams& P = B.P;
    const int W = P.probe_count[0] * P.probe_count[2] * P.rx;
    const int tiles_x = P.probe_count[0] * P.probe_count[2];
    constexpr int WARPS = 4;
    const int N = pool_slots;
    struct Block {
        std::vector<PoolVec> ray;        // N x kPoolVecs
        std::vector<int> queue[PQ_COUNT];  // rings of N
        unsigned head[PQ_COUNT] = {0, 0, 0, 0, 0}, tail[PQ_COUNT] = {0, 0, 0, 0, 0};
        int live = 0;
        bool warp_done[WARPS] = {false, false, false, false};
    };
    std::vector<Block> blocks((size_t)n_blocks);
    for (auto& b : blocks) {
        b.ray.assign((size_t)N * kPoolVecs, PoolVec{0, 0, 0, 0});
        for (int qi = 0; qi < PQ_COUNT; qi++) b.queue[qi].assign((size_t)N, 0);
        for (int i = 0; i < N; i++) {
            b.queue[PQ_FETCH][i] = i;
            b.ray[(size_t)i * kPoolVecs + 6].w = pool_bits_f(0xffffffffu);
        }
        b.tail[PQ_FETCH] = (unsigned)N;
        b.live = N;
    }
    uint32_t next = 0;
    uint64_t passes = 0;
    float stash[3] = {0, 0, 0};
    size_t running = (size_t)n_blocks * WARPS;
    const uint64_t limit = 64ull * (uint64_t)n_rays * 400ull / 32ull + 100000ull;
    struct Claim {
        int q = -1;
        unsigned n = 0;
        int slots[32];
    };
    auto claim = [&](Block& b, Claim& c) {
        unsigned bestc = 0;
        c.q = -1;
        const int order[PQ_COUNT] = {PQ_MARCH, PQ_FETCH, PQ_FEELER, PQ_BOUNCE, PQ_SLOW};
        for (int qi : order) {
            unsigned cnt = b.tail[qi] - b.head[qi];
            unsigned cc = cnt < 32u ? cnt : 32u;
            if (cc > bestc) {
                bestc = cc;
                c.q = qi;
            }
        }
        if (c.q < 0) return;
        unsigned old = b.head[c.q];
        c.n = bestc;
        b.head[c.q] += c.n;
        for (unsigned l = 0; l < c.n; l++) c.slots[l] = b.queue[c.q][(old + l) % (unsigned)N];
        stats[c.q]++;
        stats[5 + c.q] += c.n;
    };
    auto work = [&](Block& b, const Claim& c) {
        const int q = c.q;
        const unsigned n = c.n;
        int newq[32];
        bool push[32];
        auto rec = [&](unsigned l) { return &b.ray[(size_t)c.slots[l] * kPoolVecs]; };
        if (q == PQ_MARCH || q == PQ_SLOW) {
            const int run_mode = q == PQ_MARCH ? WF_MARCH : WF_MARCH_SLOW;
            WfRay R[32];
            for (unsigned l = 0; l < n; l++) {
                memset(&R[l], 0xFF, sizeof(WfRay));
                pool_unpack_march(rec(l), R[l]);
                R[l].mode = run_mode;
            }
            unsigned active;
            do {
                active = 0;
                stats[10]++;
                for (unsigned l = 0; l < n; l++) {
                    if (R[l].mode == run_mode) {
                        stats[11]++;
                        if (q == PQ_MARCH) wf_step(P, R[l]);
                        else wf_step_literal(P, R[l]);
                    }
                    active += R[l].mode == run_mode;
                }
            } while (active && active * 32u >= n * (unsigned)march_keep);
            for (unsigned l = 0; l < n; l++) {
                pool_pack_march(R[l], rec(l));
                newq[l] = pool_queue_of(R[l].mode);
                push[l] = true;
            }
        } else if (q == PQ_FETCH) {
            uint32_t base = next;
            next += n;
            for (unsigned l = 0; l < n; l++) {
                PoolVec* r = rec(l);
                uint32_t k = pool_f_bits(r[6].w);
                if (k != 0xffffffffu) {
                    const float* ry = rays + 12 * (size_t)k;
                    int p = f2i(ry[8]);
                    int yp = p / tiles_x, xp = p - yp * tiles_x;
                    size_t t = (size_t)(yp * P.ry + f2i(ry[10])) * W + (xp * P.rx + f2i(ry[9]));
                    albedo[t] = pack_rgba8(r[8].x, r[8].y, r[8].z, 1.0f);
                    if (lookups_out) lookups_out[k] = pool_f_bits(r[2].w);
                }
                uint32_t idx = base + l;
                push[l] = idx < n_rays;
                newq[l] = PQ_FETCH;
                if (push[l]) {
                    const float* ry = rays + 12 * (size_t)idx;
                    WfRay R;
                    memset(&R, 0xFF, sizeof(R));
                    wf_init(R, V3(ry[0], ry[1], ry[2]), V3(ry[4], ry[5], ry[6]), idx);
                    wf_begin_query(P, R);
                    pool_pack(R, idx, 0.0f, r);
                    newq[l] = pool_queue_of(R.mode);
                } else {
                    r[6].w = pool_bits_f(0xffffffffu);
                    b.live--;
                }
            }
        } else {
            for (unsigned l = 0; l < n; l++) {
                PoolVec* r = rec(l);
                WfRay R;
                memset(&R, 0xFF, sizeof(R));
                uint32_t k;
                float first_t;
                pool_unpack(r, R, k, first_t);
                R.mode = q == PQ_BOUNCE ? WF_BOUNCE_HIT : WF_FEELER_HIT;
                if (q == PQ_BOUNCE) {
                    bool first = R.bounce == 0;
                    float nearest = 0.0f;
                    wf_resolve_bounce<false>(P, R, stash, 1, &nearest);
                    if (first) first_t = nearest;
                } else {
                    wf_resolve_feeler<false>(P, R, stash, 1);
                }
                if (R.mode == WF_SCATTER) wf_scatter(P, R);
                if (R.mode == WF_QUERY) wf_begin_query(P, R);
                pool_pack(R, k, first_t, r);
                newq[l] = pool_queue_of(R.mode);
                push[l] = true;
            }
        }
        for (unsigned l = 0; l < n; l++)
            if (push[l]) {
                int qi = newq[l];
                b.queue[qi][b.tail[qi] % (unsigned)N] = c.slots[l];
                b.tail[qi]++;
            }
    };
    size_t turn = 0;
    while (running) {
        if (++passes > limit) return 0;
        Block& b = blocks[turn++ % blocks.size()];
        Claim claims[WARPS];
        if (lockstep) {
            for (int w = 0; w < WARPS; w++)
                if (!b.warp_done[w]) claim(b, claims[w]);
            for (int w = 0; w < WARPS; w++) {
                if (b.warp_done[w]) continue;
                if (claims[w].q >= 0) work(b, claims[w]);
            }
            for (int w = 0; w < WARPS; w++) {
                if (b.warp_done[w] || claims[w].q >= 0) continue;
                if (b.live <= 0) {
                    b.warp_done[w] = true;
                    running--;
                } else {
                    stats[12]++;
                }
            }
        } else {
            for (int w = 0; w < WARPS; w++) {
                if (b.warp_done[w]) continue;
                claim(b, claims[w]);
                if (claims[w].q >= 0) {
                    work(b, claims[w]);
                } else if (b.live <= 0) {
                    b.warp_done[w] = true;
                    running--;
                } else {
                    stats[12]++;
                }
            }
        }
    }
    return passes;
}
