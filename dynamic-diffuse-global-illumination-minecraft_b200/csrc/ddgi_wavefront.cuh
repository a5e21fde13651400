// ddgi_wavefront.cuh — the probe-ray path of ddgi_trace.cuh re-expressed as a per-lane
// state machine whose states a warp executes one at a time, most-populated first.
//
// A probe ray is a chain of nearest-hit queries: per bounce one query along the ray and
// then one shadow feeler per light (assets/shaders/probe_pass.comp:283-295, :180-215).
// Each query is a light-sphere pre-test plus a voxel march of up to 125 steps
// (assets/shaders/intersection.glsl:1244-1301, :1051-1100).  In the reference's nested
// loops a warp serialises every divergent branch of that chain.  Here each lane carries
// its ray as a WfRay and is always in exactly one state:
//
//   WF_QUERY       wf_begin_query     normalise direction, reciprocals
//   WF_MARCH       wf_step            one DDA advance + voxel test (repeated)
//   WF_MARCH_SLOW  wf_step_literal    the same for irregular directions, literal arithmetic
//   WF_BOUNCE_HIT  wf_resolve_bounce  the bounce ray's march ended: hit record, first feeler
//   WF_FEELER_HIT  wf_resolve_feeler  a shadow feeler's march ended: direct term / next feeler
//   WF_SCATTER     wf_scatter         bounce finished: cosine-weighted direction for the next
//   WF_FETCH       (kernel)           ray finished: store its texel, take the next ray
//
// and the kernel (ddgi_kernels.cu: probe_update_wavefront) repeatedly counts the lanes
// per state with ballots and runs the code of the fullest state for exactly those lanes.
// Every block of code therefore executes with many lanes active instead of once per
// divergent lane group.
//
// The arithmetic is the reference's, operation for operation, so results are
// bit-identical to ddgi_trace.cuh and to the oracle.  What differs is only how each
// correctly-rounded result is obtained:
//   * max((-f)/d, (1-f)/d) needs one division: for d > 0 the first quotient is <= 0 <=
//     the second, for d < 0 the other way round (f in [0,1]); zero / NaN / tiny
//     components take the literal two-division form (WfRay::slow).
//   * that division and x/0.1f use the FMA-corrected reciprocal of ddgi_fastmath.cuh.
//   * the voxel test reads one bit of the 4x4x2-brick occupancy word (16 MiB for 512^3
//     voxels, L1/L2 resident); the block type is fetched only on a hit.
//   * a light sphere whose discriminant is not positive yields t = INF in the reference
//     (intersection.glsl:100-113), so the two root divisions are skipped for it; and the
//     light spheres are tested after the march, only those that could beat the block hit
//     (one bounding-sphere test rejects all of them for most queries).
//   * normalize() of an axis-aligned unit vector is the identity (1/sqrt(1) = 1), which
//     removes it from the block-hit normal, the feeler's lambert term and the scatter frame.
#pragma once
#include "ddgi_fastmath.cuh"
#include "ddgi_trace.cuh"

namespace ddgi {

enum : int {
    WF_MARCH = 0,
    WF_QUERY = 1,
    WF_BOUNCE_HIT = 2,
    WF_FEELER_HIT = 3,
    WF_SCATTER = 4,
    WF_FETCH = 5,
    WF_MARCH_SLOW = 6,  // march with a zero / NaN / tiny direction component: literal arithmetic
    WF_IDLE = 7,
    WF_NUM_STATES = 7  // schedulable states (IDLE excluded)
};

struct WfRay {
    // current march
    v3 mo;   // query origin
    v3 md;   // normalize(query direction)
    v3 inv;  // 1 / md (WF_MARCH only)
    v3 sel;  // 1 where md > 0 else 0: numerator of the larger quotient is sel - fract(p)
    v3 p;    // position after the last advance
    float t; // march parameter; +INF once the march has ended without a hit
    int steps;
    int mode;
    int hit_mode;  // state a finished march hands over to: WF_BOUNCE_HIT or WF_FEELER_HIT
    // current query
    v3 qd;  // query direction as given (positions are origin + qd * t)
    // path
    int bounce;
    int phase;  // 0: the bounce ray itself; i >= 1: shadow feeler to light i-1
    v3 hpos, hnormal;
    int hblock;  // block type of the bounce hit, -1 for a light sphere (albedo 0)
    v3 direct;
    int visible;
    v3 color;
    uint32_t rng;
    uint32_t lookups;
};

// Light-sphere test of a query: nearest t over all lights, exactly as the loop of
// intersect_scene (intersection.glsl:1262-1279) evaluates it, except that it runs AFTER
// the march and skips lights that cannot beat the block hit at t_block:
// a root of |w + dir*t| = 0.1 (w = origin - light) has t*|dir| >= |w| - 0.1, and the
// reference's fp32 evaluation of it is within 1e-5 of that for |w| >= 0.101 (no
// cancellation: B^2/(A*C) <= 50), so with |w| > 1.01*t_block*|dir| + 0.101 every root it
// could report is > t_block: the block wins whatever the light's t is.  Skipping a light
// only widens the (0, closest) window of later ones by values > t_block, which lose
// to the block as well; with no block hit (t_block = INF) nothing is skipped.
// `normal` (optional) receives the un-normalised sphere normal of the winning light.
DDGI_HD float light_test(const FrameParams& P, v3 origin, v3 direction, float t_block, int* which, v3* normal)
{
    float closest = inf_f();
    *which = -1;
    float reach2 = inf_f();
    if (t_block < inf_f()) {
        float reach = (t_block * sqrtf(dot(direction, direction))) * 1.01f + 0.101f;
        reach2 = reach * reach;
        // all lights at once: they lie within lights_radius of lights_centre, so
        // |w_i| >= |origin - centre| - radius for every i; 1.0001 covers the fp32 evaluation
        v3 wc = origin - V3(P.lights_centre[0], P.lights_centre[1], P.lights_centre[2]);
        float far = reach + P.lights_radius;
        if (dot(wc, wc) > (far * far) * 1.0001f) return closest;
    }
    bool scaled = false;
    v3 d = V3(0, 0, 0);
    float A = 0.0f;
    for (int i = 0; i < P.n_lights; i++) {
        v3 w = origin - lpos(P.lights[i]);
        if (dot(w, w) > reach2) continue;
        if (!scaled) {
            d = div_tenth(direction);
            A = dot(d, d);
            scaled = true;
        }
        v3 o = div_tenth(w);
        float B = -dot(d, o);
        float C = dot(o, o) - 1.0f;
        float D = B * B - A * C;
        if (!(D > 0)) continue;  // D -> INF: both roots fail the (0, maxt) window, t = INF
        D = sqrtf(D);
        float t1 = (B - D) / A;
        float t2 = (B + D) / A;
        t1 = (0.0f < t1 && t1 < closest) ? t1 : inf_f();
        t2 = (0.0f < t2 && t2 < closest) ? t2 : inf_f();
        float t = gmin(t1, t2);
        if (t < closest) {
            *which = i;
            if (normal) *normal = o + d * t;
        }
        closest = gmin(t, closest);
    }
    return closest;
}

// WF_QUERY: starts the nearest-hit query (R.mo, R.qd): arms the march.  The light spheres
// are tested when the march has ended (light_test).
DDGI_HD void wf_begin_query(const FrameParams& P, WfRay& R)
{
    v3 origin = R.mo;
    R.md = normalize(R.qd);
    // fast-step preconditions (ddgi_fastmath.cuh): regular direction components, and no
    // origin component in (0, 2^-70) so that a position is either 0 or >= 2^-98 in magnitude
    // and |origin| < 2^20 so that |p| stays below 2^22 over 125 cells (floor_small / add_round_up)
    bool slow = !(regular_component(R.md.x) && regular_component(R.md.y) && regular_component(R.md.z)) ||
                !(regular_origin(origin.x) && regular_origin(origin.y) && regular_origin(origin.z));
    // (irregular components produce a value that is never used: WF_MARCH_SLOW divides literally)
    R.inv = V3(rcp_regular(R.md.x), rcp_regular(R.md.y), rcp_regular(R.md.z));
    R.sel = V3(R.md.x > 0 ? 1.0f : 0.0f, R.md.y > 0 ? 1.0f : 0.0f, R.md.z > 0 ? 1.0f : 0.0f);
    R.p = origin;
    R.t = 0.0f;
    R.steps = 0;
    R.hit_mode = R.phase == 0 ? WF_BOUNCE_HIT : WF_FEELER_HIT;
    R.mode = slow ? WF_MARCH_SLOW : WF_MARCH;
}

// The ray is complete: final colour, then WF_FETCH stores it.
DDGI_HD void wf_finish_ray(const FrameParams& P, WfRay& R)
{
    R.color = R.color / (float)P.max_bounces;
    R.mode = WF_FETCH;
}

// Path state of a fresh ray with first query (origin, direction).
DDGI_HD void wf_init(WfRay& R, v3 origin, v3 direction, uint32_t ray_index)
{
    R.rng = wang_hash(ray_index);
    R.color = V3(0, 0, 0);
    R.direct = V3(0, 0, 0);
    R.visible = 0;
    R.bounce = 0;
    R.phase = 0;
    R.lookups = 0;
    R.hpos = R.hnormal = V3(0, 0, 0);
    R.hblock = -1;
    R.mo = origin;
    R.qd = direction;
    R.mode = WF_QUERY;
}

// Voxel test at the new position, shared by both march flavours.
DDGI_HD void wf_test_cell(const FrameParams& P, WfRay& R, bool small_coords)
{
    R.steps++;
    int kx, ky, kz;
    if (small_coords) {
        // |p| < 2^22: p + 1.5*2^23 rounded up IS ceil(p) + 1.5*2^23 (one directed-rounding add)
        kx = float_bits(add_round_up(R.p.x, kCellMagic));
        ky = float_bits(add_round_up(R.p.y, kCellMagic));
        kz = float_bits(add_round_up(R.p.z, kCellMagic));
    } else {
        kx = cell_bits(ceilf(R.p.x));
        ky = cell_bits(ceilf(R.p.y));
        kz = cell_bits(ceilf(R.p.z));
    }
    if (cell_solid(P.scene, kx, ky, kz)) {
        R.mode = R.hit_mode;
    } else if (R.steps >= kMarchSteps) {
        R.t = inf_f();  // no block within 125 cells
        R.mode = R.hit_mode;
    }
}

// WF_MARCH: one DDA advance and voxel test (the body of the reference's 125-iteration
// loop): t += min_a(max((-f_a)/d_a, (1-f_a)/d_a)) + 1e-4 with f = fract(p).  The larger
// quotient has numerator 1-f for d > 0 and -f for d < 0 (a zero numerator may come out as
// +0 where the reference has -0: min(..)+1e-4 is the same).
// floor(p) for |p| < 2^22 comes from two full-rate adds (no conversion-pipe FRND).
DDGI_HD void wf_step(const FrameParams& P, WfRay& R)
{
    float tx = div_markstein(R.sel.x - (R.p.x - floor_small(R.p.x)), R.md.x, R.inv.x);
    float ty = div_markstein(R.sel.y - (R.p.y - floor_small(R.p.y)), R.md.y, R.inv.y);
    float tz = div_markstein(R.sel.z - (R.p.z - floor_small(R.p.z)), R.md.z, R.inv.z);
    R.t += gmin(gmin(tx, ty), tz) + 0.0001f;
    R.p = R.mo + R.md * R.t;
    wf_test_cell(P, R, true);
}

// WF_MARCH_SLOW: the literal two-division form.
DDGI_HD void wf_step_literal(const FrameParams& P, WfRay& R)
{
    march_advance(R.mo, R.md, R.t, R.p);
    wf_test_cell(P, R, false);
}

// Arms the shadow feeler to light R.phase-1 from the current bounce hit.
DDGI_HD void wf_aim_feeler(const FrameParams& P, WfRay& R)
{
    R.mo = R.hpos;
    R.qd = normalize(lpos(P.lights[R.phase - 1]) - R.hpos);
    R.mode = WF_QUERY;
}

// WF_BOUNCE_HIT: the bounce ray's march ended.  Nearest of light sphere / block; on a
// miss the ray is complete, else record the hit and aim the first feeler.
// `stash` (3 floats, `stride` apart) keeps the procedural colour of the bounce hit until the
// feelers are resolved (colour mode 1 only; the palette mode re-reads it by block type).
// kLiteral is the colour mode as a compile-time constant: the palette kernel carries none of
// the texture code.  `nearest_t` (optional) receives the query's t (INF on a miss): the
// distance-moment mode keeps it for the ray's first query.
template <bool kLiteral>
DDGI_HD void wf_resolve_bounce(const FrameParams& P, WfRay& R, float* stash, int stride, float* nearest_t = nullptr)
{
    R.lookups += (uint32_t)R.steps;
    int which;
    float closest = light_test(P, R.mo, R.qd, R.t, &which, nullptr);
    bool block_hit = R.t < closest;
    if (block_hit) closest = R.t;
    if (nearest_t) *nearest_t = closest;
    if (!(closest < inf_f())) {
        wf_finish_ray(P, R);
        return;
    }
    v3 normal;
    if (block_hit) {
        // face_normal is axis-aligned (or zero for a NaN position): both normalize() calls
        // of the reference (grid_march :1088, intersect_scene :1294) are identities on it
        v3 cell = V3(ceilf(R.p.x), ceilf(R.p.y), ceilf(R.p.z));
        normal = face_normal_unit(R.p, cell);
        R.hblock = scene_type_at(P.scene, cell);
        if (kLiteral) {
            v3 c = block_color_literal(R.p, R.hblock, normal);
            stash[0] = c.x;
            stash[stride] = c.y;
            stash[2 * stride] = c.z;
        }
    } else {
        // a light sphere is the nearest hit (rare): redo the test for its normal
        v3 n;
        light_test(P, R.mo, R.qd, inf_f(), &which, &n);
        normal = normalize(n);
        R.hblock = -1;
    }
    R.hpos = (R.mo + R.qd * closest) + normal * 0.001f;
    R.hnormal = normal;
    R.direct = V3(0, 0, 0);
    R.visible = 0;
    if (P.n_lights == 0) {
        R.mode = WF_SCATTER;  // direct term 0
        return;
    }
    R.phase = 1;
    wf_aim_feeler(P, R);
}

// WF_FEELER_HIT: the feeler to light R.phase-1 ended (probe_pass.comp:186-212).
template <bool kLiteral>
DDGI_HD v3 wf_base_color(const FrameParams& P, const WfRay& R, const float* stash, int stride)
{
    if (R.hblock < 0) return V3(0, 0, 0);  // light sphere: base_color is zero-initialised
    if (kLiteral) return V3(stash[0], stash[stride], stash[2 * stride]);
    return scene_albedo(P.scene, R.hblock);
}

template <bool kLiteral>
DDGI_HD void wf_resolve_feeler(const FrameParams& P, WfRay& R, const float* stash, int stride)
{
    R.lookups += (uint32_t)R.steps;
    int which;
    float closest = light_test(P, R.mo, R.qd, R.t, &which, nullptr);
    bool block_hit = R.t < closest;
    if (block_hit) closest = R.t;
    const Light& l = P.lights[R.phase - 1];
    if (closest < inf_f()) {
        // normalize(info.normal) (probe_pass.comp:194): a block-hit normal is an axis-aligned
        // unit vector (or all-NaN), a fixed point of normalize
        v3 n = R.hblock >= 0 ? R.hnormal : normalize(R.hnormal);
        float lambert = gclamp(dot(n, R.qd), 0.0f, 1.0f);
        if (!block_hit) {
            float dist = length(lpos(l) - R.hpos);
            R.direct = R.direct + ((lcol(l) * lambert) * l.intensity) / dist;
            R.visible++;
        } else {
            // blocked by a voxel: ambient term, remaining lights are skipped
            v3 base = wf_base_color<kLiteral>(P, R, stash, stride);
            R.color = R.color + (base * 0.2f) * lambert;
            R.mode = WF_SCATTER;
            return;
        }
    }
    R.phase++;
    if (R.phase > P.n_lights) {
        v3 result = V3(0, 0, 0);
        if (R.visible != 0) {
            v3 base = wf_base_color<kLiteral>(P, R, stash, stride);
            result = (base * R.direct) / (float)R.visible;
        }
        R.color = R.color + result;
        R.mode = WF_SCATTER;
        return;
    }
    wf_aim_feeler(P, R);
}

// WF_SCATTER: the bounce's direct term is in; pick the next bounce direction
// (probe_pass.comp:292) or finish after max_bounces.
DDGI_HD void wf_scatter(const FrameParams& P, WfRay& R)
{
    v3 o = R.hpos + R.hnormal * 0.0001f;
    v3 d = hemisphere_dir(R.hnormal, R.rng, R.hblock >= 0);
    R.bounce++;
    if (R.bounce >= P.max_bounces) {
        wf_finish_ray(P, R);
        return;
    }
    R.phase = 0;
    R.mo = o;
    R.qd = d;
    R.mode = WF_QUERY;
}

// Scalar driver (tests/hostsim): the state machine stepped for a single ray.
DDGI_HD v3 wavefront_trace_scalar(const FrameParams& P, v3 origin, v3 direction, uint32_t ray_index,
                                  uint32_t& lookups, float* first_t = nullptr)
{
    if (first_t) *first_t = 0.0f;
    WfRay R;
    float stash[3] = {0, 0, 0};
    wf_init(R, origin, direction, ray_index);
    if (P.max_bounces <= 0) wf_finish_ray(P, R);
    while (R.mode != WF_FETCH) {
        switch (R.mode) {
            case WF_MARCH: wf_step(P, R); break;
            case WF_MARCH_SLOW: wf_step_literal(P, R); break;
            case WF_QUERY: wf_begin_query(P, R); break;
            case WF_BOUNCE_HIT: {
                float* ft = R.bounce == 0 ? first_t : nullptr;
                if (P.scene.color_mode != 0) wf_resolve_bounce<true>(P, R, stash, 1, ft);
                else wf_resolve_bounce<false>(P, R, stash, 1, ft);
                break;
            }
            case WF_FEELER_HIT:
                if (P.scene.color_mode != 0) wf_resolve_feeler<true>(P, R, stash, 1);
                else wf_resolve_feeler<false>(P, R, stash, 1);
                break;
            default: wf_scatter(P, R); break;
        }
    }
    lookups += R.lookups;
    return R.color;
}

}  // namespace ddgi
