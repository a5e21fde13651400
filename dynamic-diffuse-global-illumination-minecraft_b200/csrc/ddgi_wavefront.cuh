// ddgi_wavefront.cuh — the probe-ray path of ddgi_trace.cuh re-expressed as a per-lane
// state machine whose states a warp executes one at a time, most-populated first.
//
// A probe ray is a chain of nearest-hit queries: per bounce one query along the ray and
// then one shadow feeler per light (assets/shaders/probe_pass.comp:283-295, :180-215).
// Each query is a light-sphere pre-test plus a voxel march of up to 125 steps
// (assets/shaders/intersection.glsl:1244-1301, :1051-1100).  In the reference's nested
// loops a warp serialises every divergent branch of that chain.  Here each lane carries
// its ray as a WfRay and is always in exactly one scheduled state:
//
//   WF_MARCH       wf_step            one DDA advance + voxel test (repeated)
//   WF_MARCH_SLOW  wf_step_literal    the same for irregular directions, literal arithmetic
//   WF_HIT         wf_resolve_hit     a march ended: nearest of light spheres / block, then the
//                                     bounce's hit record and first feeler, or the feeler's
//                                     direct term and the next feeler
//   WF_FETCH       (kernel)           ray finished: store its texel, take the next ray
//
// (WF_SCATTER: wf_scatter, the next bounce direction, and WF_QUERY: wf_begin_query, normalised
// direction and reciprocals, are transient: the pass that produced them runs them before it
// ends.)  The kernel (ddgi_kernels.cu: probe_update_wavefront) keeps marching while enough
// lanes march, else runs the fullest other state for exactly its lanes, so every block of code
// executes with many lanes active instead of once per divergent lane group.  Bounce hits and
// feeler hits share ONE state: what they have in common (the light-sphere test before, aiming a
// feeler and arming the next query after) is issued once for both kinds of lanes.
//
// The arithmetic is the reference's, operation for operation, so results are
// bit-identical to ddgi_trace.cuh and to the oracle.  What differs is only how each
// correctly-rounded result is obtained:
//   * max((-f)/d, (1-f)/d) needs one division: for d > 0 the first quotient is <= 0 <=
//     the second, for d < 0 the other way round (f in [0,1]); zero / NaN / tiny
//     components take the literal two-division form (WF_MARCH_SLOW).
//   * that division and x/0.1f use the FMA-corrected reciprocal of ddgi_fastmath.cuh.
//   * the voxel test reads one bit of the cell's 32-cell brick word (16 MiB for 512^3
//     voxels, L1/L2 resident); the block type is fetched only on a hit.
//   * a light sphere whose discriminant is not positive yields t = INF in the reference
//     (intersection.glsl:100-113), so the two root divisions are skipped for it; and the
//     light spheres are tested after the march, only those that could beat the block hit
//     (one bounding-sphere test rejects all of them for most queries).
//   * normalize() of an axis-aligned unit vector is the identity (1/sqrt(1) = 1), which
//     removes it from the block-hit normal, the feeler's lambert term and the scatter frame.
//   * (FrameParams::early_out, kernel variant 2) a shadow feeler only decides "does a voxel lie in
//     front of the light"; its march is ended a little behind the light and resolved by the
//     target light's own sphere test (wf_resolve_hit), resumed in the rare case that test
//     cannot settle it.  Texels are unchanged; fewer voxel lookups than the reference performs.
//   * the scatter direction of the last bounce is never used and is not computed.
#pragma once
#include "ddgi_fastmath.cuh"
#include "ddgi_trace.cuh"

#ifndef DDGI_STEP_TAIL
#define DDGI_STEP_TAIL 1
#endif

namespace ddgi {

enum : int {
    WF_MARCH = 0,
    WF_HIT = 1,
    WF_FETCH = 2,
    WF_MARCH_SLOW = 3,  // march with a zero / NaN / tiny direction component: literal arithmetic
    WF_QUERY = 4,       // transient
    WF_SCATTER = 5,     // transient
    WF_LIMIT = 6,       // transient: a march ended without a hit (wf_end_march)
    WF_IDLE = 7
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
    // FrameParams::early_out: a shadow feeler's march may end once t exceeds t_stop (its light is
    // behind it; +INF = never, as for every bounce ray); wf_test_cell sets it to -1 when it did
    float t_stop;
    // current query
    v3 qd;       // query direction as given (positions are origin + qd * t)
    float qlen;  // |qd| as wf_begin_query's normalize() evaluated it
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

// The march flavour of the current query: the fast step (ddgi_fastmath.cuh) needs regular direction
// components, no origin component in (0, 2^-70) so that a position is either 0 or >= 2^-98 in
// magnitude, and |origin| < 2^20 so that |p| stays below 2^22 over 125 cells (floor_small /
// add_round_up); anything else marches with the literal arithmetic.
DDGI_HD int wf_march_mode(const WfRay& R)
{
    bool fast = regular_direction(R.md.x, R.md.y, R.md.z) && regular_origin3(R.mo.x, R.mo.y, R.mo.z);
    return fast ? WF_MARCH : WF_MARCH_SLOW;
}

// WF_QUERY: starts the nearest-hit query (R.mo, R.qd): arms the march.  The light spheres
// are tested when the march has ended (light_test).
DDGI_HD void wf_begin_query(const FrameParams& P, WfRay& R)
{
    // normalize(R.qd), keeping the length for light_test
    R.qlen = sqrtf(dot(R.qd, R.qd));
    R.md = R.qd * rcp_exact(R.qlen);
    // (irregular components produce a value that is never used: WF_MARCH_SLOW divides literally)
    R.inv = V3(rcp_regular(R.md.x), rcp_regular(R.md.y), rcp_regular(R.md.z));
    R.sel = V3(R.md.x > 0 ? 1.0f : 0.0f, R.md.y > 0 ? 1.0f : 0.0f, R.md.z > 0 ? 1.0f : 0.0f);
    R.p = R.mo;
    R.t = 0.0f;
    R.steps = 0;
    R.mode = wf_march_mode(R);
}

// The ray is complete: WF_FETCH stores wf_final_color.
DDGI_HD void wf_finish_ray(const FrameParams& P, WfRay& R) { R.mode = WF_FETCH; }
// color /= max_bounces (probe_pass.comp:295), evaluated once where the texel is stored
DDGI_HD v3 wf_final_color(const FrameParams& P, const WfRay& R) { return R.color / (float)P.max_bounces; }

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
    R.t_stop = inf_f();
    R.mo = origin;
    R.qd = direction;
    R.mode = WF_QUERY;
}

// Voxel test at the new position of the literal march.  Returns true while the march
// goes on; when it ends the lane is in WF_HIT (a solid cell: R.t is the hit) or WF_LIMIT.
DDGI_HD bool wf_test_cell(const FrameParams& P, WfRay& R)
{
    R.steps++;
    const bool solid = cell_solid(P.scene, cell_bits(ceilf(R.p.x)), cell_bits(ceilf(R.p.y)), cell_bits(ceilf(R.p.z)));
    const bool limit = R.steps >= kMarchSteps || R.t > R.t_stop;
    if (solid) R.mode = WF_HIT;
    else if (limit) R.mode = WF_LIMIT;
    return !(solid || limit);
}

// WF_LIMIT (transient: the pass that ends a march runs this before anything else looks at the
// lane): the march ended without a solid cell.  After 125 cells there is no block hit, t = INF
// (intersection.glsl:1059, :1098); otherwise a feeler has left its light behind (early_out).
DDGI_HD void wf_end_march(WfRay& R)
{
    if (R.mode != WF_LIMIT) return;
    if (R.steps >= kMarchSteps) R.t = inf_f();
    else R.t_stop = -1.0f;
    R.mode = WF_HIT;
}

// WF_MARCH: one DDA advance and voxel test (the body of the reference's 125-iteration
// loop, intersection.glsl:1059-1097): t += min_a(max((-f_a)/d_a, (1-f_a)/d_a)) + 1e-4 with f = fract(p).
// The larger quotient has numerator 1-f for d > 0 and -f for d < 0 (a zero numerator may come out as
// +0 where the reference has -0: min(..)+1e-4 is the same).  floor: see floor_small.
// (A software-pipelined form - fetch the brick word of the new cell, test it one call later behind the
// next advance's quotients - hid the load but cost one set of quotients per march: slower on three of
// four workloads once the brick layout had raised the L1 hit rate, profiles/r2_ab.md c.)
DDGI_HD bool wf_step(const FrameParams& P, WfRay& R)
{
    float tx = div_markstein(R.sel.x - (R.p.x - floor_small(R.p.x)), R.md.x, R.inv.x);
    float ty = div_markstein(R.sel.y - (R.p.y - floor_small(R.p.y)), R.md.y, R.inv.y);
    float tz = div_markstein(R.sel.z - (R.p.z - floor_small(R.p.z)), R.md.z, R.inv.z);
    R.t += gmin(gmin(tx, ty), tz) + 0.0001f;
    R.p = R.mo + R.md * R.t;
    R.steps++;
    // |p| < 2^22: p + 1.5*2^23 rounded up IS ceil(p) + 1.5*2^23 (one directed-rounding add)
    const bool solid = cell_solid(P.scene, float_bits(add_round_up(R.p.x, kCellMagic)), float_bits(add_round_up(R.p.y, kCellMagic)),
                                  float_bits(add_round_up(R.p.z, kCellMagic)));
#if DDGI_STEP_TAIL == 1
    const bool limit = (R.steps >= kMarchSteps) | (R.t > R.t_stop);
    R.mode = solid ? WF_HIT : (limit ? WF_LIMIT : WF_MARCH);
    return !(solid | limit);
#else
    const bool limit = R.steps >= kMarchSteps || R.t > R.t_stop;
    if (solid) R.mode = WF_HIT;
    else if (limit) R.mode = WF_LIMIT;
    return !(solid || limit);
#endif
}

// WF_MARCH_SLOW: the literal two-division form.
DDGI_HD bool wf_step_literal(const FrameParams& P, WfRay& R)
{
    march_advance(R.mo, R.md, R.t, R.p);
    return wf_test_cell(P, R);
}

// Arms the shadow feeler to light R.phase-1 from the current bounce hit.
DDGI_HD void wf_aim_feeler(const FrameParams& P, WfRay& R)
{
    R.mo = R.hpos;
    v3 w = lpos(P.lights[R.phase - 1]) - R.hpos;
    float len = sqrtf(dot(w, w));
    R.qd = w * rcp_exact(len);  // normalize(w)
    // the light's sphere (radius 0.1) ends about len + 0.1 along the feeler; the margin only decides
    // how often wf_resolve_hit has to resume a march, never a result (a NaN length never stops)
    R.t_stop = P.early_out ? len + 0.25f : inf_f();
    R.mode = WF_QUERY;
}

template <bool kLiteral>
DDGI_HD v3 wf_base_color(const FrameParams& P, const WfRay& R, const float* stash, int stride)
{
    if (R.hblock < 0) return V3(0, 0, 0);  // light sphere: base_color is zero-initialised
    if (kLiteral) return V3(stash[0], stash[stride], stash[2 * stride]);
    return scene_albedo(P.scene, R.hblock);
}

// WF_HIT: the march of the current query ended.  phase 0, the bounce ray (probe_pass.comp:283-295
// with intersect_scene, intersection.glsl:1244-1301): nearest of light sphere / block; on a miss
// the ray is complete, else record the hit and aim the first feeler.  phase i >= 1, the feeler to
// light i-1 (probe_pass.comp:186-212): its direct term or the ambient term, then the next feeler
// or the bounce's sum.  Leaves the lane in WF_QUERY (a feeler to march), WF_SCATTER (the bounce is
// done), WF_FETCH (the ray is done) or, early_out only, back in its march.
// `stash` (3 floats, `stride` apart) keeps the procedural colour of the bounce hit until the
// feelers are resolved (colour mode 1 only; the palette mode re-reads it by block type).
// kLiteral is the colour mode as a compile-time constant: the palette kernel carries none of
// the texture code.  `nearest_t` (optional) receives a bounce query's t (INF on a miss): the
// distance-moment mode keeps it for the ray's first query.
template <bool kLiteral>
DDGI_HD void wf_resolve_hit(const FrameParams& P, WfRay& R, float* stash, int stride, float* nearest_t = nullptr)
{
    const bool feeler = R.phase != 0;
    float closest;
    bool block_hit;
    if (feeler && R.t_stop < 0.0f) {
        // The feeler's march was ended at R.t without a block hit (early_out).  Whatever block the
        // full march would still find lies at t_block >= R.t, and the nearest light-sphere root of
        // the reference's loop (intersection.glsl:1262-1279) is at most the target light's own root
        // ti: with ti < R.t the block loses (`block_hit` false) and some light is hit (`closest`
        // finite) - all the feeler needs.  Otherwise (the fp32 sphere test missed the light it was
        // aimed at, or the margin was too small) the march resumes where it stopped.
        v3 n;
        float ti = light_sphere(R.mo, R.qd, P.lights[R.phase - 1], inf_f(), &n);
        if (!(ti < R.t)) {
            R.t_stop = inf_f();
            R.mode = wf_march_mode(R);
            return;
        }
        closest = ti;
        block_hit = false;
    } else {
        int which;
        closest = light_test(P, R.mo, R.qd, R.qlen, R.t, &which, nullptr);
        block_hit = R.t < closest;
        if (block_hit) closest = R.t;
    }
    R.lookups += (uint32_t)R.steps;
    bool aim = false;
    if (!feeler) {
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
            normal = face_normal_axis(R.p, cell);
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
            int which;
            light_test(P, R.mo, R.qd, R.qlen, inf_f(), &which, &n);
            normal = normalize(n);
            R.hblock = -1;
        }
        R.hpos = (R.mo + R.qd * closest) + normal * 0.001f;
        R.hnormal = normal;
        R.direct = V3(0, 0, 0);
        R.visible = 0;
        if (P.n_lights == 0) {
            R.mode = WF_SCATTER;  // direct term 0
        } else {
            R.phase = 1;
            aim = true;
        }
    } else {
        const Light& l = P.lights[R.phase - 1];
        bool blocked = false;
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
                blocked = true;
            }
        }
        if (!blocked) {
            R.phase++;
            if (R.phase > P.n_lights) {
                v3 result = V3(0, 0, 0);
                if (R.visible != 0) {
                    v3 base = wf_base_color<kLiteral>(P, R, stash, stride);
                    result = (base * R.direct) / (float)R.visible;
                }
                R.color = R.color + result;
                R.mode = WF_SCATTER;
            } else {
                aim = true;
            }
        }
    }
    if (aim) wf_aim_feeler(P, R);
}

// WF_SCATTER: the bounce's direct term is in; pick the next bounce direction
// (probe_pass.comp:292) or finish after max_bounces.
DDGI_HD void wf_scatter(const FrameParams& P, WfRay& R)
{
    R.bounce++;
    if (R.bounce >= P.max_bounces) {
        // (the reference draws the last bounce's direction too, probe_pass.comp:292; nothing reads it)
        wf_finish_ray(P, R);
        return;
    }
    R.mo = R.hpos + R.hnormal * 0.0001f;
    R.qd = hemisphere_dir(R.hnormal, R.rng, R.hblock >= 0);
    R.phase = 0;
    R.t_stop = inf_f();
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
            case WF_LIMIT: wf_end_march(R); break;
            case WF_QUERY: wf_begin_query(P, R); break;
            case WF_HIT: {
                float* ft = (R.bounce == 0 && R.phase == 0) ? first_t : nullptr;
                if (P.scene.color_mode != 0) wf_resolve_hit<true>(P, R, stash, 1, ft);
                else wf_resolve_hit<false>(P, R, stash, 1, ft);
                break;
            }
            default: wf_scatter(P, R); break;
        }
    }
    lookups += R.lookups;
    return wf_final_color(P, R);
}

}  // namespace ddgi
