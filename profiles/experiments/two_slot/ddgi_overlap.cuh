// ddgi_overlap.cuh — the probe-ray state machine of ddgi_wavefront.cuh with TWO marches in flight per lane:
// the shadow feelers of bounce k and the ray of bounce k+1.
//
// A probe ray is the chain  B0, F0.., B1, F1.., ...  (B = the bounce ray's march, F = its hit's shadow feelers,
// one per light until one is blocked; assets/shaders/probe_pass.comp:283-295, :180-215).  But the next bounce
// ray depends on the bounce HIT alone - its origin is the hit, its direction the next two numbers of the ray's
// random sequence, which the feelers never touch (probe_pass.comp:150-178, :292) - not on what the feelers find:
// those only add to the colour.  So once bounce k is resolved, its feelers and the ray of bounce k+1 are marched
// at the same time, in two march slots of the lane:
//
//   slot B (WfRay, as before)   the bounce ray's march and the path state
//   slot F (WfFeeler)           the march of the current feeler of the LAST resolved bounce hit (origin = R.hpos)
//
// The hit of bounce k+1 is resolved only when slot F is idle, i.e. when bounce k's direct term has been added to
// the colour: the additions happen in bounce order and the hit record (hpos, hnormal, hblock, the stash of its
// procedural colour) has one owner at a time, so the result is bit-identical to the sequential chain - the same
// arithmetic on the same values in the same order per accumulator; only the interleaving of independent work
// differs.  What it buys: the dependent chain of a ray is max(B, F) instead of B + F per bounce, a lane is in the
// march loop whenever EITHER slot marches, and the two steps of an iteration are independent instruction
// streams (the occupancy loads of one hide behind the arithmetic of the other).
#pragma once
#include "ddgi_wavefront.cuh"

namespace ddgi {

enum : int { WF_F_IDLE = 0, WF_F_MARCH = 1, WF_F_SLOW = 2, WF_F_HIT = 3 };
// (R.mode additionally takes WF_WAIT: the bounce chain is over - a miss is impossible here, see ov_resolve_bounce -
// after max_bounces; the ray is complete once slot F is idle)
enum : int { WF_WAIT = 9 };

struct WfFeeler {
    v3 md;   // normalize(query direction)
    v3 inv;  // 1 / md
    v3 p;    // position after the last advance
    float t, t_stop;
    int steps;
    int mode;
};

// Query direction of the feeler to light `phase - 1` from the current bounce hit, exactly as wf_aim_feeler
// computes it (recomputed where it is needed instead of being held across the march).
DDGI_HD v3 ov_feeler_dir(const FrameParams& P, const WfRay& R, float* len)
{
    v3 w = lpos(P.lights[R.phase - 1]) - R.hpos;
    *len = sqrtf(dot(w, w));
    return w * rcp_exact(*len);  // normalize(w)
}

// Arms slot F with the feeler to light R.phase - 1 (wf_aim_feeler + wf_begin_query).
DDGI_HD void ov_begin_feeler(const FrameParams& P, const WfRay& R, WfFeeler& F)
{
    float len;
    v3 qd = ov_feeler_dir(P, R, &len);
    F.t_stop = P.early_out ? len + 0.25f : inf_f();
    float qlen = sqrtf(dot(qd, qd));
    F.md = qd * rcp_exact(qlen);
    F.inv = V3(rcp_regular(F.md.x), rcp_regular(F.md.y), rcp_regular(F.md.z));
    F.p = R.hpos;
    F.t = 0.0f;
    F.steps = 0;
    bool fast = regular_direction(F.md.x, F.md.y, F.md.z) && regular_origin3(R.hpos.x, R.hpos.y, R.hpos.z);
    F.mode = fast ? WF_F_MARCH : WF_F_SLOW;
}

// One DDA advance + voxel test of slot F (wf_step on the feeler's state; origin = R.hpos).
DDGI_HD void ov_step_feeler(const FrameParams& P, const WfRay& R, WfFeeler& F)
{
    const float sx = F.md.x > 0 ? 1.0f : 0.0f, sy = F.md.y > 0 ? 1.0f : 0.0f, sz = F.md.z > 0 ? 1.0f : 0.0f;
    float tx = div_markstein(sx - (F.p.x - floor_small(F.p.x)), F.md.x, F.inv.x);
    float ty = div_markstein(sy - (F.p.y - floor_small(F.p.y)), F.md.y, F.inv.y);
    float tz = div_markstein(sz - (F.p.z - floor_small(F.p.z)), F.md.z, F.inv.z);
    F.t += gmin(gmin(tx, ty), tz) + 0.0001f;
    F.p = R.hpos + F.md * F.t;
    F.steps++;
    const bool solid = cell_solid(P.scene, float_bits(add_round_up(F.p.x, kCellMagic)), float_bits(add_round_up(F.p.y, kCellMagic)),
                                  float_bits(add_round_up(F.p.z, kCellMagic)));
    // a march that ends without a solid cell: after 125 cells there is no block hit (t = INF), else the feeler
    // has left its light behind (early_out: t_stop = -1 marks it for ov_resolve_feeler); wf_end_march folded in
    const bool out_of_steps = F.steps >= kMarchSteps, passed = F.t > F.t_stop;
    if (!solid && (out_of_steps | passed)) {
        if (out_of_steps) F.t = inf_f();
        else F.t_stop = -1.0f;
    }
    F.mode = (solid | out_of_steps | passed) ? WF_F_HIT : WF_F_MARCH;
}

// The literal two-division form for irregular feelers (wf_step_literal).
DDGI_HD void ov_step_feeler_literal(const FrameParams& P, const WfRay& R, WfFeeler& F)
{
    march_advance(R.hpos, F.md, F.t, F.p);
    F.steps++;
    const bool solid = cell_solid(P.scene, cell_bits(ceilf(F.p.x)), cell_bits(ceilf(F.p.y)), cell_bits(ceilf(F.p.z)));
    const bool out_of_steps = F.steps >= kMarchSteps, passed = F.t > F.t_stop;
    if (!solid && (out_of_steps | passed)) {
        if (out_of_steps) F.t = inf_f();
        else F.t_stop = -1.0f;
    }
    F.mode = (solid | out_of_steps | passed) ? WF_F_HIT : WF_F_SLOW;
}

// Slot F's march ended (WF_F_HIT): the feeler branch of wf_resolve_hit.  Leaves slot F marching again (the next
// feeler, or - early_out only - the same one resumed) or idle (the bounce's direct term is in R.color); a ray
// whose bounce chain is over (R.mode == WF_WAIT) is complete then.
template <bool kLiteral>
DDGI_HD void ov_resolve_feeler(const FrameParams& P, WfRay& R, WfFeeler& F, float* stash, int stride)
{
    float len;
    const v3 qd = ov_feeler_dir(P, R, &len);
    float closest;
    bool block_hit;
    if (F.t_stop < 0.0f) {
        // ended behind its light without a block hit (see wf_resolve_hit): settled by the target light's own
        // sphere test, or resumed
        v3 n;
        float ti = light_sphere(R.hpos, qd, P.lights[R.phase - 1], inf_f(), &n);
        if (!(ti < F.t)) {
            F.t_stop = inf_f();
            bool fast = regular_direction(F.md.x, F.md.y, F.md.z) && regular_origin3(R.hpos.x, R.hpos.y, R.hpos.z);
            F.mode = fast ? WF_F_MARCH : WF_F_SLOW;
            return;
        }
        closest = ti;
        block_hit = false;
    } else {
        int which;
        const float qlen = sqrtf(dot(qd, qd));
        closest = light_test(P, R.hpos, qd, qlen, F.t, &which, nullptr);
        block_hit = F.t < closest;
        if (block_hit) closest = F.t;
    }
    R.lookups += (uint32_t)F.steps;
    const Light& l = P.lights[R.phase - 1];
    bool more = false;
    bool blocked = false;
    if (closest < inf_f()) {
        v3 n = R.hblock >= 0 ? R.hnormal : normalize(R.hnormal);
        float lambert = gclamp(dot(n, qd), 0.0f, 1.0f);
        if (!block_hit) {
            float dist = length(lpos(l) - R.hpos);
            R.direct = R.direct + ((lcol(l) * lambert) * l.intensity) / dist;
            R.visible++;
        } else {
            v3 base = wf_base_color<kLiteral>(P, R, stash, stride);
            R.color = R.color + (base * 0.2f) * lambert;
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
        } else {
            more = true;
        }
    }
    if (more) {
        ov_begin_feeler(P, R, F);
    } else {
        F.mode = WF_F_IDLE;
        if (R.mode == WF_WAIT) wf_finish_ray(P, R);
    }
}

// Slot B's march ended (R.mode == WF_HIT) and slot F is idle: the bounce branch of wf_resolve_hit, then at once
// the first feeler (slot F) AND the next bounce ray (wf_scatter + wf_begin_query in slot B).
template <bool kLiteral>
DDGI_HD void ov_resolve_bounce(const FrameParams& P, WfRay& R, WfFeeler& F, float* stash, int stride, float* nearest_t = nullptr)
{
    int which;
    float closest = light_test(P, R.mo, R.qd, R.qlen, R.t, &which, nullptr);
    const bool block_hit = R.t < closest;
    if (block_hit) closest = R.t;
    R.lookups += (uint32_t)R.steps;
    if (nearest_t) *nearest_t = closest;
    if (!(closest < inf_f())) {
        wf_finish_ray(P, R);  // (slot F is idle: nothing is pending)
        return;
    }
    v3 normal;
    if (block_hit) {
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
        v3 n;
        light_test(P, R.mo, R.qd, R.qlen, inf_f(), &which, &n);
        normal = normalize(n);
        R.hblock = -1;
    }
    R.hpos = (R.mo + R.qd * closest) + normal * 0.001f;
    R.hnormal = normal;
    R.direct = V3(0, 0, 0);
    R.visible = 0;
    if (P.n_lights != 0) {
        R.phase = 1;
        ov_begin_feeler(P, R, F);
    }
    // the next bounce ray (wf_scatter), or the end of the bounce chain
    R.bounce++;
    if (R.bounce >= P.max_bounces) {
        if (F.mode == WF_F_IDLE) wf_finish_ray(P, R);
        else R.mode = WF_WAIT;
        return;
    }
    R.mo = R.hpos + R.hnormal * 0.0001f;
    R.qd = hemisphere_dir(R.hnormal, R.rng, R.hblock >= 0);
    R.t_stop = inf_f();
    wf_begin_query(P, R);
}

DDGI_HD void ov_init(WfRay& R, WfFeeler& F, v3 origin, v3 direction, uint32_t ray_index)
{
    wf_init(R, origin, direction, ray_index);
    F.mode = WF_F_IDLE;
    F.md = F.inv = F.p = V3(0, 0, 0);
    F.t = 0.0f;
    F.t_stop = inf_f();
    F.steps = 0;
}

// Scalar driver (tests/hostsim): both slots stepped for a single ray, `feeler_steps` steps of slot F for every
// `bounce_steps` steps of slot B - any interleaving of the two gives the same result.
DDGI_HD v3 overlap_trace_scalar(const FrameParams& P, v3 origin, v3 direction, uint32_t ray_index, uint32_t& lookups,
                                float* first_t = nullptr, int feeler_steps = 1, int bounce_steps = 1)
{
    if (first_t) *first_t = 0.0f;
    WfRay R;
    WfFeeler F;
    float stash[3] = {0, 0, 0};
    ov_init(R, F, origin, direction, ray_index);
    if (P.max_bounces <= 0) wf_finish_ray(P, R);
    if (R.mode == WF_QUERY) wf_begin_query(P, R);
    while (R.mode != WF_FETCH) {
        for (int i = 0; i < feeler_steps; i++) {
            if (F.mode == WF_F_MARCH) ov_step_feeler(P, R, F);
            else if (F.mode == WF_F_SLOW) ov_step_feeler_literal(P, R, F);
            if (F.mode == WF_F_HIT) {
                if (P.scene.color_mode != 0) ov_resolve_feeler<true>(P, R, F, stash, 1);
                else ov_resolve_feeler<false>(P, R, F, stash, 1);
            }
        }
        for (int i = 0; i < bounce_steps && R.mode != WF_FETCH; i++) {
            if (R.mode == WF_MARCH) wf_step(P, R);
            else if (R.mode == WF_MARCH_SLOW) wf_step_literal(P, R);
            wf_end_march(R);
            if (R.mode == WF_HIT && F.mode == WF_F_IDLE) {
                float* ft = R.bounce == 0 ? first_t : nullptr;
                if (P.scene.color_mode != 0) ov_resolve_bounce<true>(P, R, F, stash, 1, ft);
                else ov_resolve_bounce<false>(P, R, F, stash, 1, ft);
            }
        }
    }
    lookups += R.lookups;
    return wf_final_color(P, R);
}

}  // namespace ddgi
