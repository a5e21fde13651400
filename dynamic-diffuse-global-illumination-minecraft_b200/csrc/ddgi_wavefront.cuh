// ddgi_wavefront.cuh — the probe-ray path of ddgi_trace.cuh re-expressed as a per-lane
// state machine, so a warp can keep all lanes inside the DDA step loop and regroup the
// expensive, divergent "a march just ended" work.
//
// A probe ray is a chain of nearest-hit queries: per bounce one query along the ray and
// then one shadow feeler per light (assets/shaders/probe_pass.comp:283-295, :180-215).
// Each query is a light-sphere pre-test plus a voxel march of up to 125 steps
// (assets/shaders/intersection.glsl:1244-1301, :1051-1100).  March lengths are
// geometrically distributed, so in the reference's nested-loop form a warp idles on its
// longest march ~40 times per ray.  Here every lane is either MARCHING (wf_step: one
// DDA advance + voxel test) or PENDING (wf_transition: resolve the query, shade, start
// the next query); the kernel runs wf_step while enough lanes march and batches the
// transitions (ddgi_kernels.cu: probe_update_wavefront).
//
// Every floating-point operation, and its order, is the same as in ddgi_trace.cuh, so
// both kernel variants (and the oracle) produce identical bits.
#pragma once
#include "ddgi_trace.cuh"

namespace ddgi {

enum : int { WF_MARCH = 0, WF_PENDING = 1, WF_DONE = 2 };

struct WfRay {
    // current march
    v3 mo;       // query origin
    v3 md;       // normalize(query direction)
    v3 p;        // position after the last advance
    v3 cell;     // ceil(p)
    float t;
    int steps;
    int mode;
    int block;   // block type the march ended on (0: 125 steps without a hit)
    // current query
    v3 qd;       // query direction as given (positions are origin + qd * t)
    float light_t;
    int light_i; // nearest light sphere so far, -1 none
    v3 light_n;
    // path
    int bounce;
    int phase;   // 0: the bounce ray itself; i >= 1: shadow feeler to light i-1
    v3 hpos, hnormal, hbase;
    v3 direct;
    int visible;
    v3 color;
    uint32_t rng;
    uint32_t lookups;
};

// Starts a nearest-hit query: light spheres first (they do not depend on the march),
// then arm the march.
DDGI_HD void wf_begin_query(const FrameParams& P, WfRay& R, v3 origin, v3 direction)
{
    R.mo = origin;
    R.qd = direction;
    R.md = normalize(direction);
    R.p = origin;
    R.t = 0.0f;
    R.steps = 0;
    R.block = 0;
    float closest = inf_f();
    R.light_i = -1;
    R.light_n = V3(0, 0, 0);
    for (int i = 0; i < P.n_lights; i++) {
        v3 n;
        float t = light_sphere(origin, direction, P.lights[i], closest, &n);
        if (t < closest) {
            R.light_i = i;
            R.light_n = n;
        }
        closest = gmin(t, closest);
    }
    R.light_t = closest;
    R.mode = WF_MARCH;
}

DDGI_HD void wf_finish_ray(const FrameParams& P, WfRay& R)
{
    R.color = R.color / (float)P.max_bounces;
    R.mode = WF_DONE;
}

DDGI_HD void wf_init(const FrameParams& P, WfRay& R, v3 origin, v3 direction, uint32_t ray_index)
{
    R.rng = wang_hash(ray_index);
    R.color = V3(0, 0, 0);
    R.direct = V3(0, 0, 0);
    R.visible = 0;
    R.bounce = 0;
    R.phase = 0;
    R.lookups = 0;
    R.hpos = R.hnormal = R.hbase = V3(0, 0, 0);
    R.cell = V3(0, 0, 0);
    if (P.max_bounces <= 0) {
        wf_finish_ray(P, R);
        return;
    }
    wf_begin_query(P, R, origin, direction);
}

// One DDA advance and voxel test (the body of the reference's 125-iteration loop).
DDGI_HD void wf_step(const FrameParams& P, WfRay& R)
{
    march_advance(R.mo, R.md, R.t, R.p);
    R.cell = V3(ceilf(R.p.x), ceilf(R.p.y), ceilf(R.p.z));
    R.lookups++;
    R.steps++;
    int type = scene_lookup(P.scene, R.cell);
    if (type > 0) {
        R.block = type;
        R.mode = WF_PENDING;
    } else if (R.steps >= kMarchSteps) {
        R.block = 0;
        R.mode = WF_PENDING;
    }
}

// A march ended: resolve the query (nearest of light sphere / block), advance the
// bounce / feeler bookkeeping and arm the next query (single wf_begin_query site).
DDGI_HD void wf_transition(const FrameParams& P, WfRay& R)
{
    float closest = R.light_t;
    int type = R.light_i >= 0 ? 2 : 0;
    float t = closest;
    v3 n = R.light_n;
    v3 base = V3(0, 0, 0);
    if (R.block > 0 && R.t < closest) {
        t = R.t;
        n = normalize(face_normal(R.p, R.cell));
        base = scene_albedo(P.scene, R.block);
        closest = R.t;
        type = 3;
    }
    bool hit = closest < inf_f();
    v3 normal = hit ? normalize(n) : V3(0, 0, 0);
    v3 pos = hit ? R.mo + R.qd * t : V3(0, 0, 0);
    pos = pos + normal * 0.001f;

    bool end_bounce = false;
    v3 result = V3(0, 0, 0);
    if (R.phase == 0) {
        // the bounce ray itself
        if (!hit) {
            wf_finish_ray(P, R);
            return;
        }
        R.hpos = pos;
        R.hnormal = normal;
        R.hbase = base;
        R.direct = V3(0, 0, 0);
        R.visible = 0;
        if (P.n_lights == 0) end_bounce = true;
        else R.phase = 1;
    } else {
        // shadow feeler to light phase-1 (probe_pass.comp:186-207)
        const Light& l = P.lights[R.phase - 1];
        if (hit) {
            float lambert = gclamp(dot(normalize(R.hnormal), R.qd), 0.0f, 1.0f);
            if (type == 2) {
                float dist = length(lpos(l) - R.hpos);
                R.direct = R.direct + ((lcol(l) * lambert) * l.intensity) / dist;
                R.visible++;
            } else {
                end_bounce = true;
                result = (R.hbase * 0.2f) * lambert;
            }
        }
        if (!end_bounce) {
            R.phase++;
            if (R.phase > P.n_lights) {
                end_bounce = true;
                if (R.visible != 0) result = (R.hbase * R.direct) / (float)R.visible;
            }
        }
    }
    v3 o, d;
    if (end_bounce) {
        // probe_pass.comp:286-292: accumulate, pick the next bounce direction
        R.color = R.color + result;
        o = R.hpos + R.hnormal * 0.0001f;
        d = hemisphere_dir(R.hnormal, R.rng);
        R.bounce++;
        if (R.bounce >= P.max_bounces) {
            wf_finish_ray(P, R);
            return;
        }
        R.phase = 0;
    } else {
        o = R.hpos;
        d = normalize(lpos(P.lights[R.phase - 1]) - R.hpos);
    }
    wf_begin_query(P, R, o, d);
}

// Scalar driver (tests/hostsim): the state machine stepped for a single ray.
DDGI_HD v3 wavefront_trace_scalar(const FrameParams& P, v3 origin, v3 direction, uint32_t ray_index,
                                  uint32_t& lookups)
{
    WfRay R;
    wf_init(P, R, origin, direction, ray_index);
    while (R.mode != WF_DONE) {
        if (R.mode == WF_MARCH) wf_step(P, R);
        else wf_transition(P, R);
    }
    lookups += R.lookups;
    return R.color;
}

}  // namespace ddgi
