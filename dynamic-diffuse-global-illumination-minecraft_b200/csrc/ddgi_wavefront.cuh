// ddgi_wavefront.cuh — the probe-ray path of ddgi_trace.cuh re-expressed as a per-lane
// state machine, so a warp can keep all lanes inside the DDA step loop and regroup the
// expensive, divergent "a march just ended" work.
//
// A probe ray is a chain of nearest-hit queries: per bounce one query along the ray and
// then one shadow feeler per light (assets/shaders/probe_pass.comp:283-295, :180-215).
// Each query is a light-sphere pre-test plus a voxel march of up to 125 steps
// (assets/shaders/intersection.glsl:1244-1301, :1051-1100).  Here every lane is either
// MARCHING (wf_step: one DDA advance + voxel test) or PENDING (wf_transition: resolve
// the query, shade, start the next query); the kernel runs wf_step while enough lanes
// march and batches the transitions (ddgi_kernels.cu: probe_update_wavefront).
//
// The arithmetic is the reference's, operation for operation, so results are
// bit-identical to ddgi_trace.cuh and to the oracle.  What differs is only how each
// correctly-rounded result is obtained:
//   * max((-f)/d, (1-f)/d) needs one division: for d > 0 the first quotient is <= 0 <=
//     the second, for d < 0 the other way round (f in [0,1]); zero / NaN / tiny
//     components take the literal two-division form (WfRay::slow).
//   * that division and x/0.1f use the FMA-corrected reciprocal of ddgi_fastmath.cuh.
//   * floor(p) is ceil(p)-1 unless p is an integer; ceil(p) is needed anyway for the
//     voxel id, so a step costs three FRND instead of six (integers take the literal form).
//   * the voxel test reads one bit of the 4x4x4-brick occupancy word (16 MiB for 512^3
//     voxels, L1/L2 resident); the block type is fetched only on a hit.
//   * a light sphere whose discriminant is not positive yields t = INF in the reference
//     (intersection.glsl:100-113), so the two root divisions are skipped for it.
#pragma once
#include "ddgi_fastmath.cuh"
#include "ddgi_trace.cuh"

namespace ddgi {

// WF_HIT / WF_MISS: the march ended on a solid cell / after 125 empty cells
enum : int { WF_MARCH = 0, WF_HIT = 1, WF_MISS = 2, WF_DONE = 3, WF_IDLE = 4 };

struct WfRay {
    // current march
    v3 mo;    // query origin
    v3 md;    // normalize(query direction)
    v3 inv;   // 1 / md (valid when !slow)
    v3 p;     // position after the last advance
    v3 c;     // ceil(p)
    float t;
    int steps;
    int mode;
    int slow;  // a direction component is zero, NaN or tiny: literal step arithmetic
    // current query
    v3 qd;  // query direction as given (positions are origin + qd * t)
    float light_t;
    int light_i;  // nearest light sphere, -1 none
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

// Light-sphere pre-test of a query: nearest t over all lights and which light, exactly
// as the loop of intersect_scene (intersection.glsl:1262-1279) evaluates it.  `normal`
// (optional) receives the un-normalised sphere normal of the winning light.
DDGI_HD float light_pretest(const FrameParams& P, v3 origin, v3 direction, int* which, v3* normal)
{
    float closest = inf_f();
    *which = -1;
    v3 d = div_tenth(direction);
    float A = dot(d, d);
    for (int i = 0; i < P.n_lights; i++) {
        v3 o = div_tenth(origin - lpos(P.lights[i]));
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

// Starts a nearest-hit query: light spheres first (they do not depend on the march),
// then arm the march.
DDGI_HD void wf_begin_query(const FrameParams& P, WfRay& R, v3 origin, v3 direction)
{
    R.mo = origin;
    R.qd = direction;
    R.md = normalize(direction);
    // fast-step preconditions (ddgi_fastmath.cuh): regular direction components, and no
    // origin component in (0, 2^-70) so that a position is either 0 or >= 2^-98 in magnitude
    R.slow = (int)!(regular_component(R.md.x) && regular_component(R.md.y) && regular_component(R.md.z)) |
             (int)(tiny_nonzero(origin.x) || tiny_nonzero(origin.y) || tiny_nonzero(origin.z));
    R.inv = R.slow ? V3(0, 0, 0) : V3(1.0f / R.md.x, 1.0f / R.md.y, 1.0f / R.md.z);
    R.p = origin;
    R.c = V3(ceilf(origin.x), ceilf(origin.y), ceilf(origin.z));
    R.t = 0.0f;
    R.steps = 0;
    R.light_t = light_pretest(P, origin, direction, &R.light_i, nullptr);
    R.mode = WF_MARCH;
}

DDGI_HD void wf_finish_ray(const FrameParams& P, WfRay& R)
{
    R.color = R.color / (float)P.max_bounces;
    R.mode = WF_DONE;
}

// Path state of a fresh ray (the caller starts its first query with wf_begin_query).
DDGI_HD void wf_init(WfRay& R, uint32_t ray_index)
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
}

// One DDA advance and voxel test (the body of the reference's 125-iteration loop).
DDGI_HD void wf_step(const FrameParams& P, WfRay& R)
{
    // ---- advance: t += min_a(max((-f_a)/d_a, (1-f_a)/d_a)) + 1e-4 ----
    if (!R.slow) {
        v3 fl = V3(R.c.x - 1.0f, R.c.y - 1.0f, R.c.z - 1.0f);  // floor(p) unless p is an integer
        if (DDGI_UNLIKELY(R.p.x == R.c.x || R.p.y == R.c.y || R.p.z == R.c.z))
            fl = V3(floorf(R.p.x), floorf(R.p.y), floorf(R.p.z));
        v3 f = R.p - fl;
        // numerator of the larger quotient: 1-f for d > 0, -f for d < 0 (a zero numerator
        // may come out as +0 where the reference has -0: min(..)+1e-4 is the same)
        float nx = (R.md.x > 0 ? 1.0f : 0.0f) - f.x;
        float ny = (R.md.y > 0 ? 1.0f : 0.0f) - f.y;
        float nz = (R.md.z > 0 ? 1.0f : 0.0f) - f.z;
        float tx = div_markstein(nx, R.md.x, R.inv.x);
        float ty = div_markstein(ny, R.md.y, R.inv.y);
        float tz = div_markstein(nz, R.md.z, R.inv.z);
        float step = gmin(gmin(tx, ty), tz) + 0.0001f;
        R.t += step;
        R.p = R.mo + R.md * R.t;
    } else {
        march_advance(R.mo, R.md, R.t, R.p);
    }
    R.c = V3(ceilf(R.p.x), ceilf(R.p.y), ceilf(R.p.z));
    R.steps++;
    // ---- voxel test: one bit of the brick occupancy word ----
    if (cell_solid(P.scene, cell_bits(R.c.x), cell_bits(R.c.y), cell_bits(R.c.z))) R.mode = WF_HIT;
    else if (R.steps >= kMarchSteps) R.mode = WF_MISS;
}

// A march ended: resolve the query (nearest of light sphere / block) and advance the
// bounce / feeler bookkeeping.  Returns true when the ray is finished (R.color final),
// otherwise the next query's origin / direction.
DDGI_HD bool wf_resolve(const FrameParams& P, WfRay& R, v3& o, v3& d)
{
    float closest = R.light_t;
    int type = R.light_i >= 0 ? 2 : 0;
    bool block_hit = R.mode == WF_HIT && R.t < closest;
    R.lookups += (uint32_t)R.steps;
    if (block_hit) {
        closest = R.t;
        type = 3;
    }
    bool hit = closest < inf_f();

    bool end_bounce = false;
    v3 result = V3(0, 0, 0);
    if (R.phase == 0) {
        // the bounce ray itself
        if (!hit) {
            wf_finish_ray(P, R);
            return true;
        }
        v3 n;
        if (block_hit) {
            n = normalize(face_normal(R.p, R.c));
            R.hblock = scene_type_at(P.scene, R.c);
        } else {
            // a light sphere is the nearest hit (rare): redo the pre-test for its normal
            int which;
            light_pretest(P, R.mo, R.qd, &which, &n);
            R.hblock = -1;
        }
        v3 normal = normalize(n);
        R.hpos = (R.mo + R.qd * closest) + normal * 0.001f;
        R.hnormal = normal;
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
                v3 base = R.hblock >= 0 ? scene_albedo(P.scene, R.hblock) : V3(0, 0, 0);
                end_bounce = true;
                result = (base * 0.2f) * lambert;
            }
        }
        if (!end_bounce) {
            R.phase++;
            if (R.phase > P.n_lights) {
                end_bounce = true;
                if (R.visible != 0) {
                    v3 base = R.hblock >= 0 ? scene_albedo(P.scene, R.hblock) : V3(0, 0, 0);
                    result = (base * R.direct) / (float)R.visible;
                }
            }
        }
    }
    if (end_bounce) {
        // probe_pass.comp:286-292: accumulate, pick the next bounce direction
        R.color = R.color + result;
        o = R.hpos + R.hnormal * 0.0001f;
        d = hemisphere_dir(R.hnormal, R.rng);
        R.bounce++;
        if (R.bounce >= P.max_bounces) {
            wf_finish_ray(P, R);
            return true;
        }
        R.phase = 0;
    } else {
        o = R.hpos;
        d = normalize(lpos(P.lights[R.phase - 1]) - R.hpos);
    }
    return false;
}

// Scalar driver (tests/hostsim): the state machine stepped for a single ray.
DDGI_HD v3 wavefront_trace_scalar(const FrameParams& P, v3 origin, v3 direction, uint32_t ray_index,
                                  uint32_t& lookups)
{
    WfRay R;
    wf_init(R, ray_index);
    if (P.max_bounces <= 0) {
        wf_finish_ray(P, R);
        return R.color;
    }
    wf_begin_query(P, R, origin, direction);
    for (;;) {
        if (R.mode == WF_MARCH) {
            wf_step(P, R);
            continue;
        }
        v3 o, d;
        if (wf_resolve(P, R, o, d)) break;
        wf_begin_query(P, R, o, d);
    }
    lookups += R.lookups;
    return R.color;
}

}  // namespace ddgi
