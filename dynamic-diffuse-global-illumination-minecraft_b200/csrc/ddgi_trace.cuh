// ddgi_trace.cuh — per-ray building blocks of the probe update and the pixel pass:
// RNG, hemisphere sampling, light-sphere test, voxel DDA march, nearest-hit query and
// the shadow-tested direct term.  Operation order follows the reference shaders so
// results are bit-identical to the CPU oracle (fp32, no contraction):
//   RNG                      assets/shaders/probe_pass.comp:45-71
//   hemisphere_dir           assets/shaders/probe_pass.comp:150-178
//   light_sphere             assets/shaders/intersection.glsl:78-121 (via :1264-1279)
//   march / nearest_hit      assets/shaders/intersection.glsl:1051-1100, :1244-1301
//   probe_direct_lighting    assets/shaders/probe_pass.comp:180-215
#pragma once
#include "ddgi_fastmath.cuh"
#include "ddgi_texture.cuh"

namespace ddgi {

struct Light {
    float intensity;
    float col[3];
    float pos[3];
};

constexpr int kMaxLights = 8;
constexpr int kMarchSteps = 125;  // intersection.glsl:1059

struct FrameParams {
    SceneView scene;
    int n_lights;
    Light lights[kMaxLights];
    // irradiance field (src/rvpt/rvpt.h:82-90)
    int probe_count[3];
    int side_length;
    int rx, ry;  // ray-tile shape; the reference has rx == ry == sqrt_rays_per_probe
    float field_origin[3];
    // render settings (src/rvpt/rvpt.h:70-80)
    int max_bounces;
    int screen_w, screen_h;
    // camera block (src/rvpt/camera.cpp:100-111) + host-evaluated 1/tan(hfov/2)
    float cam[20];
    float cam_w;
    // bounding sphere of the light positions (light_bounds below): one test rejects all
    // lights for a query that ends far from every one of them
    float lights_centre[3];
    float lights_radius;
    // pixel pass: integrator (compute_pass.comp:58-87) and probe markers (integrators.glsl:45-67)
    int render_mode;
    int visualize_probes;
    // 1: the reference's commented-out `weight *= chebyshevWeight` restored (intersection.glsl:1382)
    int weight_mode;
    // distance moments are stored and compared in units of distance_scale (1 = the reference's text)
    float distance_scale;
    // probe-texture layout: 0 = the reference's ray tile (one texel per ray, rx x ry), 1 = octahedral
    // tile of oct x oct texels (ddgi_octahedral.cuh); tile_w x tile_h is the tile either way
    int layout, oct;
    int tile_w, tile_h;
    // 1: result-preserving early-outs of the wavefront kernel (ddgi_wavefront.cuh: a shadow feeler's
    // march ends once it has left its light behind); texels are unchanged, the voxel-lookup counts
    // are no longer those of the reference algorithm (kernel variant 2, the default)
    int early_out;
};

struct Hit {
    float t;
    v3 pos;
    v3 normal;
    v3 base_color;
    v3 emissive;
    int type;  // 2 light sphere, 3 block, 0 none
};

DDGI_HD float inf_f()
{
#ifdef __CUDA_ARCH__
    return __int_as_float(0x7f800000);
#else
    return INFINITY;
#endif
}

// Host-side: a sphere that contains every light position, radius rounded up generously
// (1e-4 relative + 1e-4 absolute, far above the fp32 error of evaluating it).
inline void light_bounds(FrameParams& P)
{
    double c[3] = {0, 0, 0};
    for (int i = 0; i < P.n_lights; i++)
        for (int a = 0; a < 3; a++) c[a] += (double)P.lights[i].pos[a] / (P.n_lights > 0 ? P.n_lights : 1);
    double r = 0;
    for (int i = 0; i < P.n_lights; i++) {
        double d2 = 0;
        for (int a = 0; a < 3; a++) {
            double d = (double)P.lights[i].pos[a] - (double)(float)c[a];
            d2 += d * d;
        }
        if (!(sqrt(d2) <= r)) r = sqrt(d2);  // a NaN / Inf position makes the radius NaN / Inf: never rejects
    }
    for (int a = 0; a < 3; a++) P.lights_centre[a] = (float)c[a];
    P.lights_radius = (float)(r * 1.0001 + 1e-4);
}

DDGI_HD v3 lpos(const Light& l) { return V3(l.pos[0], l.pos[1], l.pos[2]); }
DDGI_HD v3 lcol(const Light& l) { return V3(l.col[0], l.col[1], l.col[2]); }

// The hysteresis blend the reference has commented out (probe_pass.comp:298-299):
// old_color = imageLoad(albedo, texel).rgb; color = mix(old_color, color, hysteresis).
DDGI_HD v3 blend_hysteresis(uint32_t old_texel, v3 color, float hysteresis)
{
    v3 old = unpack_rgb8(old_texel);
    return V3(gmix(old.x, color.x, hysteresis), gmix(old.y, color.y, hysteresis), gmix(old.z, color.z, hysteresis));
}

// ------------------------------------------------------------------ RNG
DDGI_HD uint32_t wang_hash(uint32_t seed)
{
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}
DDGI_HD float rng_next(uint32_t& st)
{
    st ^= (st << 13);
    st ^= (st >> 17);
    st ^= (st << 5);
    // float(uint) rounds to nearest; the scale by 2^-32 is exact (can yield 1.0)
    return (float)st / 4294967296.0f;
}

// `axis_normal`: the caller knows `normal` is an axis-aligned unit vector (or all-NaN).  Then
// cross(normal, other) and cross(normal, p1) are axis-aligned unit vectors too (signed zeros
// elsewhere) and normalize() returns them unchanged, so both calls are skipped.
DDGI_HD v3 hemisphere_dir(v3 normal, uint32_t& st, bool axis_normal = false)
{
    const float two_pi = 6.2831853071795864769252867665590057683943f;
    const float sqrt_third = 0.5773502691896257645091487805019574556476f;
    float up = sqrtf(rng_next(st));
    float over = sqrtf(1.0f - up * up);
    float around = rng_next(st) * two_pi;
    v3 other;
    if (fabsf(normal.x) < sqrt_third) other = V3(1, 0, 0);
    else if (fabsf(normal.y) < sqrt_third) other = V3(0, 1, 0);
    else other = V3(0, 0, 1);
    v3 p1 = cross(normal, other);
    if (!axis_normal) p1 = normalize(p1);
    v3 p2 = cross(normal, p1);
    if (!axis_normal) p2 = normalize(p2);
    float sn, cs;
    pin_sincos(around, &sn, &cs);
    return (normal * up + p1 * (cs * over)) + p2 * (sn * over);
}

// ------------------------------------------------------------------ light spheres
// Ray vs. the radius-0.1 sphere of light l: the unit-sphere quadratic on the
// ray scaled by 1/0.1.  Returns t (INF on a miss) and the unnormalised normal.
DDGI_HD float light_sphere(v3 origin, v3 direction, const Light& l, float maxt, v3* normal)
{
    v3 o = div_tenth(origin - lpos(l));  // x / 0.1f, exact (ddgi_fastmath.cuh)
    v3 d = div_tenth(direction);
    float A = dot(d, d);
    float B = -dot(d, o);
    float C = dot(o, o) - 1.0f;
    float D = B * B - A * C;
    D = D > 0 ? sqrtf(D) : inf_f();
    float t1 = (B - D) / A;
    float t2 = (B + D) / A;
    t1 = (0.0f < t1 && t1 < maxt) ? t1 : inf_f();
    t2 = (0.0f < t2 && t2 < maxt) ? t2 : inf_f();
    float t = gmin(t1, t2);
    *normal = o + d * t;
    return t;
}

// ------------------------------------------------------------------ voxel march
// Axis normal of the face a hit point lies on: sign of the dominant component of
// normalize(p - cell_centre), first axis wins ties.
DDGI_HD v3 face_normal(v3 p, v3 cell)
{
    v3 centre = V3(cell.x - 0.5f, cell.y - 0.5f, cell.z - 0.5f);
    v3 diff = normalize(p - centre);
    v3 n = V3(0, 0, 0);
    float mx = 0.0f;
    if (fabsf(diff.x) > mx) {
        mx = fabsf(diff.x);
        n = V3(gsign(diff.x), 0, 0);
    }
    if (fabsf(diff.y) > mx) {
        mx = fabsf(diff.y);
        n = V3(0, gsign(diff.y), 0);
    }
    if (fabsf(diff.z) > mx) {
        mx = fabsf(diff.z);
        n = V3(0, 0, gsign(diff.z));
    }
    return n;
}

// normalize(normalize(face_normal(p, cell))) — what grid_march (:1088) and intersect_scene
// (:1294) make of it: an axis-aligned unit vector is a fixed point of normalize (dot = 1,
// sqrt(1) = 1, 1/1 = 1, v * 1 = v) and the zero vector (no component of diff compares
// > 0: all zero or NaN) becomes 0 * (1/sqrt(0)) = 0 * Inf = NaN in every component.
DDGI_HD v3 face_normal_unit(v3 p, v3 cell)
{
    v3 n = face_normal(p, cell);
    if (n.x == 0.0f && n.y == 0.0f && n.z == 0.0f) {
        float q = 0.0f * inf_f();
        return V3(q, q, q);
    }
    return n;
}

// face_normal_unit for the common case, without the normalize(): when one component of
// p - cell_centre is the clear winner - in [2^-20, 4] and more than 2^-20 (relative) above the
// other two - scaling all three by the same finite positive factor 1/|diff| and rounding cannot
// change which one is largest (rounding is monotone and moves a value by at most 2^-24 relative)
// nor its sign, so face_normal's comparisons on the normalised vector pick that axis.  Anything
// else (near ties at cell edges, zero / NaN / huge positions) takes the literal path.
DDGI_HD v3 face_normal_axis(v3 p, v3 cell)
{
    v3 centre = V3(cell.x - 0.5f, cell.y - 0.5f, cell.z - 0.5f);
    v3 d = p - centre;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    const float margin = 1.00000095367431640625f;  // 1 + 2^-20
    if (ax >= 9.5367431640625e-07f && ax <= 4.0f && ax > ay * margin && ax > az * margin) return V3(d.x > 0.0f ? 1.0f : -1.0f, 0, 0);
    if (ay >= 9.5367431640625e-07f && ay <= 4.0f && ay > ax * margin && ay > az * margin) return V3(0, d.y > 0.0f ? 1.0f : -1.0f, 0);
    if (az >= 9.5367431640625e-07f && az <= 4.0f && az > ax * margin && az > ay * margin) return V3(0, 0, d.z > 0.0f ? 1.0f : -1.0f);
    return face_normal_unit(p, cell);
}

// One DDA advance: distance to the next lattice plane along each axis, the
// smallest plus the 1e-4 nudge, accumulate, re-evaluate the position.
DDGI_HD void march_advance(v3 origin, v3 dir, float& t, v3& p)
{
    v3 f = V3(gfract(p.x), gfract(p.y), gfract(p.z));
    float tx = gmax((-f.x) / dir.x, (1.0f - f.x) / dir.x);
    float ty = gmax((-f.y) / dir.y, (1.0f - f.y) / dir.y);
    float tz = gmax((-f.z) / dir.z, (1.0f - f.z) / dir.z);
    float step = gmin(gmin(tx, ty), tz) + 0.0001f;
    t += step;
    p = origin + dir * t;
}

// Marches at most 125 cells.  On a hit fills t, the (un-normalised) axis normal and
// the albedo.  `lookups` counts voxel queries (the reference's getBlockAt calls).
// Regular directions and origins (every component passes ddgi_fastmath.cuh's range checks —
// all but axis-parallel / degenerate rays) take the exact one-division step of the wavefront
// kernel: max((-f)/d, (1-f)/d) = (sel - f)/d with sel = 1 for d > 0 else 0, the division by the
// FMA-corrected reciprocal, floor / ceil by directed-rounding adds.  Bit-identical to the literal
// form below (tests/hostsim on the CPU, tests/selftest_div.cu exhaustively on the GPU).
// `want_surface` false (shadow feelers): only out.t is filled - no block type fetch, normal or colour.
DDGI_HD bool march(const SceneView& S, v3 origin, v3 direction, Hit& out, uint32_t& lookups, bool want_surface = true)
{
    v3 dir = normalize(direction);
    v3 p = origin;
    float t = 0.0f;
    const bool fast = regular_direction(dir.x, dir.y, dir.z) && regular_origin3(origin.x, origin.y, origin.z);
    bool hit = false;
    if (fast) {
        const v3 inv = V3(rcp_regular(dir.x), rcp_regular(dir.y), rcp_regular(dir.z));
        const v3 sel = V3(dir.x > 0 ? 1.0f : 0.0f, dir.y > 0 ? 1.0f : 0.0f, dir.z > 0 ? 1.0f : 0.0f);
        for (int i = 0; i < kMarchSteps; i++) {
            float tx = div_markstein(sel.x - (p.x - floor_small(p.x)), dir.x, inv.x);
            float ty = div_markstein(sel.y - (p.y - floor_small(p.y)), dir.y, inv.y);
            float tz = div_markstein(sel.z - (p.z - floor_small(p.z)), dir.z, inv.z);
            t += gmin(gmin(tx, ty), tz) + 0.0001f;
            p = origin + dir * t;
            lookups++;
            // |p| < 2^22: p + 1.5*2^23 rounded up IS ceil(p) + 1.5*2^23
            if (cell_solid(S, float_bits(add_round_up(p.x, kCellMagic)), float_bits(add_round_up(p.y, kCellMagic)),
                           float_bits(add_round_up(p.z, kCellMagic)))) {
                hit = true;
                break;
            }
        }
    } else {
        for (int i = 0; i < kMarchSteps; i++) {
            march_advance(origin, dir, t, p);
            lookups++;
            if (cell_solid(S, cell_bits(ceilf(p.x)), cell_bits(ceilf(p.y)), cell_bits(ceilf(p.z)))) {
                hit = true;
                break;
            }
        }
    }
    if (!hit) return false;
    out.t = t;
    if (!want_surface) return true;
    v3 cell = V3(ceilf(p.x), ceilf(p.y), ceilf(p.z));
    int type = scene_type_at(S, cell);
    v3 n = face_normal(p, cell);
    out.normal = normalize(n);
    out.base_color = scene_color(S, p, type, out.normal);
    out.emissive = V3(0, 0, 0);
    return true;
}

// Nearest of {light spheres, voxel march}.
DDGI_HD bool nearest_hit(const FrameParams& P, v3 origin, v3 direction, Hit& info, uint32_t& lookups)
{
    float closest = inf_f();
    info.t = closest;
    info.pos = V3(0, 0, 0);
    info.normal = V3(0, 0, 0);
    info.base_color = V3(0, 0, 0);
    info.emissive = V3(0, 0, 0);
    info.type = 0;
    for (int i = 0; i < P.n_lights; i++) {
        v3 n;
        float t = light_sphere(origin, direction, P.lights[i], closest, &n);
        if (t < closest) {
            info.t = t;
            info.normal = n;
            info.base_color = V3(0, 0, 0);
            info.emissive = lcol(P.lights[i]);
            info.type = 2;
        }
        closest = gmin(t, closest);
    }
    Hit m;
    if (march(P.scene, origin, direction, m, lookups)) {
        if (m.t < closest) {
            info.t = m.t;
            info.normal = m.normal;
            info.base_color = m.base_color;
            info.emissive = m.emissive;
            closest = m.t;
            info.type = 3;
        }
    }
    bool hit = closest < inf_f();
    info.normal = hit ? normalize(info.normal) : V3(0, 0, 0);
    info.pos = hit ? origin + direction * info.t : V3(0, 0, 0);
    info.pos = info.pos + info.normal * 0.001f;
    return hit;
}

// Light-sphere test of a query: nearest t over all lights, exactly as the loop of
// intersect_scene (intersection.glsl:1262-1279) evaluates it, except that it runs AFTER
// the march and skips lights that cannot beat the block hit at t_block:
// a root of |w + dir*t| = 0.1 (w = origin - light) has t*|dir| >= |w| - 0.1, and the
// reference's fp32 evaluation of it is within 1e-5 of that for |w| >= 0.101 (no
// cancellation: B^2/(A*C) <= 50), so with |w| > 1.01*t_block*|dir| + 0.101 every root it
// could report is > t_block: the block wins whatever the light's t is.  Skipping a light
// only widens the (0, closest) window of later ones by values > t_block, which lose
// to the block as well; with no block hit (t_block = INF) nothing is skipped.
// `dir_len` = sqrtf(dot(direction, direction)).
// `normal` (optional) receives the un-normalised sphere normal of the winning light.
DDGI_HD float light_test(const FrameParams& P, v3 origin, v3 direction, float dir_len, float t_block, int* which, v3* normal)
{
    float closest = inf_f();
    *which = -1;
    float reach2 = inf_f();
    unsigned near = 0xffu;  // bit i: light i may beat the block hit
    if (t_block < inf_f()) {
        float reach = (t_block * dir_len) * 1.01f + 0.101f;
        reach2 = reach * reach;
        // all lights at once: they lie within lights_radius of lights_centre, so
        // |w_i| >= |origin - centre| - radius for every i; 1.0001 covers the fp32 evaluation
        v3 wc = origin - V3(P.lights_centre[0], P.lights_centre[1], P.lights_centre[2]);
        float far = reach + P.lights_radius;
        if (dot(wc, wc) > (far * far) * 1.0001f) return closest;
        // light by light, without a branch per light (the loop count is the same for every lane)
        near = 0u;
        for (int i = 0; i < P.n_lights; i++) {
            v3 w = origin - lpos(P.lights[i]);
            near |= (dot(w, w) > reach2 ? 0u : 1u) << i;
        }
        if (near == 0u) return closest;
    }
    v3 d = div_tenth(direction);
    float A = dot(d, d);
    for (int i = 0; i < P.n_lights; i++) {
        if (!((near >> i) & 1u)) continue;
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

// nearest_hit with the voxel march FIRST and the light spheres tested afterwards by light_test (only those that
// could beat the block hit): the same `info` and the same voxel lookups as nearest_hit - the reference's march
// does not depend on the lights, the block wins iff m.t < closest either way, and light_test reports the
// reference loop's winner (first light on ties) with its un-normalised normal - for a fraction of the
// light arithmetic.  `want_surface` false (shadow feelers): info.type and info.t only.
DDGI_HD bool nearest_hit_marchfirst(const FrameParams& P, v3 origin, v3 direction, Hit& info, uint32_t& lookups, bool want_surface = true)
{
    info.pos = V3(0, 0, 0);
    info.normal = V3(0, 0, 0);
    info.base_color = V3(0, 0, 0);
    info.emissive = V3(0, 0, 0);
    info.type = 0;
    Hit m;
    const bool block = march(P.scene, origin, direction, m, lookups, want_surface);
    const float t_block = block ? m.t : inf_f();
    int which;
    v3 n = V3(0, 0, 0);
    float closest = light_test(P, origin, direction, sqrtf(dot(direction, direction)), t_block, &which, want_surface ? &n : nullptr);
    info.t = closest;
    if (t_block < closest) {
        info.t = t_block;
        closest = t_block;
        info.type = 3;
        if (want_surface) {
            info.normal = m.normal;
            info.base_color = m.base_color;
            info.emissive = m.emissive;
        }
    } else if (closest < inf_f()) {
        info.type = 2;
        info.normal = n;
        if (want_surface) info.emissive = lcol(P.lights[which]);
    }
    const bool hit = closest < inf_f();
    if (!want_surface) return hit;
    info.normal = hit ? normalize(info.normal) : V3(0, 0, 0);
    info.pos = hit ? origin + direction * info.t : V3(0, 0, 0);
    info.pos = info.pos + info.normal * 0.001f;
    return hit;
}

// Direct term at a probe-ray hit: one shadow feeler per light; a feeler blocked by a
// voxel ends the loop with the 0.2*albedo*lambert ambient term.
DDGI_HD v3 probe_direct_lighting(const FrameParams& P, const Hit& info, uint32_t& lookups)
{
    v3 direct = V3(0, 0, 0);
    int visible = 0;
    for (int i = 0; i < P.n_lights; i++) {
        const Light& l = P.lights[i];
        v3 to_light = normalize(lpos(l) - info.pos);
        Hit fh;
        if (nearest_hit(P, info.pos, to_light, fh, lookups)) {
            float lambert = gclamp(dot(normalize(info.normal), to_light), 0.0f, 1.0f);
            if (fh.type == 2) {
                float dist = length(lpos(l) - info.pos);
                direct = direct + ((lcol(l) * lambert) * l.intensity) / dist;
            } else {
                return (info.base_color * 0.2f) * lambert;
            }
            visible++;
        }
    }
    if (visible != 0) return (info.base_color * direct) / (float)visible;
    return V3(0, 0, 0);
}

// Whole probe-ray path: up to max_bounces hits, each adding its direct term.
// `first_t` (optional) receives t of the first query (INF on a miss, 0 without bounces).
DDGI_HD v3 trace_probe_ray(const FrameParams& P, v3 origin, v3 direction, uint32_t ray_index,
                           uint32_t& lookups, float* first_t = nullptr)
{
    uint32_t rng = wang_hash(ray_index);
    v3 color = V3(0, 0, 0);
    v3 o = origin, d = direction;
    Hit hit;
    if (first_t) *first_t = 0.0f;
    for (int b = 0; b < P.max_bounces; b++) {
        bool isect = nearest_hit(P, o, d, hit, lookups);
        if (first_t && b == 0) *first_t = hit.t;
        if (!isect) break;
        color = color + probe_direct_lighting(P, hit, lookups);
        o = hit.pos + hit.normal * 0.0001f;
        d = hemisphere_dir(hit.normal, rng);
    }
    return color / (float)P.max_bounces;
}

// Probe lattice position and tile addressing (src/rvpt/rvpt.cpp:1190-1205,
// assets/shaders/probe_pass.comp:139-145).
DDGI_HD v3 probe_origin(const FrameParams& P, int p)
{
    int X = P.probe_count[0], Y = P.probe_count[1], Z = P.probe_count[2];
    int py = p / (X * Z);
    int rest = p - py * X * Z;
    int pz = rest / X;
    int px = rest - pz * X;
    v3 o = V3((float)(px - (X - 1) / 2), (float)(py - (Y - 1) / 2), (float)(pz - (Z - 1) / 2));
    o = o * (float)P.side_length;
    return o + V3(P.field_origin[0], P.field_origin[1], P.field_origin[2]);
}

}  // namespace ddgi
