// ddgi_shade.cuh — per-pixel side of the path: probe-tile lookup, the 8-probe-cage
// weighted sample, the integrators and the pinhole camera.
//   tile_origin / sample_tile   assets/shaders/intersection.glsl:1152-1240
//   cage_irradiance             assets/shaders/intersection.glsl:1306-1409
//   probe_marker_t              assets/shaders/intersection.glsl:314-392, :1102-1129
//   shade_ddgi                  assets/shaders/integrators.glsl:27-106
//   shade_direct .. shade_normal  assets/shaders/integrators.glsl:110-271 (the debug views)
//   shade_pixel                 assets/shaders/compute_pass.comp:58-87 (eval_integrator)
//   pinhole_ray                 assets/shaders/camera.glsl:29-51
#pragma once
#include "ddgi_octahedral.cuh"
#include "ddgi_trace.cuh"

namespace ddgi {

// Tile origin of probe `p` in the packed texture, or (-1,-1).
DDGI_HD void tile_origin(const FrameParams& P, int p, int* ox, int* oy)
{
    int x_dim = P.probe_count[0] * P.probe_count[2];
    *ox = -1;
    *oy = -1;
    if (p >= x_dim * P.probe_count[1]) return;
    if (p < 0 || x_dim < 0) return;
    int col = f2i(gmod((float)p, (float)x_dim));
    int row = p / x_dim;
    if (row >= P.probe_count[1]) return;
    *ox = col * P.tile_w;
    *oy = row * P.tile_h;
}

// Octahedral layout: bilinear fetch of probe p's tile of `tex` at octEncode(dir).
DDGI_HD v3 sample_tile_oct(const FrameParams& P, const uint32_t* tex, int W, int p, v3 dir)
{
    int cx, cy;
    tile_origin(P, p, &cx, &cy);
    if (cx == -1 && cy == -1) return V3(1, 0, 1);
    return oct_sample_tile(tex, W, cx, cy, P.oct, dir);
}

// Direction -> texel of a probe tile (intersection.glsl:1196-1207): depends on the direction alone, so the
// cage sample evaluates it ONCE for its eight probes (the reference recomputes it, acos included, per probe).
DDGI_HD void tile_texel(const FrameParams& P, v3 dir, int* relx, int* rely)
{
    const float pi = 3.1415926535897932384626433832795f;
    v3 d = normalize(dir);
    *relx = f2i(((-1.0f * (d.z - 1.0f)) / 2.0f) * (float)P.rx);
    if (*relx == P.rx) *relx = 0;
    float sq = sqrtf(1.0f - (d.z * d.z));
    *rely = f2i((pin_acos(d.x / sq) / (2.0f * pi)) * (float)P.ry);
}

// The mean of the centre texel plus the in-tile part of the 5x5 window around it (centre counted
// twice).  `taps` is the image the window is read from: the albedo texture, or the distance texture for
// texture_to_sample = 1 — the centre texel comes from the ALBEDO texture either way (intersection.glsl:1213).
DDGI_HD v3 gather_tile(const FrameParams& P, const uint32_t* tex, const uint32_t* taps, int W, int p, int relx, int rely)
{
    int cx, cy;
    tile_origin(P, p, &cx, &cy);
    if (cx == -1 && cy == -1) return V3(1, 0, 1);
    int sx = cx + relx, sy = cy + rely;
    v3 sum = unpack_rgb8(tex[(size_t)sy * W + sx]);
    // the in-tile part of the window [-2, 2]^2 around (relx, rely): the same taps in the same order as the
    // reference's two loops with their four range tests per tap (intersection.glsl:1218-1236)
    const int x0 = relx < 2 ? -relx : -2, x1 = relx + 2 > P.rx - 1 ? P.rx - 1 - relx : 2;
    const int y0 = rely < 2 ? -rely : -2, y1 = rely + 2 > P.ry - 1 ? P.ry - 1 - rely : 2;
    int count = 0;
    for (int x = x0 < -2 ? -2 : x0; x <= x1; x++) {
        const uint32_t* col = taps + (size_t)(sy + (y0 < -2 ? -2 : y0)) * W + (sx + x);
        for (int y = y0 < -2 ? -2 : y0; y <= y1; y++) {
            count++;
            sum = sum + unpack_rgb8(*col);
            col += W;
        }
    }
    return sum / (float)count;
}
DDGI_HD v3 sample_tile(const FrameParams& P, const uint32_t* tex, const uint32_t* taps, int W, int p, v3 dir)
{
    int relx, rely;
    tile_texel(P, dir, &relx, &rely);
    return gather_tile(P, tex, taps, W, p, relx, rely);
}

// pow(x, 3) of intersection.glsl:1379: the cube in fp64 rounded once (oracle PIN 11)
DDGI_HD float pin_pow3(float x) { return (float)(((double)x * (double)x) * (double)x); }

// kExt = false compiles the reference as shipped only (no Chebyshev term, no probe markers): the
// kernel the default frame runs carries none of the optional code.
template <bool kExt>
DDGI_HD v3 cage_irradiance(const FrameParams& P, const uint32_t* tex, const uint32_t* dist_tex, int W, const Hit& info)
{
    v3 pos = info.pos;
    v3 N = normalize(info.normal);
    v3 fo = V3(P.field_origin[0], P.field_origin[1], P.field_origin[2]);
    float side = (float)P.side_length;
    v3 q = (pos - fo) / side;
    int base[3] = {f2i(floorf(q.x)), f2i(floorf(q.y)), f2i(floorf(q.z))};
    // the reference bounds every axis with the x probe count (int(vec3) = .x)
    int lo = f2i(-floorf((float)P.probe_count[0] / 2.0f));
    int hi = f2i(floorf((float)P.probe_count[0] / 2.0f) - 1.0f);
    for (int i = 0; i < 3; i++)
        if (base[i] < lo || base[i] > hi) return V3(1, 0, 1);

    v3 base_world = V3((float)(base[0] * P.side_length), (float)(base[1] * P.side_length),
                       (float)(base[2] * P.side_length)) + fo;
    v3 a = (pos - base_world) / side;
    v3 alpha = V3(gclamp(a.x, 0.0f, 1.0f), gclamp(a.y, 0.0f, 1.0f), gclamp(a.z, 0.0f, 1.0f));
    int X = P.probe_count[0], Y = P.probe_count[1], Z = P.probe_count[2];
    v3 irradiance = V3(0, 0, 0);
    float sum_w = 0.0f;
    int relx = 0, rely = 0;
    if (!(kExt && P.layout == 1)) tile_texel(P, N, &relx, &rely);  // the same texel of every probe's tile
    for (int i = 0; i < 8; i++) {
        int ox = (i >> 2) & 1, oy = (i >> 1) & 1, oz = i & 1;
        int sx = base[0] + ox + X / 2, sy = base[1] + oy + Y / 2, sz = base[2] + oz + Z / 2;
        int p = sy * X * Z + sz * X + sx;
        if (p < 0 || p >= X * Y * Z) return V3(1, 0, 1);
        v3 off = V3((float)ox, (float)oy, (float)oz);
        v3 tri = V3(gmix(1.0f - alpha.x, alpha.x, off.x), gmix(1.0f - alpha.y, alpha.y, off.y),
                    gmix(1.0f - alpha.z, alpha.z, off.z));
        v3 probe_pos = base_world + off * side;
        v3 dir = normalize(probe_pos - pos);
        float bf = gmax(0.0001f, (dot(dir, N) + 1.0f) * 0.5f);
        float w = bf * bf + 0.2f;
        // the reference computes a Chebyshev visibility term and discards it
        // (intersection.glsl:1367-1383); weight_mode 1 restores `weight *= chebyshevWeight`
        if (kExt && P.weight_mode == 1) {
            float probe_dist = length(pos - probe_pos) / P.distance_scale;
            v3 to_pos = V3(-dir.x, -dir.y, -dir.z);
            v3 mms = P.layout == 1 ? sample_tile_oct(P, dist_tex, W, p, to_pos) : sample_tile(P, tex, dist_tex, W, p, to_pos);
            float mean = mms.x;
            float variance = fabsf(mean * mean - mms.y);
            float over = gmax(probe_dist - mean, 0.0f);
            float cheb = variance / (variance + over * over);
            cheb = gmax(pin_pow3(cheb), 0.0f);
            if (!(probe_dist <= mean)) w *= cheb;
        }
        w = gmax(0.000001f, w);
        const float crush = 0.2f;
        if (w < crush) w *= w * w * (1.f / (crush * crush));
        w *= tri.x * tri.y * tri.z;
        v3 e = (kExt && P.layout == 1) ? sample_tile_oct(P, tex, W, p, N) : gather_tile(P, tex, tex, W, p, relx, rely);
        irradiance = irradiance + e * w;
        sum_w += w;
    }
    return irradiance / sum_w;
}

// Probe markers ("Visualize Probes"): sphere-traces the repeated radius-0.2 sphere at every
// lattice point (sceneSDF -> opRepLim, intersection.glsl:332-346) while t < 100
// (implicit_surface, :366-392).  Only the hit parameter reaches the image
// (integrators.glsl:49-64), so the marker normal (estimateNormal) is not evaluated.
// Returns t, or INF on a miss.
DDGI_HD float probe_marker_sdf(const FrameParams& P, v3 point)
{
    v3 p = point - V3(P.field_origin[0], P.field_origin[1], P.field_origin[2]);
    float c = (float)P.side_length;
    v3 l = V3((float)(P.probe_count[0] / 2), (float)(P.probe_count[1] / 2), (float)(P.probe_count[2] / 2));
    v3 r = V3(roundf(p.x / c), roundf(p.y / c), roundf(p.z / c));
    v3 cl = V3(gmin(gmax(r.x, -l.x), l.x), gmin(gmax(r.y, -l.y), l.y), gmin(gmax(r.z, -l.z), l.z));
    v3 q = p - cl * c;
    return length(q) - 0.2f;
}
DDGI_HD float probe_marker_t(const FrameParams& P, v3 origin, v3 direction)
{
    v3 dir = normalize(direction);
    float t = 0.f;
    while (t < 100.0f) {
        float dist = probe_marker_sdf(P, origin + dir * t);
        if (dist < 0.001f) return t;
        t += dist;
    }
    return inf_f();
}
DDGI_HD bool probe_marker_in_front(const FrameParams& P, v3 origin, v3 direction, const Hit& info)
{
    if (!P.visualize_probes) return false;
    float t = probe_marker_t(P, origin, direction);
    return t < inf_f() && t < info.t;
}

// The shadow-tested direct term of the pixel integrators (integrators.glsl:78-97, :131-148): no
// early return and no ambient term, unlike the probe pass's (probe_pass.comp:180-215).
DDGI_HD int pixel_direct_term(const FrameParams& P, const Hit& info, v3* direct, uint32_t& lookups)
{
    int visible = 0;
    *direct = V3(0, 0, 0);
    for (int i = 0; i < P.n_lights; i++) {
        const Light& l = P.lights[i];
        v3 to_light = normalize(lpos(l) - info.pos);
        Hit fh;
        if (nearest_hit_marchfirst(P, info.pos, to_light, fh, lookups, false) && fh.type == 2) {
            float lambert = gclamp(dot(normalize(info.normal), to_light), 0.0f, 1.0f);
            float dist = length(lpos(l) - info.pos);
            *direct = *direct + ((lcol(l) * lambert) * l.intensity) / dist;
            visible++;
        }
    }
    return visible;
}

template <bool kExt>
DDGI_HD v3 shade_ddgi(const FrameParams& P, const uint32_t* tex, const uint32_t* dist_tex, int W, v3 origin, v3 direction,
                      uint32_t& lookups)
{
    Hit info;
    bool hit = nearest_hit_marchfirst(P, origin, direction, info, lookups);
    if (kExt && probe_marker_in_front(P, origin, direction, info)) return V3(0, 1, 1);
    if (!hit) return V3(0.898f, 0.968f, 1.0f);
    if (info.type == 2) return info.emissive;
    v3 indirect = cage_irradiance<kExt>(P, tex, dist_tex, W, info);
    v3 direct;
    int visible = pixel_direct_term(P, info, &direct, lookups);
    v3 half_base = info.base_color * 0.5f;
    if (visible != 0) return half_base * (direct / (float)visible) + half_base * indirect;
    return (indirect * 0.5f) * info.base_color;
}

// integrator_direct, integrators.glsl:110-156
DDGI_HD v3 shade_direct(const FrameParams& P, v3 origin, v3 direction, uint32_t& lookups)
{
    Hit info;
    if (!nearest_hit_marchfirst(P, origin, direction, info, lookups)) return V3(0, 0, 0);
    v3 direct;
    int visible = pixel_direct_term(P, info, &direct, lookups);
    if (visible != 0) return (info.base_color * 0.5f) * (direct / (float)visible);
    return V3(0, 0, 0);
}

// integrator_indirect, integrators.glsl:160-207
DDGI_HD v3 shade_indirect(const FrameParams& P, const uint32_t* tex, const uint32_t* dist_tex, int W, v3 origin, v3 direction,
                          uint32_t& lookups)
{
    Hit info;
    bool hit = nearest_hit_marchfirst(P, origin, direction, info, lookups);
    if (probe_marker_in_front(P, origin, direction, info)) return V3(0, 1, 1);
    if (!hit) return V3(0, 0, 0);
    return cage_irradiance<true>(P, tex, dist_tex, W, info) * 0.5f;
}

// integrator_color / _depth / _normal, integrators.glsl:211-271
DDGI_HD v3 shade_color(const FrameParams& P, v3 origin, v3 direction, uint32_t& lookups)
{
    Hit info;
    if (!nearest_hit_marchfirst(P, origin, direction, info, lookups)) return V3(0, 0, 0);
    return info.base_color;
}
DDGI_HD v3 shade_depth(const FrameParams& P, v3 origin, v3 direction, uint32_t& lookups)
{
    Hit info;
    nearest_hit_marchfirst(P, origin, direction, info, lookups);
    float inv_dist = 1.0f / (length(direction) * info.t);
    return V3(inv_dist, inv_dist, inv_dist);
}
DDGI_HD v3 shade_normal(const FrameParams& P, v3 origin, v3 direction, uint32_t& lookups)
{
    Hit info;
    float isect = nearest_hit_marchfirst(P, origin, direction, info, lookups) ? 1.0f : 0.0f;
    float h = 0.5f * isect;
    return info.normal * 0.5f + V3(h, h, h);
}

// eval_integrator, compute_pass.comp:58-87: modes outside 1..5 take the DDGI branch.
// kExt = false: render_mode 0 as shipped (the caller has checked that nothing optional is on).
template <bool kExt>
DDGI_HD v3 shade_pixel(const FrameParams& P, const uint32_t* tex, const uint32_t* dist_tex, int W, v3 origin, v3 direction,
                       uint32_t& lookups)
{
    if (!kExt) return shade_ddgi<false>(P, tex, dist_tex, W, origin, direction, lookups);
    switch (P.render_mode) {
        case 1: return shade_direct(P, origin, direction, lookups);
        case 2: return shade_indirect(P, tex, dist_tex, W, origin, direction, lookups);
        case 3: return shade_color(P, origin, direction, lookups);
        case 4: return shade_normal(P, origin, direction, lookups);
        case 5: return shade_depth(P, origin, direction, lookups);
        default: return shade_ddgi<true>(P, tex, dist_tex, W, origin, direction, lookups);
    }
}

DDGI_HD void pinhole_ray(const FrameParams& P, float x, float y, v3* origin, v3* direction)
{
    const float* m = P.cam;
    float u = m[16] * (2.0f * x - 1.0f);
    float v = 2.0f * y - 1.0f;
    float w = P.cam_w;
    *origin = V3(m[12], m[13], m[14]);
    v3 d;
    d.x = ((m[0] * u + m[4] * v) + m[8] * w) + m[12] * 0.0f;
    d.y = ((m[1] * u + m[5] * v) + m[9] * w) + m[13] * 0.0f;
    d.z = ((m[2] * u + m[6] * v) + m[10] * w) + m[14] * 0.0f;
    *direction = normalize(d);
}

}  // namespace ddgi
