// ddgi_octahedral.cuh — the textbook probe layout the north star names and the reference only
// gestures at: an oct x oct octahedral tile per probe whose texels hold the cosine-weighted mean
// of ALL the probe's rays (irradiance in the albedo plane, first-hit distance moments in the
// distance plane), sampled bilinearly at octEncode(direction).
//
// The reference ships the mapping — assets/shaders/octahedral.glsl:16-35 (octEncode / octDecode,
// G3D; signNotZero from the g3dmath.glsl it includes but does not ship) — and never includes it:
// its tile has one texel per ray (probe_pass.comp:269-271).  So this layout has NO reference
// output to pin against ("parity unpinned" for this mode); the contract is the
// oracle's restatement of the operation order below, and the engine must match it bit for bit.
//
// Operation order (what makes the warp-shuffle reduction reproducible):
//   * lane l of a warp accumulates the rays i = l, l+32, l+64, ... in that order;
//   * the 32 partial sums are combined by the xor butterfly, offsets 16, 8, 4, 2, 1:
//     acc[l] = acc[l] + acc[l ^ off]  (fp addition commutes, so every lane ends with the same sum);
//   * texel = sum(w c) / sum(w), w = max(0, dot(texel direction, ray direction)); 0 if sum(w) == 0.
#pragma once
#include "ddgi_math.cuh"

namespace ddgi {

DDGI_HD float sign_not_zero(float x) { return x >= 0.0f ? 1.0f : -1.0f; }

// octEncode, octahedral.glsl:16-23 (v a unit vector) -> [-1, 1]^2
DDGI_HD void oct_encode(v3 v, float* ox, float* oy)
{
    float l1 = (fabsf(v.x) + fabsf(v.y)) + fabsf(v.z);
    float inv = 1.0f / l1;
    float rx = v.x * inv, ry = v.y * inv;
    if (v.z < 0.0f) {
        float tx = (1.0f - fabsf(ry)) * sign_not_zero(rx);
        float ty = (1.0f - fabsf(rx)) * sign_not_zero(ry);
        rx = tx;
        ry = ty;
    }
    *ox = rx;
    *oy = ry;
}

// octDecode, octahedral.glsl:28-35
DDGI_HD v3 oct_decode(float ox, float oy)
{
    v3 v = V3(ox, oy, (1.0f - fabsf(ox)) - fabsf(oy));
    if (v.z < 0.0f) {
        float tx = (1.0f - fabsf(v.y)) * sign_not_zero(v.x);
        float ty = (1.0f - fabsf(v.x)) * sign_not_zero(v.y);
        v.x = tx;
        v.y = ty;
    }
    return normalize(v);
}

// Direction of the centre of texel (u, v) of an oct x oct tile.
DDGI_HD v3 oct_texel_dir(int u, int v, int oct)
{
    float ox = (((float)u + 0.5f) / (float)oct) * 2.0f - 1.0f;
    float oy = (((float)v + 0.5f) / (float)oct) * 2.0f - 1.0f;
    return oct_decode(ox, oy);
}

struct OctAcc {
    float r, g, b, d, d2, w;
};
DDGI_HD OctAcc oct_add(OctAcc a, OctAcc b)
{
    OctAcc c;
    c.r = a.r + b.r;
    c.g = a.g + b.g;
    c.b = a.b + b.b;
    c.d = a.d + b.d;
    c.d2 = a.d2 + b.d2;
    c.w = a.w + b.w;
    return c;
}

// A ray's distance sample: first-hit t in units of `scale`, capped at 1 (a miss has t = INF).
DDGI_HD float oct_ray_distance(float first_t, float scale) { return gmin(first_t / scale, 1.0f); }

// Lane `lane`'s partial sums for the texel looking along dir_t: rays lane, lane + 32, ...
// dirs: n normalised ray directions (xyz); radiance: n x (r, g, b, first_t).
DDGI_HD OctAcc oct_lane_partial(v3 dir_t, const float* dirs, const float* radiance, int n, int lane, float scale)
{
    OctAcc a = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    for (int i = lane; i < n; i += 32) {
        v3 di = V3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        float w = gmax(0.0f, dot(dir_t, di));
        float d = oct_ray_distance(radiance[4 * i + 3], scale);
        a.r += w * radiance[4 * i];
        a.g += w * radiance[4 * i + 1];
        a.b += w * radiance[4 * i + 2];
        a.d += w * d;
        a.d2 += w * (d * d);
        a.w += w;
    }
    return a;
}

// sum -> the two texels; `blend` mixes with the old texels by the reference's hysteresis rule
// (probe_pass.comp:298-299: mix(old, new, hysteresis)).
DDGI_HD void oct_finalize(OctAcc t, int blend, float hysteresis, uint32_t old_albedo, uint32_t old_distance,
                          uint32_t* albedo, uint32_t* distance)
{
    v3 e = V3(0, 0, 0);
    float d = 0.0f, d2 = 0.0f;
    if (t.w > 0.0f) {
        e = V3(t.r / t.w, t.g / t.w, t.b / t.w);
        d = t.d / t.w;
        d2 = t.d2 / t.w;
    }
    if (blend) {
        v3 oa = unpack_rgb8(old_albedo), od = unpack_rgb8(old_distance);
        e = V3(gmix(oa.x, e.x, hysteresis), gmix(oa.y, e.y, hysteresis), gmix(oa.z, e.z, hysteresis));
        d = gmix(od.x, d, hysteresis);
        d2 = gmix(od.y, d2, hysteresis);
    }
    *albedo = pack_rgba8(e.x, e.y, e.z, 1.0f);
    *distance = pack_rgba8(d, d2, 0.0f, 0.0f);
}

// Bilinear fetch of an oct x oct tile with origin (cx, cy) at octEncode(dir), clamped to the tile.
DDGI_HD v3 oct_sample_tile(const uint32_t* tex, int W, int cx, int cy, int oct, v3 dir)
{
    float ox, oy;
    oct_encode(normalize(dir), &ox, &oy);
    float fx = (ox * 0.5f + 0.5f) * (float)oct - 0.5f;
    float fy = (oy * 0.5f + 0.5f) * (float)oct - 0.5f;
    float x0f = floorf(fx), y0f = floorf(fy);
    float ax = fx - x0f, ay = fy - y0f;
    int x0 = f2i(x0f), y0 = f2i(y0f);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = x0 < 0 ? 0 : (x0 > oct - 1 ? oct - 1 : x0);
    x1 = x1 < 0 ? 0 : (x1 > oct - 1 ? oct - 1 : x1);
    y0 = y0 < 0 ? 0 : (y0 > oct - 1 ? oct - 1 : y0);
    y1 = y1 < 0 ? 0 : (y1 > oct - 1 ? oct - 1 : y1);
    v3 c00 = unpack_rgb8(tex[(size_t)(cy + y0) * W + cx + x0]);
    v3 c10 = unpack_rgb8(tex[(size_t)(cy + y0) * W + cx + x1]);
    v3 c01 = unpack_rgb8(tex[(size_t)(cy + y1) * W + cx + x0]);
    v3 c11 = unpack_rgb8(tex[(size_t)(cy + y1) * W + cx + x1]);
    v3 top = c00 * (1.0f - ax) + c10 * ax;
    v3 bot = c01 * (1.0f - ax) + c11 * ax;
    return top * (1.0f - ay) + bot * ay;
}

}  // namespace ddgi
