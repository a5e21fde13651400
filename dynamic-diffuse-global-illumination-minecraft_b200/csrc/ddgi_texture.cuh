// ddgi_texture.cuh — the reference's procedural block colours (colour mode 1, "literal"),
// evaluated at a voxel hit: assets/shaders/intersection.glsl:872-1047 (getColorAt) with its
// helpers getUVs :828-863, dotsPattern :865-870, worleyNoise :465-499, fbm :421-435,
// fbm (1-D) :437-463, random1 :400.  Colour mode 0 reads the flat palette instead
// (README.md:266 "flat colors" variant, ddgi_scene.cuh: scene_albedo).
//
// Operation order follows the shader so that the result is bit-identical to the oracle and to
// the transpiled reference shaders (tests/golden/cave_3x3x3.npz); sin is the pinned fp64
// evaluation of ddgi_math.cuh.  Only executed once per bounce hit / primary hit.
#pragma once
#include "ddgi_scene.cuh"

namespace ddgi {

struct v2 {
    float x, y;
};
DDGI_HD v2 V2(float x, float y)
{
    v2 r;
    r.x = x;
    r.y = y;
    return r;
}
DDGI_HD float length2(float x, float y) { return sqrtf(x * x + y * y); }
DDGI_HD v3 mix3(v3 a, v3 b, float t) { return V3(gmix(a.x, b.x, t), gmix(a.y, b.y, t), gmix(a.z, b.z, t)); }

// intersection.glsl:400
DDGI_HD float hash3(v3 p)
{
    float d = (p.x * 127.1f + p.y * 311.7f) + p.z * 191.999f;
    return gfract(pin_sin(d) * 43758.5453f);
}
// 1-D value noise and its 8-octave sum, intersection.glsl:437-463 (octaves 0..7)
DDGI_HD float hash1(float i) { return gfract(pin_sin(203.311f * i)); }
DDGI_HD float interp_noise1D(float x)
{
    float ix = floorf(x);
    float fx = gfract(x);
    return gmix(hash1(ix), hash1(ix + 1.0f), fx);
}
DDGI_HD float fbm1D(float x)
{
    float total = 0.0f;
    float freq = 1.0f, amp = 1.0f;
    for (int i = 0; i < 8; i++) {
        total += interp_noise1D(x * freq) * amp;
        freq = freq * 2.0f;  // pow(2, i), exact
        amp = amp * 0.5f;    // pow(0.5, i), exact
    }
    return total;
}

// Worley noise on 5-unit cells, intersection.glsl:465-499.  The feature point of a cell is
// cell + fract(sin(vec2(dot(cell, a), dot(cell, b) * 43758.5453))) — the scale sits inside
// the second sine's argument.
DDGI_HD v2 worley_point(float cx, float cy)
{
    const float cell_size = 5.0f;
    float a = cx * 127.1f + cy * 311.7f;
    float b = (cx * 269.5f + cy * 183.3f) * 43758.5453f;
    return V2((cx + gfract(pin_sin(a))) * cell_size, (cy + gfract(pin_sin(b))) * cell_size);
}
DDGI_HD float worley(float px, float py)
{
    const float cell_size = 5.0f;
    float cx = floorf(px / cell_size), cy = floorf(py / cell_size);
    v2 q = worley_point(cx, cy);
    float shortest = length2(px - q.x, py - q.y);
    for (int i = -1; i <= 1; i++)
        for (int j = -1; j <= 1; j++) {
            v2 n = worley_point(cx + (float)i, cy + (float)j);
            float d = length2(px - n.x, py - n.y);
            if (d < shortest) shortest = d;
        }
    return shortest / cell_size;
}

// Face-local texture coordinates of a hit point, intersection.glsl:828-863.
DDGI_HD v2 face_uv(v3 p, v3 n)
{
    float fy = p.y - floorf(p.y);
    if (n.y == 0) {
        if (n.x == 0) return V2(gsign(n.z) > 0 ? ceilf(p.x) - p.x : p.x - floorf(p.x), fy);
        return V2(gsign(n.x) < 1 ? ceilf(p.z) - p.z : p.z - floorf(p.z), fy);
    }
    float fx = p.x - floorf(p.x);
    return V2(fx, gsign(n.y) < 0 ? ceilf(p.z) - p.z : p.z - floorf(p.z));
}

// Signed distance to a lattice of dots, intersection.glsl:865-870.
DDGI_HD float dots(v2 p, float radius, float cell)
{
    float c = 4.0f * radius * cell;
    float h = c / 2.0f;
    return length2(gmod(p.x + h, c) - h, gmod(p.y + h, c) - h) - radius;
}

// getColorAt(point, block_type, normal).rgb, intersection.glsl:872-1047.  Unknown types fall
// off the end of the shader function: (0,0,0) (oracle PIN 4).
DDGI_HD v3 block_color_literal(v3 p, int type, v3 n)
{
    switch (type) {
        case 1: {  // :875-906, the random1() result is overwritten by 0.3
            const float r = 0.3f;
            if (p.x < 0 && p.z > 0) return p.x < -16 ? V3(0.8f, 0.4f, 0.2f) : V3(0.1f, r, 0.2f);
            if (p.x < 0 && p.z < 0) return p.x < -16 ? V3(0.4f, 0.8f, 0.2f) : V3(0.99f, r, r);
            if (p.x > 0 && p.z < 0) return V3(0.1f, r, 0.5f);
            return V3(0.99f, r, r);
        }
        case 2: return V3(0.95f, 0, 0);
        case 3: return V3(0, 0.95f, 0);
        case 4: return V3(0, 0, 0.95f);
        case 5: return V3(0.95f, 0.95f, 0.95f);
        case 6:  // mushroom cap 1, :920-927
            return worley(p.x, p.z) < 0.35f ? V3(1, 0, 0.223f) : V3(1, 0.2f, 0);
        case 7: {  // mushroom cap 2, :928-936
            float w = worley(p.x + 5.0f, p.z + 5.0f);
            if (w < 0.25f) {
                v3 green = V3(0.8f, 1, 0);
                return V3(green.x - (w * (0.5f - green.x)), green.y - (w * (0.5f - green.y)), green.z - (w * (0.5f - green.z)));
            }
            return V3(1, 0, 0.011f);
        }
        case 8: {  // dotted cap, :937-953: uv rotated by mat2(0.707, -0.707, 0.707, 0.707) (column major)
            v2 g = face_uv(p, n);
            v2 uv = V2(0.707f * g.x + 0.707f * g.y, -0.707f * g.x + 0.707f * g.y);
            const float radius = 0.05f;
            float circle = (radius - dots(uv, radius, 1.8f)) * 100.0f;
            return mix3(V3(1, 0.313f, 0), V3(1, 0, 0.223f), gclamp(circle, 0.0f, 1.0f));
        }
        case 9: {  // stem, :954-963
            v2 uv = face_uv(p, n);
            float val = fbm2D(uv.x * 5.0f, p.z);
            val += 0.5f * fbm1D(p.x);
            return mix3(V3(0.3f, 0.1f, 0.3f), V3(0.9f, 0.9f, 0.9f), gclamp(val, 0.0f, 1.0f));
        }
        case 10: {  // cave wall, :964-1006: height bands blended with a per-cell two-tone pattern
            v3 band = V3(0.568f, 0.133f, 0.439f);
            if (p.y < -8) band = V3(0.349f, 0.133f, 0.427f);
            else if (p.y < -6) band = V3(0.568f, 0.133f, 0.439f);
            else if (p.y < -5) band = V3(0.639f, 0.176f, 0.725f);
            else if (p.y < 0) band = V3(0.274f, 0.188f, 0.772f);
            else if (p.y < 4) band = V3(0.341f, 0.270f, 0.768f);
            else if (p.y < 6) band = V3(0.368f, 0.203f, 0.415f);
            else if (p.y < 11) band = V3(0.470f, 0.270f, 0.729f);
            v2 uv = face_uv(p, n);
            float r = fbm2D(0.05f, (uv.y + p.y) * 0.3f);
            v3 wall = V3(0, 0.666f, 1);
            if (p.x < -1) {
                wall = V3(0.294f, 0.007f, 0.152f);
            } else if (p.x < 6 && p.x >= -1) {
                float gradient = p.x / 7.0f;
                float r2 = hash3(V3(ceilf(p.x), ceilf(p.y), ceilf(p.z)));
                wall = r2 < gradient ? V3(0, 0.666f, 1) : V3(0.294f, 0.007f, 0.152f);
            }
            return mix3(wall, band, r);
        }
        case 11: {  // cave ground, :1007-1021
            v3 dark = V3(0.294f, 0.007f, 0.152f);
            float r = hash3(V3(ceilf(p.x), ceilf(p.y), ceilf(p.z))) / 3.0f;
            v3 combined = mix3(dark, V3(0.901f, 0.992f, 0.427f), r);
            v2 uv = face_uv(p, n);
            r = fbm2D(uv.x * 2.0f, uv.y * 2.0f);
            return mix3(combined, dark, r / 2.0f);
        }
        case 12:
        case 13: {  // moss / mold, :1022-1046: radial gradient + noise of the normalised uv offset
            v2 uv = face_uv(p, n);
            v3 inner = type == 12 ? V3(0.356f, 1, 0.101f) : V3(0.803f, 1, 0.341f);
            float ax = uv.x - 0.5f, ay = uv.y - 0.5f;
            float inv = rcp_exact(sqrtf(ax * ax + ay * ay));  // normalize(vec2)
            float r = interp_noise2D(ax * inv, ay * inv);
            float dist = length2(uv.x - 0.5f, uv.y - 0.5f);
            return mix3(inner, V3(0.619f, 1, 0.278f), 2.0f * dist + r * 0.3f);
        }
        default: return V3(0, 0, 0);
    }
}

// Albedo of a voxel hit at continuous position p on the face with unit normal n.
DDGI_HD v3 scene_color(const SceneView& S, v3 p, int type, v3 n)
{
    if (S.color_mode == 0) return scene_albedo(S, type);
    return block_color_literal(p, type, n);
}

}  // namespace ddgi
