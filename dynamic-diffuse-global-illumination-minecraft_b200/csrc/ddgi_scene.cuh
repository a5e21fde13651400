// ddgi_scene.cuh — stored voxel field (what replaces the reference's compiled-in
// procedural getBlockAt, assets/shaders/intersection.glsl:699-826) and the built-in
// scene bakers that evaluate that procedural function once per voxel.
//
// HBM layout (DESIGN.md "Data layout"):
//   occ   : one uint32 per brick of 1x4x8 voxels (kBrickL*), bit occ_shift(x, y, z) & 31 set = solid,
//           x/y/z = voxel id relative to `borg` (a multiple of 8 per axis).  Brick index
//           ((bz*nby)+by)*nbx+bx, x fastest: a 32-byte sector holds 8x4x8 voxels.
//           512^3 voxels -> 16 MiB, L1/L2 resident.
//   types : one uint8 block type per voxel, linear x-fastest; read only on a hit.
//   palette: 256 x rgb fp32 albedo by block type (flat-colour variant, README.md:266).
// Voxel id c (an integer-valued float, c = ceil(position)) covers (c-1, c] per axis;
// grid cell g = c - vorg.  Anything outside the grid, or a NaN id, is empty.
#pragma once
#include "ddgi_math.cuh"

namespace ddgi {

struct SceneView {
    const uint32_t* occ;
    const uint8_t* types;
    const float* palette;
    int vorg[3];
    int vdim[3];
    int borg[3];  // voxel id of brick (0,0,0)'s first cell: vorg rounded down to a multiple of kBrickAlign
    int kneg[3];  // -(kCellBias + borg): biased cell coordinate -> borg-relative cell coordinate in one add
    int nb[3];    // bricks per axis
    float lo[3];  // (float)vorg
    float hi[3];  // (float)(vorg + vdim - 1)
    int color_mode;  // 0: flat palette by block type, 1: the reference's procedural colours (ddgi_texture.cuh)
};

DDGI_HD int float_bits(float x)
{
#ifdef __CUDA_ARCH__
    return __float_as_int(x);
#else
    union {
        float f;
        int i;
    } u;
    u.f = x;
    return u.i;
#endif
}

// Integer-valued float c -> kCellBias + (int)c, exact for |c| < 2^22 (one FADD instead
// of a conversion).  Anything else (NaN, Inf, |c| >= 2^22) maps to a value whose brick
// falls outside every grid the engine accepts (|voxel id| < 2^22), i.e. reads as empty.
constexpr int kCellBias = 0x4B400000;
DDGI_HD int cell_bits(float c) { return float_bits(c + 12582912.0f); }

// word >> (s mod 32): SHF.R.W on the device, so callers need not mask the shift count
DDGI_HD uint32_t shr_wrap(uint32_t w, int s)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(w, 0u, (unsigned)s);
#else
    return w >> ((unsigned)s & 31u);
#endif
}

// Brick shape: a brick is 2^kBrickLx x 2^kBrickLy x 2^kBrickLz cells = 32 cells = one uint32.  The
// words are stored x fastest, so a 32-byte sector holds 8 bricks along x.  DDGI_BRICK 148 (default):
// bricks of 1 x 4 x 8 cells, a sector covers 8 x 4 x 8 cells - close to a cube, so a march crosses
// into a new sector about equally rarely along every axis (round 1's 4 x 4 x 2 bricks made a sector
// 32 x 4 x 2 cells: every second step along z left it).  Measured in profiles/r2_ab.md.
#ifndef DDGI_BRICK
#define DDGI_BRICK 148
#endif
#if DDGI_BRICK == 442
constexpr int kBrickLx = 2, kBrickLy = 2, kBrickLz = 1;
#else
constexpr int kBrickLx = 0, kBrickLy = 2, kBrickLz = 3;
#endif
constexpr int kBrickAlign = 8;  // borg is a multiple of this on every axis (>= the largest brick side)

// Bit of the cell with borg-relative coordinates (gx, gy, gz) inside its brick word, as a shift
// count that is taken modulo 32.  1x4x8 bricks: bits 0-1 = gy & 3, bits 2-4 = gz & 7.  4x4x2 bricks:
// bits 0-1 = gx & 3, bits 2-3 = gy & 3, bit 4 = (gz ^ (gy >> 2)) & 1 - the z layer of a brick is
// skewed by the brick row's parity so that the count is three shift-adds with no masking of gy, gz.
// Writers (build_occupancy_kernel, tests/hostsim) and readers (cell_solid / cell_fetch) share this
// one definition.
DDGI_HD int occ_shift(int gx, int gy, int gz)
{
#if DDGI_BRICK == 442
    return (gx & 3) + (gy << 2) + (gz << 4);
#else
    (void)gx;
    return (gy & 3) + (gz << 2);
#endif
}

// Where a cell's bit sits in its word.  DDGI_OCC_MSB (default): counted from the top, bit 31 - (shift mod 32), so
// that the test is a wrapping left shift and a sign test (2 instructions of the march step instead of shift,
// mask, compare).  Writers use occ_mask, readers occ_test: one definition.
#ifndef DDGI_OCC_MSB
#define DDGI_OCC_MSB 1
#endif
DDGI_HD uint32_t occ_mask(int shift)
{
#if DDGI_OCC_MSB
    return 0x80000000u >> ((unsigned)shift & 31u);
#else
    return 1u << ((unsigned)shift & 31u);
#endif
}
DDGI_HD bool occ_test(uint32_t word, int shift)
{
#if DDGI_OCC_MSB
#ifdef __CUDA_ARCH__
    return (int)__funnelshift_l(0u, word, (unsigned)shift) < 0;  // SHF.L.W: the count is taken mod 32
#else
    return (int)(word << ((unsigned)shift & 31u)) < 0;
#endif
#else
    return (bool)(shr_wrap(word, shift) & 1u);
#endif
}

// The occupancy word `idx` when `inside`, else 0 (bricks outside the grid are empty): one
// predicated load, no branch, no divergence.
DDGI_HD uint32_t occ_word(const uint32_t* occ, unsigned idx, bool inside)
{
#ifdef __CUDA_ARCH__
    uint32_t w = 0u;
    if (inside) w = __ldg(occ + idx);
    return w;
#else
    return inside ? occ[idx] : 0u;
#endif
}

// Occupancy test of the cell with biased integer coordinates (kx,ky,kz) = cell_bits(c)
// per axis; bricks outside the grid read as empty.
DDGI_HD bool cell_solid(const SceneView& S, int kx, int ky, int kz)
{
    int gx = kx + S.kneg[0];
    int gy = ky + S.kneg[1];
    int gz = kz + S.kneg[2];
    unsigned bx = (unsigned)gx >> kBrickLx, by = (unsigned)gy >> kBrickLy, bz = (unsigned)gz >> kBrickLz;  // (negative: huge)
    bool inside = (bx < (unsigned)S.nb[0]) & (by < (unsigned)S.nb[1]) & (bz < (unsigned)S.nb[2]);
    uint32_t word = occ_word(S.occ, (bz * (unsigned)S.nb[1] + by) * (unsigned)S.nb[0] + bx, inside);
    return occ_test(word, occ_shift(gx, gy, gz));
}

// Block type of an occupied cell (only called after its occupancy bit tested set).
DDGI_HD int scene_type_at(const SceneView& S, v3 c)
{
    int gx = (int)c.x - S.vorg[0], gy = (int)c.y - S.vorg[1], gz = (int)c.z - S.vorg[2];
    return S.types[((size_t)gz * S.vdim[1] + gy) * S.vdim[0] + gx];
}

// Returns the block type (0 = empty) of voxel id c.
DDGI_HD int scene_lookup(const SceneView& S, v3 c)
{
    if (!cell_solid(S, cell_bits(c.x), cell_bits(c.y), cell_bits(c.z))) return 0;
    return scene_type_at(S, c);
}

DDGI_HD v3 scene_albedo(const SceneView& S, int type)
{
    const float* c = S.palette + 3 * (type & 255);
    return V3(c[0], c[1], c[2]);
}

// ---------------------------------------------------------------------------------
// Procedural block function, evaluated by the bakers only.
// Noise: intersection.glsl:400-435.  Mushrooms: :538-697.  getBlockAt: :699-826.
// ---------------------------------------------------------------------------------
DDGI_HD float noise2D(float px, float py)
{
    float d = px * 127.1f + py * 311.7f;
    return gfract(pin_sin(d) * 43758.5453f);
}
DDGI_HD float interp_noise2D(float x, float y)
{
    int ix = f2i(floorf(x));
    float fx = gfract(x);
    int iy = f2i(floorf(y));
    float fy = gfract(y);
    float v1 = noise2D((float)ix, (float)iy);
    float v2 = noise2D((float)(ix + 1), (float)iy);
    float v3_ = noise2D((float)ix, (float)(iy + 1));
    float v4 = noise2D((float)(ix + 1), (float)(iy + 1));
    float i1 = gmix(v1, v2, fx);
    float i2 = gmix(v3_, v4, fx);
    return gmix(i1, i2, fy);
}
DDGI_HD float fbm2D(float x, float y)
{
    float total = 0.0f;
    float freq = 1.0f, amp = 1.0f;
    for (int i = 1; i <= 8; i++) {
        freq = freq * 2.0f;  // pow(2, i), exact
        amp = amp * 0.5f;    // pow(0.5, i), exact
        total += interp_noise2D(x * freq, y * freq) * amp;
    }
    return total;
}

DDGI_HD float sd_round_box(v3 p, v3 b, float r)
{
    v3 q = V3(fabsf(p.x) - b.x, fabsf(p.y) - b.y, fabsf(p.z) - b.z);
    v3 qm = V3(gmax(q.x, 0.0f), gmax(q.y, 0.0f), gmax(q.z, 0.0f));
    return length(qm) + gmin(gmax(q.x, gmax(q.y, q.z)), 0.0f) - r;
}

// The four mushroom shapes: a rounded-box cap split into three layers by sign(y)
// and a stem made of up to three vertical segments.
struct CapSpec {
    float bx, by, bz, r;
    int above, level, below;  // block types for y>0, y==0, y<0 inside the cap
};
DDGI_HD int cap_layers(v3 p, CapSpec c)
{
    if (sd_round_box(p, V3(c.bx, c.by, c.bz), c.r) <= 0) {
        if (p.y > 0) return c.above;
        if (p.y == 0) return c.level;
        if (p.y < 0) return c.below;
    }
    return -1;
}
DDGI_HD int mushroom_tiny(v3 p)
{
    if (sd_round_box(p, V3(1.0f, 0.5f, 1.0f), 0.0f) <= 0) return 7;
    if (p.x == 0 && p.z == 0 && p.y < 0) return 9;
    return 0;
}
DDGI_HD int mushroom_small(v3 p)
{
    CapSpec c = {1.0f, 0.5f, 1.0f, 1.0f, 8, 7, 6};
    int k = cap_layers(p, c);
    if (k >= 0) return k;
    if (p.x == 0 && p.z == 0 && p.y < 0) return 9;
    return 0;
}
DDGI_HD int mushroom_medium(v3 p)
{
    CapSpec c = {2.0f, 0.5f, 2.0f, 1.0f, 6, 7, 8};
    int k = cap_layers(p, c);
    if (k >= 0) return k;
    if (p.x == 0 && p.z == 0 && p.y < 0 && p.y > -7) return 9;
    if (p.x == 1 && p.z == 0 && p.y < -5 && p.y > -12) return 9;
    if (p.x == 2 && p.z == 0 && p.y < -10) return 9;
    return 0;
}
DDGI_HD int mushroom_large(v3 p, int dir)
{
    CapSpec c = {3.0f, 0.5f, 3.0f, 1.5f, 6, 8, 7};
    int k = cap_layers(p, c);
    if (k >= 0) return k;
    if (p.x == 0 && p.z == 0 && p.y < 0 && p.y > -9) return 9;
    if (p.x == 0 && p.z == (float)dir && p.y < -7 && p.y > -18) return 9;
    if (p.x == 0 && p.z == (float)(2 * dir) && p.y < -16) return 9;
    return 0;
}

// Placement table of the cave's mushrooms by quadrant (intersection.glsl:630-697).
DDGI_HD int cave_mushrooms(v3 c)
{
    if (c.x < 0 && c.z > 0) {
        if (c.x < -16) {
            if (c.z > 20) return mushroom_tiny(c - V3(-19, -12, 22));
            if (c.z < 4) return mushroom_tiny(c - V3(-18, -12, 2));
            int k = mushroom_large(c - V3(-22, 3, 8), -1);
            if (k != 0) return k;
            return mushroom_medium(c - V3(-27, -4, 16));
        }
        if (c.z > 10 && c.x > -6) return mushroom_tiny(c - V3(-4, -14, 12));
        if (c.z < 14) return mushroom_medium(c - V3(-4, -1, 6));
        return mushroom_small(c - V3(-10, -8, 18));
    }
    if (c.x < 0 && c.z < 0) {
        if (c.x < -16) {
            if (c.x < -28) {
                if (c.z < -16) return mushroom_tiny(c - V3(-32, -14, -20));
                return mushroom_tiny(c - V3(-30, -12, -12));
            }
            if (c.z > -10) return mushroom_small(c - V3(-25, -7, -4));
            return mushroom_medium(c - V3(-20, -3, -20));
        }
        if (c.x < -12 && c.z > -12) return mushroom_tiny(c - V3(-14, -15, -10));
        if (c.z > -10 && c.x > -4) return mushroom_tiny(c - V3(-2, -12, -2));
        if (c.z < -10) return mushroom_small(c - V3(-5, -9, -14));
        return mushroom_large(c - V3(-8, 8, -6), 1);
    }
    if (c.x > 0 && c.z < 0) {
        if (c.z > -5) return mushroom_tiny(c - V3(6, -14, -3));
        if (c.z < -14) {
            if (c.x > 18) return mushroom_tiny(c - V3(20, -7, -16));
            return mushroom_large(c - V3(14, 10, -20), -1);
        }
        return mushroom_medium(c - V3(6, -6, -10));
    }
    return 0;
}

DDGI_HD int block_cave(v3 c)
{
    if (c.y > 17.0f) return 0;
    if (c.y < -15) {
        if (c.y < -18) {
            float r = fbm2D(c.x * 0.3f, c.z * 0.3f);
            if (f2i(floorf(r * 2.0f)) == 0) return 12;
        }
        float r = fbm2D(c.x * 0.058f, c.z * 0.058f);
        int d = f2i(floorf(r * 5.0f));
        if ((float)(-21 + d) >= c.y) return c.y == -18 ? 13 : 11;
    }
    bool outside = length(c) - 20.0f > 0.0f && length(c + V3(16, 8, -10)) - 20.0f > 0.0f &&
                   length(c + V3(-13, -1, 19)) - 18.0f > 0.0f && length(c + V3(20, 15, 15)) - 21.0f > 0.0f;
    if (outside) return 10;
    return cave_mushrooms(c);
}
DDGI_HD int block_cornell(v3 c)
{
    bool in_yz = fabsf(c.y) < 10 && fabsf(c.z - 15) < 10;
    if (c.x == -10 && in_yz) return 2;
    if (c.x == 10 && in_yz) return 3;
    if (fabsf(c.y) == 10 && fabsf(c.x) < 10 && fabsf(c.z - 15) < 10) return 5;
    if (c.z == 25 && fabsf(c.x) < 10 && fabsf(c.y) < 10) return 5;
    if (fabsf(c.x + 3) < 3 && fabsf(c.y + 7) < 3 && fabsf(c.z - 13) < 3) return 5;
    if (fabsf(c.x - 4) < 3 && fabsf(c.y + 4) < 6 && fabsf(c.z - 16) < 3) return 5;
    return 0;
}
DDGI_HD int block_house(v3 c)
{
    if (c.y == -5) return 1;
    if (fabsf(c.x) == 25 && fabsf(c.y) < 5 && fabsf(c.z) < 15) return 2;
    if (c.y == 5 && fabsf(c.x) < 25 && fabsf(c.z) < 15) return 5;
    if (c.z == -15 && fabsf(c.x) < 25 && fabsf(c.y) < 5) return 3;
    if (c.z == 15) {
        if (fabsf(c.x - 10) < 2 && fabsf(c.y + 1) < 4) return 0;
        if (fabsf(c.x) < 25 && fabsf(c.y) < 5) return 3;
    }
    return 0;
}
DDGI_HD int block_procedural(v3 c, int scene)
{
    if (scene == 0) return block_cave(c);
    if (scene == 1) return block_cornell(c);
    if (scene == 2) return block_house(c);
    return 0;
}

}  // namespace ddgi
