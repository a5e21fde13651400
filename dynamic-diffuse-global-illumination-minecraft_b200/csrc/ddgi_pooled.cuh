// ddgi_pooled.cuh — EXPERIMENTAL kernel variant 2: the wavefront state machine of
// ddgi_wavefront.cuh with the rays of a thread block held in a shared-memory POOL instead of in
// the lanes' registers, so that a warp can gather any 32 rays that are in the same state.
//
// Why (profiles/r1_policy_model.md): variant 1 is issue-bound at ~20 of 32 active lanes per
// instruction because a warp can only group the 32 rays it owns; regrouping over the 4 warps of
// a block has an upper bound of 0.70x the instructions.  The price is moving ray state between
// shared memory and registers on every issue, which is why the record is packed into nine
// float4 (LDS.128 / STS.128) and the march — the frequent, cheap state — touches only four of
// them on load and three on store.
//
// The per-ray arithmetic is the SAME wf_* functions as variant 1 (bit-identical results by
// construction); this header only defines the pool record.  tests/hostsim round-trips every
// ray through pack/unpack after every state execution with all other fields poisoned, so a
// field missing from a record shows up in the CPU parity tests.
#pragma once
#include "ddgi_wavefront.cuh"

namespace ddgi {

struct PoolVec {
    float x, y, z, w;
};
constexpr int kPoolVecs = 9;  // 144 bytes per ray

DDGI_HD float pool_bits_f(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union {
        uint32_t u;
        float f;
    } c;
    c.u = u;
    return c.f;
#endif
}
DDGI_HD uint32_t pool_f_bits(float f) { return (uint32_t)float_bits(f); }

// v0: mo, t        v1: md, steps | hit_mode << 8     v2: inv, lookups     v3: p, rng
// v4: qd, bounce | phase << 8 | visible << 16        v5: hpos, hblock     v6: hnormal, k
// v7: direct, first_t                                v8: color, -
DDGI_HD PoolVec pool_vec(v3 a, float w)
{
    PoolVec v;
    v.x = a.x;
    v.y = a.y;
    v.z = a.z;
    v.w = w;
    return v;
}

// What a march (WF_MARCH / WF_MARCH_SLOW) reads: v0..v3.
DDGI_HD void pool_unpack_march(const PoolVec* r, WfRay& R)
{
    R.mo = V3(r[0].x, r[0].y, r[0].z);
    R.t = r[0].w;
    R.md = V3(r[1].x, r[1].y, r[1].z);
    uint32_t a = pool_f_bits(r[1].w);
    R.steps = (int)(a & 255u);
    R.hit_mode = (int)(a >> 8);
    R.inv = V3(r[2].x, r[2].y, r[2].z);
    R.lookups = pool_f_bits(r[2].w);
    R.p = V3(r[3].x, r[3].y, r[3].z);
    R.rng = pool_f_bits(r[3].w);
    R.sel = V3(R.md.x > 0 ? 1.0f : 0.0f, R.md.y > 0 ? 1.0f : 0.0f, R.md.z > 0 ? 1.0f : 0.0f);  // as wf_begin_query sets it
}
// What a march writes back: t (v0), steps (v1), p (v3).
DDGI_HD void pool_pack_march(const WfRay& R, PoolVec* r)
{
    r[0] = pool_vec(R.mo, R.t);
    r[1] = pool_vec(R.md, pool_bits_f((uint32_t)R.steps | ((uint32_t)R.hit_mode << 8)));
    r[3] = pool_vec(R.p, pool_bits_f(R.rng));
}

DDGI_HD void pool_unpack(const PoolVec* r, WfRay& R, uint32_t& k, float& first_t)
{
    pool_unpack_march(r, R);
    R.qd = V3(r[4].x, r[4].y, r[4].z);
    uint32_t b = pool_f_bits(r[4].w);
    R.bounce = (int)(b & 255u);
    R.phase = (int)((b >> 8) & 255u);
    R.visible = (int)((b >> 16) & 255u);
    R.hpos = V3(r[5].x, r[5].y, r[5].z);
    R.hblock = (int)pool_f_bits(r[5].w);
    R.hnormal = V3(r[6].x, r[6].y, r[6].z);
    k = pool_f_bits(r[6].w);
    R.direct = V3(r[7].x, r[7].y, r[7].z);
    first_t = r[7].w;
    R.color = V3(r[8].x, r[8].y, r[8].z);
}
DDGI_HD void pool_pack(const WfRay& R, uint32_t k, float first_t, PoolVec* r)
{
    pool_pack_march(R, r);
    r[2] = pool_vec(R.inv, pool_bits_f(R.lookups));
    r[4] = pool_vec(R.qd, pool_bits_f((uint32_t)R.bounce | ((uint32_t)R.phase << 8) | ((uint32_t)R.visible << 16)));
    r[5] = pool_vec(R.hpos, pool_bits_f((uint32_t)R.hblock));
    r[6] = pool_vec(R.hnormal, pool_bits_f(k));
    r[7] = pool_vec(R.direct, first_t);
    r[8] = pool_vec(R.color, 0.0f);
}

// Queue a ray waits in for its state (FETCH also holds the pool's empty slots at start).
enum : int { PQ_MARCH = 0, PQ_BOUNCE = 1, PQ_FEELER = 2, PQ_FETCH = 3, PQ_SLOW = 4, PQ_COUNT = 5 };
DDGI_HD int pool_queue_of(int mode)
{
    switch (mode) {
        case WF_MARCH: return PQ_MARCH;
        case WF_BOUNCE_HIT: return PQ_BOUNCE;
        case WF_FEELER_HIT: return PQ_FEELER;
        case WF_MARCH_SLOW: return PQ_SLOW;
        default: return PQ_FETCH;
    }
}

}  // namespace ddgi
