// ddgi_kernels.cu — sm_100a kernels of the DDGI probe-field engine.
//
//   probe_update_wavefront   persistent warps, one probe ray per lane as a state machine
//                            (ddgi_wavefront.cuh); the default probe update
//   probe_update_direct      one thread per probe ray, reference loop order
//                            (assets/shaders/probe_pass.comp:253-303)
//   probe_blend_octahedral   optional textbook layout: ray results -> octahedral tile texels,
//                            warp-shuffle reduction (ddgi_octahedral.cuh)
//   peer_barrier_kernel      completion barrier of the fused multi-GPU exchange
//   render_frame_kernel      one thread per pixel (assets/shaders/compute_pass.comp:162-191)
//   bake_* / build_occupancy / edit_voxels   scene preparation and per-frame edits
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false (see csrc/Makefile).
#include "ddgi_internal.h"
#include "ddgi_shade.cuh"
#include "ddgi_wavefront.cuh"

namespace ddgi {

// ------------------------------------------------------------------ ray fetch
struct RayIn {
    v3 origin, direction;
    int tx, ty;  // destination texel
};

// Ray k either from the caller's storage buffer (48-byte records, three 16-byte
// loads) or derived from the lattice formula and the per-probe direction table.
__device__ __forceinline__ RayIn fetch_ray(const FrameParams& P, const ProbeJob& J, uint32_t k)
{
    RayIn r;
    int tiles_x = P.probe_count[0] * P.probe_count[2];
    if (J.rays) {
        float4 a = __ldg(J.rays + 3 * (size_t)k);
        float4 b = __ldg(J.rays + 3 * (size_t)k + 1);
        float4 c = __ldg(J.rays + 3 * (size_t)k + 2);
        r.origin = V3(a.x, a.y, a.z);
        r.direction = V3(b.x, b.y, b.z);
        int p = f2i(c.x);
        int yp = p / tiles_x;
        int xp = p - yp * tiles_x;
        r.tx = xp * P.rx + f2i(c.y);
        r.ty = yp * P.ry + f2i(c.z);
    } else {
        int n = P.rx * P.ry;
        int p = (int)(k / (uint32_t)n);
        int i = (int)(k - (uint32_t)p * (uint32_t)n);
        r.origin = probe_origin(P, p);
        r.direction = V3(__ldg(J.dirs + 3 * i), __ldg(J.dirs + 3 * i + 1), __ldg(J.dirs + 3 * i + 2));
        int yp = p / tiles_x;
        int xp = p - yp * tiles_x;
        int iy = i / P.rx;
        r.tx = xp * P.rx + (i - iy * P.rx);
        r.ty = yp * P.ry + iy;
    }
    return r;
}

// Destination texel of ray k (what fetch_ray also returns): the wavefront kernel recomputes it when the ray is
// stored instead of holding two registers for the ray's whole life.
__device__ __forceinline__ void texel_of_ray(const FrameParams& P, const ProbeJob& J, uint32_t k, int* tx, int* ty)
{
    int tiles_x = P.probe_count[0] * P.probe_count[2];
    int p, x, y;
    if (J.rays) {
        float4 c = __ldg(J.rays + 3 * (size_t)k + 2);
        p = f2i(c.x);
        x = f2i(c.y);
        y = f2i(c.z);
    } else {
        int n = P.rx * P.ry;
        p = (int)(k / (uint32_t)n);
        int i = (int)(k - (uint32_t)p * (uint32_t)n);
        y = i / P.rx;
        x = i - y * P.rx;
    }
    int yp = p / tiles_x;
    int xp = p - yp * tiles_x;
    *tx = xp * P.rx + x;
    *ty = yp * P.ry + y;
}

// Linear ray index (the reference's position in the ProbeRay list) of the idx-th ray of this shard.
__device__ __forceinline__ uint32_t shard_ray(const ProbeJob& J, uint32_t idx)
{
    uint32_t slot = idx / J.slot_rays;
    uint32_t i = idx - slot * J.slot_rays;
    return __ldg(J.order + slot) * J.slot_rays + i;
}

// `first_t`: t of the ray's first nearest-hit query (INF on a miss); only read in distance mode 1.
__device__ __forceinline__ void store_texel(const ProbeJob& J, int tx, int ty, v3 color, uint32_t k,
                                            uint32_t lookups, float first_t)
{
    if (J.ray_out) {
        // octahedral layout: the ray's radiance and first-hit t go to the ray buffer; the tile
        // texels are made from all rays of the probe by probe_blend_octahedral
        J.ray_out[k] = make_float4(color.x, color.y, color.z, first_t);
        if (J.lookups) J.lookups[k] = lookups;
        if (J.slot_cost) atomicMax(J.slot_cost + k / J.slot_rays, lookups);
        return;
    }
    // literal storage-buffer mode: the texel comes from the caller's probe_info floats; an
    // out-of-range imageStore is discarded, as Vulkan does for the reference (probe_pass.comp:301-302)
    if (J.rays && ((unsigned)tx >= (unsigned)J.tex_w || (unsigned)ty >= (unsigned)J.tex_h)) return;
    size_t t = (size_t)ty * J.tex_w + tx;
    if (J.blend) color = blend_hysteresis(J.albedo_old[t], color, J.hysteresis);
    uint32_t rgba = pack_rgba8(color.x, color.y, color.z, 1.0f);
    J.albedo[t] = rgba;
    // probe_pass.comp:276,302: distances = vec2(0) as shipped; distance mode 1 stores the
    // first-hit moments (d, d*d), d = t / distance_scale
    uint32_t moments = 0u;
    if (J.distance_mode == 1) {
        float d = first_t / J.distance_scale;
        moments = pack_rgba8(d, d * d, 0.0f, 0.0f);
    }
    // (J.distance == nullptr: the plane is known to hold the zeros the reference would store again)
    if (J.distance) J.distance[t] = moments;
#pragma unroll 1
    for (int g = 0; g < J.n_peers; g++) {
        J.peer_albedo[g][t] = rgba;
        if (J.distance) J.peer_distance[g][t] = moments;
    }
    if (J.albedo_f32) J.albedo_f32[t] = make_float4(color.x, color.y, color.z, 1.0f);
    if (J.lookups) J.lookups[k] = lookups;
    if (J.slot_cost) atomicMax(J.slot_cost + k / J.slot_rays, lookups);
}

// ------------------------------------------------------------------ variant 0
__global__ void __launch_bounds__(256) probe_update_direct(const __grid_constant__ FrameParams P,
                                                           const __grid_constant__ ProbeJob J)
{
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= J.n_owned * J.slot_rays) return;
    uint32_t k = shard_ray(J, idx);
    RayIn r = fetch_ray(P, J, k);
    uint32_t lookups = 0;
    float first_t = 0.0f;
    v3 color = trace_probe_ray(P, r.origin, r.direction, k, lookups, &first_t);
    store_texel(J, r.tx, r.ty, color, k, lookups, first_t);
}

// ------------------------------------------------------------------ variant 1
// Persistent warps over a global ray counter; each lane owns one probe ray as a WfRay
// state machine (ddgi_wavefront.cuh).  The warp steps the DDA march while at least
// march_min/32 of its rays are marching; below that it counts its lanes per remaining
// state (ballots), picks the fullest one and runs that state's code for exactly those
// lanes.  The other lanes wait and accumulate, so each code block executes with many
// lanes active instead of once per divergent lane group.  Every ended march - a bounce ray's
// or a shadow feeler's - waits in the same state (WF_HIT), so what the two have in common
// is issued once for both.
// Finished rays store their texel in WF_FETCH and take the next ray index there (the
// warp draws indices from the global counter 32 at a time).
// Threads per block: the kernel is compiled for up to kWfThreads; the warps of a block share nothing, so the block size
// only decides how many warps must have drained before the SM takes the next block (of this launch or, with two frames in
// flight, of the next frame's).  The launcher picks ONE warp per block when there are at least 7 rays per resident lane
// (field_32 6.410 -> 6.361 ms; 28 blocks of 32 threads are resident per SM) and four warps per block for smaller
// workloads, which are bound by their longest ray and measured up to 10 % slower with one-warp blocks
// (profiles/r2_ab.md h, p).  DDGI_WF_THREADS forces one size (A/B builds).
#ifndef DDGI_WF_THREADS
#define DDGI_WF_THREADS 0
#endif
constexpr int kWfThreads = 128;
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#ifndef DDGI_ENOUGH_HOISTED
#define DDGI_ENOUGH_HOISTED 1
#endif
#ifndef DDGI_WF_MIN_BLOCKS
#define DDGI_WF_MIN_BLOCKS 6  // asks for 6 blocks of 128; the kernel needs 72 registers, so 7 are resident (with a bound of 7 ptxas works AT its register limit and generates two more instructions in the march loop): 72 registers / thread, 28 resident warps per SM: measured faster than 8 (64 registers, spills)
#endif

// kCount: the per-ray voxel-lookup count is kept (debug buffers, the calibration launch of the schedule); the
// normal launch does not carry the counter.
template <bool kLiteral, bool kTimed, bool kCount>
__global__ void __launch_bounds__(kWfThreads, DDGI_WF_MIN_BLOCKS) probe_update_wavefront(const __grid_constant__ FrameParams P,
                                                                     const __grid_constant__ ProbeJob J,
                                                                     uint32_t* __restrict__ next_ray,
                                                                     int march_min, int lanes_used)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t n_rays = J.n_owned * J.slot_rays;
    __shared__ float s_base[kLiteral ? 3 * kWfThreads : 1];  // procedural colour of the lane's bounce hit
    __shared__ float s_first_t[kWfThreads];                  // t of the lane's first query (distance mode 1)
    float* stash = s_base + (kLiteral ? threadIdx.x : 0);
    const bool want_first_t = J.distance_mode == 1 || J.ray_out != nullptr;
    WfRay R;
    // (a workload with fewer rays than the GPU has resident lanes is bound by the longest warp, not by
    // throughput: it is spread over all resident warps with only `lanes_used` lanes of each holding rays)
    R.mode = lane < lanes_used ? WF_FETCH : WF_IDLE;
    uint32_t k = 0xffffffffu;  // no ray yet
    uint32_t chunk_next = 0, chunk_end = 0;  // warp-uniform: ray indices already reserved
    bool exhausted = false;
    // debug level 2 (kTimed instantiation only — measured: even a never-taken test of
    // J.warp_times in the fetch path costs the normal kernel 1.6 %): this warp's (start, last
    // ray taken, exit) times
#define DDGI_WARP_TIME(slot)                                                                                  \
    do {                                                                                                      \
        if (kTimed)                                                                                           \
            J.warp_times[3 * (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) + (slot)] = globaltimer_ns(); \
    } while (0)
    if (lane == 0) DDGI_WARP_TIME(0);

    // lanes that hold a ray or may still take one (FETCH retires lanes) in the low byte, the number of scheduling
    // rounds above it; warp-uniform, one register
    int n_live = lanes_used;
    // lanes that must be marching for the march loop to go on: march_min/32 of the live lanes, at least one.
    // (Kept in a register the compiler cannot re-derive - DDGI_OPAQUE - or it recomputes the five
    // instructions from n_live and march_min in EVERY march step to save that register.)
#if DDGI_ENOUGH_HOISTED
#define DDGI_OPAQUE(x) asm volatile("" : "+r"(x))
#else
#define DDGI_OPAQUE(x)
#endif
    int enough = (n_live * march_min + 31) >> 5 > 1 ? (n_live * march_min + 31) >> 5 : 1;
    DDGI_OPAQUE(enough);
    for (;;) {
        // ---- march while at least march_min/32 of the lanes holding a ray are marching ----
        if ((n_live & 255) == 0) break;
        while (__popc(__ballot_sync(full, R.mode == WF_MARCH)) >= enough) {
            if (R.mode == WF_MARCH) wf_step(P, R);
        }
        wf_end_march(R);
        // ---- otherwise run the fullest of the other states (ties: the later stage).  Ended marches
        //      of bounce rays and of shadow feelers run the same code (wf_resolve_hit) but are
        //      scheduled as two states: the kind that waits keeps accumulating lanes, which the
        //      model (profiles/policy_sim.py) and round 1's measurements favour over one merged pass ----
        const bool hit_feeler = R.mode == WF_HIT && R.phase != 0, hit_bounce = R.mode == WF_HIT && R.phase == 0;
        const int n_hit_f = __popc(__ballot_sync(full, hit_feeler));
        const int n_hit_b = __popc(__ballot_sync(full, hit_bounce));
        const unsigned fetching = __ballot_sync(full, R.mode == WF_FETCH);
        const int n_fetch = __popc(fetching);
        const int n_hit = n_hit_f > n_hit_b ? n_hit_f : n_hit_b;
        // (marches with the literal arithmetic - axis-parallel or degenerate rays - are rare: they are only looked
        // for when nothing else waits, or once in a while so that they cannot starve)
        int n_slow = 0;
        n_live += 256;
        if (n_hit + n_fetch == 0 || (n_live & 0x700) == 0) {
            n_slow = __popc(__ballot_sync(full, R.mode == WF_MARCH_SLOW));
            if (n_hit + n_fetch + n_slow == 0) continue;  // (only marching lanes: cannot be fewer than `enough`)
        }

        if (n_slow > n_hit && n_slow > n_fetch) {
            if (R.mode == WF_MARCH_SLOW) {
                wf_step_literal(P, R);
                wf_end_march(R);
            }
        } else if (n_hit > n_fetch) {
            if (n_hit_f >= n_hit_b ? hit_feeler : hit_bounce) {
                float* first_t = (want_first_t && R.bounce == 0 && R.phase == 0) ? &s_first_t[threadIdx.x] : nullptr;
                wf_resolve_hit<kLiteral>(P, R, stash, kWfThreads, first_t);
                if (!kCount) R.lookups = 0;
            }
        } else {
            // WF_FETCH: store the finished ray, take the next one
            const bool need = R.mode == WF_FETCH;
            if (need && k != 0xffffffffu) {
                int tx, ty;
                texel_of_ray(P, J, k, &tx, &ty);
                store_texel(J, tx, ty, wf_final_color(P, R), k, kCount ? R.lookups : 0u, want_first_t ? s_first_t[threadIdx.x] : 0.0f);
            }
            const unsigned want = fetching;
            uint32_t cnt = (uint32_t)__popc(want);
            uint32_t rank = (uint32_t)__popc(want & ((1u << lane) - 1u));
            uint32_t avail = chunk_end - chunk_next;
            uint32_t idx;
            if (cnt <= avail) {
                idx = chunk_next + rank;
                chunk_next += cnt;
            } else {
                // Reserve ray indices for the warp: 32 at a time while plenty are left (one
                // atomic per 32 rays, and a warp's lanes work on neighbouring rays), exactly the
                // number asked for once fewer than two rays per resident lane remain — a warp
                // must not sit on reserved rays while others run dry, or it alone is the tail.
                const uint32_t more = cnt - avail;
                const uint32_t take = (n_rays - chunk_end <= 2u * gridDim.x * blockDim.x) ? more : (uint32_t)lanes_used;
                uint32_t base = n_rays;
                if (!exhausted) {
                    if (lane == 0) base = atomicAdd(next_ray, take);
                    base = __shfl_sync(full, base, 0);
                }
                uint32_t nend = base + take < n_rays ? base + take : n_rays;
                if (base >= n_rays) {
                    exhausted = true;
                    base = nend = n_rays;
                }
                idx = rank < avail ? chunk_next + rank : base + (rank - avail);
                chunk_next = base + more;
                chunk_end = nend;
                if (chunk_next > chunk_end) chunk_next = chunk_end;
            }
            if (need) {
                if (idx < n_rays) {
                    DDGI_WARP_TIME(1);  // (any lane: the last writer wins, same instant)
                    k = shard_ray(J, idx);
                    RayIn r = fetch_ray(P, J, k);
                    wf_init(R, r.origin, r.direction, k);
                } else {
                    R.mode = WF_IDLE;
                }
            }
            const int retired = __popc(__ballot_sync(full, need && R.mode == WF_IDLE));
            if (retired) {
                n_live -= retired;
                enough = ((n_live & 255) * march_min + 31) >> 5 > 1 ? ((n_live & 255) * march_min + 31) >> 5 : 1;
                DDGI_OPAQUE(enough);
            }
        }
        // a finished bounce scatters at once, and every state above hands over to WF_QUERY or
        // ends the ray: arm the new queries right away.  Neither is scheduled as a state of
        // its own (one scheduling round less per bounce and per query).
        if (R.mode == WF_SCATTER) wf_scatter(P, R);
        if (R.mode == WF_QUERY) wf_begin_query(P, R);
    }
    if (lane == 0) DDGI_WARP_TIME(2);
#undef DDGI_WARP_TIME
}

// ------------------------------------------------------------------ octahedral blend
// One block per owned probe.  The probe's ray directions and ray results are staged in shared
// memory; each warp takes texels of the oct x oct tile in turn: lanes accumulate their share of
// the rays (ddgi_octahedral.cuh: oct_lane_partial), the partial sums meet in an xor-butterfly
// of warp shuffles — the north star's "warp-shuffle reduction of ray radiance into texels" — and
// lane 0 blends and stores the two texels (into every replica under the fused exchange).
#ifndef DDGI_OCT_TMA
#define DDGI_OCT_TMA 1
#endif
constexpr int kOctThreads = 128;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ OctAcc oct_shuffle_xor(OctAcc a, int off)
{
    OctAcc b;
    b.r = __shfl_xor_sync(0xffffffffu, a.r, off);
    b.g = __shfl_xor_sync(0xffffffffu, a.g, off);
    b.b = __shfl_xor_sync(0xffffffffu, a.b, off);
    b.d = __shfl_xor_sync(0xffffffffu, a.d, off);
    b.d2 = __shfl_xor_sync(0xffffffffu, a.d2, off);
    b.w = __shfl_xor_sync(0xffffffffu, a.w, off);
    return b;
}
__global__ void __launch_bounds__(kOctThreads) probe_blend_octahedral(const __grid_constant__ FrameParams P,
                                                                      const __grid_constant__ OctJob J)
{
    extern __shared__ __align__(16) float s_oct[];
    __shared__ __align__(8) unsigned long long s_bar;
    float* s_dirs = s_oct;                 // n x 3
    float* s_rad = s_oct + 3 * J.n_rays;   // n x 4
    const uint32_t p = J.probes[blockIdx.x];
    const float4* mine = J.ray_out + (size_t)p * J.n_rays;
    if (DDGI_OCT_TMA && J.n_rays % 4 == 0) {
        // TMA staging: the probe's ray directions (12 n bytes) and ray results (16 n bytes) are two contiguous,
        // 16-byte aligned ranges that every warp of the block reads for every texel of the tile.  One thread
        // issues two bulk copies (cp.async.bulk global -> shared, completion counted in bytes on an mbarrier);
        // the block waits on the barrier's first phase - no per-thread load / store loop, no register staging.
        const uint32_t bar = smem_u32(&s_bar);
        const uint32_t bytes_dirs = 12u * (uint32_t)J.n_rays, bytes_rad = 16u * (uint32_t)J.n_rays;
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes_dirs + bytes_rad) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s_dirs)),
                         "l"(J.dirs), "r"(bytes_dirs), "r"(bar)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s_rad)),
                         "l"(mine), "r"(bytes_rad), "r"(bar)
                         : "memory");
        }
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                         : "=r"(done)
                         : "r"(bar)
                         : "memory");
        }
    } else {
        // (a ray count that is not a multiple of 4 breaks the 16-byte granularity of the bulk copy: plain loads)
        for (int i = threadIdx.x; i < 3 * J.n_rays; i += kOctThreads) s_dirs[i] = J.dirs[i];
        for (int i = threadIdx.x; i < J.n_rays; i += kOctThreads) {
            float4 v = mine[i];
            s_rad[4 * i] = v.x;
            s_rad[4 * i + 1] = v.y;
            s_rad[4 * i + 2] = v.z;
            s_rad[4 * i + 3] = v.w;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int cx, cy;
    tile_origin(P, (int)p, &cx, &cy);
    for (int t = warp; t < P.oct * P.oct; t += kOctThreads / 32) {
        int u = t % P.oct, v = t / P.oct;
        v3 dir_t = oct_texel_dir(u, v, P.oct);
        OctAcc a = oct_lane_partial(dir_t, s_dirs, s_rad, J.n_rays, lane, J.distance_scale);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) a = oct_add(a, oct_shuffle_xor(a, off));
        if (lane == 0) {
            size_t at = (size_t)(cy + v) * J.tex_w + (cx + u);
            uint32_t alb, dist;
            oct_finalize(a, J.blend, J.hysteresis, J.albedo_old[at], J.distance_old[at], &alb, &dist);
            J.albedo[at] = alb;
            J.distance[at] = dist;
            for (int g = 0; g < J.n_peers; g++) {
                J.peer_albedo[g][at] = alb;
                J.peer_distance[g][at] = dist;
            }
        }
    }
}

cudaError_t launch_probe_blend_octahedral(const FrameParams& P, const OctJob& J, cudaStream_t s, int* launches)
{
    if (J.n_probes == 0) return cudaSuccess;
    size_t smem = (size_t)7 * J.n_rays * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(probe_blend_octahedral, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    probe_blend_octahedral<<<J.n_probes, kOctThreads, smem, s>>>(P, J);
    (*launches)++;
    return cudaGetLastError();
}

// ------------------------------------------------------------------ fused-exchange barrier
// Runs after the probe update on the same stream (so every texel this rank stored into its
// peers' replicas has been performed): lane g publishes `epoch` in peer g's flag slot for this
// rank, then waits until peer g's epoch has arrived in the local slot — i.e. until g's texels
// are in the local replica.  One launch, no collective library on the data path.  A peer that
// never arrives ends the wait after `timeout_ns` with *error set (never a hang).
__global__ void peer_barrier_kernel(PeerBarrier B)
{
    const int g = threadIdx.x;
    if (g >= B.n_ranks || g == B.self) return;
    __threadfence_system();
    // (a maximum, not a store: with two frames in flight the barriers of consecutive frames run on two
    // streams and may publish out of order; epochs only grow)
    atomicMax_system(B.peer_flags[g] + B.self, B.epoch);
    __threadfence_system();
    const volatile uint32_t* mine = B.local_flags + g;
    const unsigned long long t0 = globaltimer_ns();
    // epochs only grow; the signed difference keeps the test right across a wrap
    while ((int32_t)(*mine - B.epoch) < 0) {
        if (globaltimer_ns() - t0 > B.timeout_ns) {
            atomicExch(B.error, 1u);
            return;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

cudaError_t launch_peer_barrier(const PeerBarrier& B, cudaStream_t s, int* launches)
{
    peer_barrier_kernel<<<1, 32, 0, s>>>(B);
    (*launches)++;
    return cudaGetLastError();
}

// ------------------------------------------------------------------ packed all-gather (probe-cyclic ownership)
// Probes dealt round-robin in blocks of B to G ranks scatter a rank's tiles over the whole texture; the collective
// wants ONE contiguous chunk per rank.  pack_tiles copies the tiles of the owned probes (list `owned`, ascending)
// into the rank's chunk of the gather buffer, tile after tile (both planes when the distance plane is in use);
// after the in-place ncclAllGather unpack_tiles scatters every OTHER rank's chunk to its tiles: the j-th probe of
// rank g is probe ((j / B) * G + g) * B + j % B.
__device__ __forceinline__ void copy_tile(const TilePack& T, int p, size_t slot, bool to_pack)
{
    const int n = T.tw * T.th;
    const int ox = (p % T.tiles_x) * T.tw, oy = (p / T.tiles_x) * T.th;
    for (int pl = 0; pl < T.planes; pl++) {
        uint32_t* tex = T.tex + pl * T.plane;
        uint32_t* pk = T.pack + slot + (size_t)pl * T.max_owned * n;
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            const int row = t / T.tw, col = t - row * T.tw;
            const size_t at = (size_t)(oy + row) * T.tex_w + ox + col;
            if (to_pack) pk[t] = tex[at];
            else tex[at] = pk[t];
        }
    }
}
__global__ void __launch_bounds__(128) pack_tiles_kernel(const __grid_constant__ TilePack T)
{
    const int j = blockIdx.x;
    if (j >= T.n_owned) return;
    copy_tile(T, (int)T.owned[j], (size_t)T.self * T.chunk_texels + (size_t)j * T.tw * T.th, true);
}
__global__ void __launch_bounds__(128) unpack_tiles_kernel(const __grid_constant__ TilePack T)
{
    const int j = blockIdx.x, g = blockIdx.y;
    if (g == T.self) return;
    const int p = ((j / T.B) * T.G + g) * T.B + j % T.B;
    if (p >= T.n_probes) return;
    copy_tile(T, p, (size_t)g * T.chunk_texels + (size_t)j * T.tw * T.th, false);
}
cudaError_t launch_pack_tiles(const TilePack& T, bool unpack, cudaStream_t s, int* launches)
{
    if (T.max_owned == 0) return cudaSuccess;
    if (unpack) unpack_tiles_kernel<<<dim3(T.max_owned, T.G), 128, 0, s>>>(T);
    else if (T.n_owned) pack_tiles_kernel<<<T.n_owned, 128, 0, s>>>(T);
    else return cudaSuccess;
    (*launches)++;
    return cudaGetLastError();
}

// ------------------------------------------------------------------ pixel pass
// The reference dispatches floor(w/16) x floor(h/16) groups of 16x16: pixels beyond
// that are never written (src/rvpt/rvpt.cpp:1139-1140).
// kExt = false: the DDGI frame as shipped; true: debug integrators (render_mode 1-5), probe
// markers, restored Chebyshev weight.
template <bool kExt>
__global__ void __launch_bounds__(256) render_frame_kernel(const __grid_constant__ FrameParams P,
                                                           const __grid_constant__ PixelJob J)
{
    int gx = blockIdx.x * 16 + (threadIdx.x & 15);
    int gy = (J.group_row0 + blockIdx.y) * 16 + (threadIdx.x >> 4);
    float cx = (float)gx / (float)P.screen_w;
    float cy = (float)gy / (float)P.screen_h;
    cy = 1.0f - cy;
    v3 o, d;
    pinhole_ray(P, cx, cy, &o, &d);
    uint32_t lookups = 0;
    v3 s = shade_pixel<kExt>(P, J.albedo, J.distance, J.tex_w, o, d, lookups);
    s = V3(0, 0, 0) + s;  // `sampled += ...`
    size_t at = (size_t)gy * P.screen_w + gx;
    J.frame[at] = pack_rgba8(s.x, s.y, s.z, 1.0f);
    if (J.frame_f32) J.frame_f32[at] = make_float4(s.x, s.y, s.z, 1.0f);
    if (J.lookups) J.lookups[at] = lookups;
}

// ------------------------------------------------------------------ scene preparation
__global__ void bake_scene_kernel(int scene, int dx, int dy, int dz, int ox, int oy, int oz, uint8_t* types)
{
    size_t n = (size_t)dx * dy * dz;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % dx);
        int y = (int)((i / dx) % dy);
        int z = (int)(i / ((size_t)dx * dy));
        v3 c = V3((float)(x + ox), (float)(y + oy), (float)(z + oz));
        types[i] = (uint8_t)block_procedural(c, scene);
    }
}

// Integer-only synthetic cave: the reference cave's four spheres (radii 20,20,18,21 at
// offsets (0,0,0), (-16,-8,10), (13,1,-19), (-20,-15,-15), intersection.glsl:744-750)
// scaled by dims/64 around the grid centre; rock (type 10) outside their union, and
// inside it `permille` of the cells solid with one of six flat block types 2..7.
__global__ void bake_synthetic_kernel(int dx, int dy, int dz, int permille, uint32_t seed, uint8_t* types)
{
    const int sph[4][4] = {{0, 0, 0, 20}, {-16, -8, 10, 20}, {13, 1, -19, 18}, {-20, -15, -15, 21}};
    size_t n = (size_t)dx * dy * dz;
    int mind = min(dx, min(dy, dz));
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % dx);
        int y = (int)((i / dx) % dy);
        int z = (int)(i / ((size_t)dx * dy));
        // position relative to the grid centre in 1/64ths of the grid: compare
        // |64*(g - dim/2) - mind*centre|^2 < (mind*radius)^2 in 64-bit integers
        bool inside = false;
        for (int s = 0; s < 4; s++) {
            long long ex = 64ll * x - 32ll * dx - (long long)mind * sph[s][0];
            long long ey = 64ll * y - 32ll * dy - (long long)mind * sph[s][1];
            long long ez = 64ll * z - 32ll * dz - (long long)mind * sph[s][2];
            long long rr = (long long)mind * sph[s][3];
            if (ex * ex + ey * ey + ez * ez < rr * rr) inside = true;
        }
        uint8_t t = 10;
        if (inside) {
            uint32_t h = wang_hash((uint32_t)i ^ seed);
            h ^= (uint32_t)(i >> 32) * 0x9E3779B9u;
            h ^= h << 13;
            h ^= h >> 17;
            h ^= h << 5;
            t = (h % 1000u) < (uint32_t)permille ? (uint8_t)(2 + (h >> 10) % 6u) : 0;
        }
        types[i] = t;
    }
}

// One thread per brick of the brick box [b0, b0 + bn): gathers 32 type bytes into the
// occupancy word.  The whole grid after an upload / bake, the touched bricks after an edit.
__global__ void build_occupancy_kernel(int dx, int dy, int dz, int sx, int sy, int sz, int nbx, int nby, int b0x, int b0y,
                                       int b0z, int bnx, int bny, int bnz, const uint8_t* types, uint32_t* occ)
{
    size_t nb = (size_t)bnx * bny * bnz;
    for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < nb; b += (size_t)gridDim.x * blockDim.x) {
        int bx = b0x + (int)(b % bnx);
        int by = b0y + (int)((b / bnx) % bny);
        int bz = b0z + (int)(b / ((size_t)bnx * bny));
        uint32_t w = 0u;
        for (int z = 0; z < (1 << kBrickLz); z++)
            for (int y = 0; y < (1 << kBrickLy); y++)
                for (int x = 0; x < (1 << kBrickLx); x++) {
                    // grid cell of this brick cell: bricks start (sx,sy,sz) cells before the grid
                    const int cx = (bx << kBrickLx) + x, cy = (by << kBrickLy) + y, cz = (bz << kBrickLz) + z;
                    int gx = cx - sx, gy = cy - sy, gz = cz - sz;
                    if (gx >= 0 && gy >= 0 && gz >= 0 && gx < dx && gy < dy && gz < dz &&
                        types[((size_t)gz * dy + gy) * dx + gx] != 0)
                        w |= occ_mask(occ_shift(cx, cy, cz));
                }
        occ[((size_t)bz * nby + by) * nbx + bx] = w;
    }
}

// Per-frame voxel edit: copies an ex x ey x ez box of block types (x fastest) into the field at
// grid cell (x0, y0, z0).
__global__ void edit_voxels_kernel(int dx, int dy, int x0, int y0, int z0, int ex, int ey, int ez, const uint8_t* src,
                                   uint8_t* types)
{
    size_t n = (size_t)ex * ey * ez;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % ex), y = (int)((i / ex) % ey), z = (int)(i / ((size_t)ex * ey));
        types[((size_t)(z0 + z) * dy + (y0 + y)) * dx + (x0 + x)] = src[i];
    }
}

// ------------------------------------------------------------------ launchers
// Launch shape of the wavefront kernel for n rays: the resident grid (148 SMs x blocks per SM) when there is
// enough work to give every lane a ray; fewer blocks than that only when even 8 rays per warp do not fill them.
// *lanes = lanes per warp that hold rays: 32, or fewer (a multiple of 4, at least 8) when n is below the number
// of resident lanes, so that a small workload uses every resident warp with few rays each instead of a quarter
// of the warps with 32 each (cave_64: 65 536 rays on 132 608 resident lanes).
uint32_t wavefront_warps(uint32_t n, int grid_limit, int* lanes, int* threads)
{
    static int blocks_per_sm = 0, sms = 0;
    if (!blocks_per_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, probe_update_wavefront<false, false, false>, kWfThreads, 0);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    int per_sm = grid_limit > 0 && grid_limit < blocks_per_sm ? grid_limit : blocks_per_sm;  // (in blocks of kWfThreads)
    uint32_t resident_warps = (uint32_t)(sms * per_sm) * (kWfThreads / 32);
    int use = 32;
    if (n < resident_warps * 32u) {
        use = (int)((n + resident_warps - 1) / resident_warps);
        use = (use + 3) & ~3;
        use = use < 8 ? 8 : (use > 32 ? 32 : use);
    }
    if (lanes) *lanes = use;
    // one warp per block from 7 rays per resident lane on, else four (see kWfThreads: sweep_64 / sweep_128, 2 - 4 rays per
    // lane, are 7 - 10 % slower with one-warp blocks, 8 rays per lane are even, field_32 with 63 gains 0.8 %)
    int t = DDGI_WF_THREADS ? DDGI_WF_THREADS : (n >= resident_warps * 224u ? 32 : kWfThreads);
    if (threads) *threads = t;
    uint32_t warps_needed = (n + use - 1) / use;
    uint32_t warps = warps_needed < resident_warps ? warps_needed : resident_warps;
    uint32_t per_block = (uint32_t)t / 32u;
    return (warps + per_block - 1) / per_block * per_block;
}

cudaError_t launch_probe_update(const FrameParams& P, const ProbeJob& J, int variant, uint32_t* counter,
                                int march_min, int grid_limit, cudaStream_t s, int* launches)
{
    uint32_t n = J.n_owned * J.slot_rays;
    if (n == 0) return cudaSuccess;
    if (variant == 0 || P.max_bounces <= 0) {
        dim3 block(256), grid((n + 255) / 256);
        probe_update_direct<<<grid, block, 0, s>>>(P, J);
        (*launches)++;
        return cudaGetLastError();
    }
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    int lanes = 32, threads = kWfThreads;
    uint32_t grid = wavefront_warps(n, grid_limit, &lanes, &threads) / (uint32_t)(threads / 32);
    const bool literal = P.scene.color_mode != 0, timed = J.warp_times != nullptr;
    const bool count = J.lookups != nullptr || J.slot_cost != nullptr;  // (the timed instantiation always counts)
#define DDGI_LAUNCH_WF(L, T, C) probe_update_wavefront<L, T, C><<<grid, threads, 0, s>>>(P, J, counter, march_min, lanes)
    if (literal && timed) DDGI_LAUNCH_WF(true, true, true);
    else if (literal && count) DDGI_LAUNCH_WF(true, false, true);
    else if (literal) DDGI_LAUNCH_WF(true, false, false);
    else if (timed) DDGI_LAUNCH_WF(false, true, true);
    else if (count) DDGI_LAUNCH_WF(false, false, true);
    else DDGI_LAUNCH_WF(false, false, false);
#undef DDGI_LAUNCH_WF
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_render_frame(const FrameParams& P, const PixelJob& J, cudaStream_t s, int* launches)
{
    dim3 grid(P.screen_w / 16, J.group_rows);  // the reference dispatches floor(w/16) x floor(h/16) groups
    if (grid.x == 0 || grid.y == 0) return cudaSuccess;
    bool ext = (P.render_mode >= 1 && P.render_mode <= 5) || P.visualize_probes != 0 || P.weight_mode != 0 || P.layout != 0;
    if (ext) render_frame_kernel<true><<<grid, 256, 0, s>>>(P, J);
    else render_frame_kernel<false><<<grid, 256, 0, s>>>(P, J);
    (*launches)++;
    return cudaGetLastError();
}

static int grid_for(size_t n)
{
    size_t g = (n + 255) / 256;
    size_t cap = 148 * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

cudaError_t launch_bake_scene(int scene, const int dims[3], const int org[3], uint8_t* types,
                              cudaStream_t s, int* launches)
{
    size_t n = (size_t)dims[0] * dims[1] * dims[2];
    bake_scene_kernel<<<grid_for(n), 256, 0, s>>>(scene, dims[0], dims[1], dims[2], org[0], org[1], org[2], types);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_bake_synthetic(const int dims[3], const int org[3], int permille, uint32_t seed,
                                  uint8_t* types, cudaStream_t s, int* launches)
{
    (void)org;
    size_t n = (size_t)dims[0] * dims[1] * dims[2];
    bake_synthetic_kernel<<<grid_for(n), 256, 0, s>>>(dims[0], dims[1], dims[2], permille, seed, types);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_build_occupancy(const int dims[3], const int shift[3], const int nb[3], const int b0[3], const int bn[3],
                                   const uint8_t* types, uint32_t* occ, cudaStream_t s, int* launches)
{
    size_t n = (size_t)bn[0] * bn[1] * bn[2];
    if (n == 0) return cudaSuccess;
    build_occupancy_kernel<<<grid_for(n), 256, 0, s>>>(dims[0], dims[1], dims[2], shift[0], shift[1], shift[2], nb[0], nb[1], b0[0],
                                                        b0[1], b0[2], bn[0], bn[1], bn[2], types, occ);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_edit_voxels(const int dims[3], const int at[3], const int ext[3], const uint8_t* src, uint8_t* types,
                               cudaStream_t s, int* launches)
{
    size_t n = (size_t)ext[0] * ext[1] * ext[2];
    if (n == 0) return cudaSuccess;
    edit_voxels_kernel<<<grid_for(n), 256, 0, s>>>(dims[0], dims[1], at[0], at[1], at[2], ext[0], ext[1], ext[2], src, types);
    (*launches)++;
    return cudaGetLastError();
}

}  // namespace ddgi
