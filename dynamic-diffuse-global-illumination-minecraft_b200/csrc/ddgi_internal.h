// ddgi_internal.h — launcher interface between the C-ABI host code (ddgi_engine.cu)
// and the kernels (ddgi_kernels.cu).  Not installed; the public surface is include/ddgi.h.
#pragma once
#include <cuda_runtime.h>

#include "ddgi_trace.cuh"

namespace ddgi {

constexpr int kMaxPeers = 8;

// Everything the probe-update kernels need besides FrameParams.
struct ProbeJob {
    const float4* rays;   // literal 48-byte ProbeRay records (3 x float4) or nullptr
    const float* dirs;    // generated mode: rx*ry normalised directions (xyz)
    // The unit of scheduling is a SLOT of slot_rays consecutive rays of one probe (32 = one
    // warp's fetch = two rows of a 16x16 ray tile, i.e. neighbouring directions; the whole probe
    // when rays/probe is not a multiple of 32).  This shard updates the slots order[0 .. n_owned):
    // its idx-th ray is ray idx % slot_rays of slot order[idx / slot_rays], slot s holding the
    // rays [s * slot_rays, (s + 1) * slot_rays) of the reference's ray list.  The host lists
    // the owned slots most expensive first (ddgi_engine.cu: schedule), so the long rays start
    // early and the persistent kernel's tail is short.  The order never changes a result.
    const uint32_t* order;
    uint32_t n_owned;
    uint32_t slot_rays;
    uint32_t* slot_cost;  // calibration launch: per-slot MAX of the rays' voxel lookups, else nullptr
    int tex_w, tex_h;
    uint32_t* albedo;     // W*H RGBA8
    uint32_t* distance;   // W*H RGBA8 (the reference stores zeros); nullptr = already all zero, skip the stores
    const uint32_t* albedo_old;  // what the hysteresis blend reads: the same plane, or the previous frame's under double buffering
    float4* albedo_f32;   // debug: pre-quantisation values, or nullptr
    uint32_t* lookups;    // debug: per-ray voxel lookups, or nullptr
    int blend;            // 1: blend into the old texel with `hysteresis` (probe_pass.comp:298-299 restored)
    float hysteresis;
    int distance_mode;    // 1: store first-hit distance moments (d, d*d), d = t / distance_scale; 0: zeros as shipped
    float distance_scale;
    float4* ray_out;      // octahedral layout: per-ray (radiance rgb, first-hit t) instead of a texel store, or nullptr
    unsigned long long* warp_times;  // debug level 2: per warp (start, last fetch, exit) globaltimer ns, or nullptr
    int n_peers;          // fused exchange: replicas to store every texel into
    uint32_t* peer_albedo[kMaxPeers];
    uint32_t* peer_distance[kMaxPeers];
};

// Fused exchange: the epoch flags live behind the two texture planes of every replica
// (kFlagWords uint32 after them; slot g = the last epoch rank g has published here).
constexpr int kFlagWords = 64;
struct PeerBarrier {
    int n_ranks, self;
    uint32_t epoch;
    unsigned long long timeout_ns;
    uint32_t* local_flags;
    uint32_t* peer_flags[kMaxPeers];  // by rank; [self] unused
    uint32_t* error;                  // set to 1 when a peer did not arrive in time
};
cudaError_t launch_peer_barrier(const PeerBarrier& B, cudaStream_t s, int* launches);

// Octahedral layout, second kernel of the probe update: ray results -> tile texels.
struct OctJob {
    const uint32_t* probes;  // owned probes
    uint32_t n_probes;
    int n_rays;              // rays per probe
    const float* dirs;       // n_rays normalised directions
    const float4* ray_out;   // all rays of the field, linear ray index
    int tex_w;
    uint32_t* albedo;
    uint32_t* distance;
    const uint32_t* albedo_old;    // blend sources: the same planes, or the previous frame's under double buffering
    const uint32_t* distance_old;
    int blend;
    float hysteresis;
    float distance_scale;
    int n_peers;
    uint32_t* peer_albedo[kMaxPeers];
    uint32_t* peer_distance[kMaxPeers];
};
struct TilePack {
    uint32_t* tex;          // albedo plane; the distance plane follows at `plane` texels
    uint32_t* pack;         // G chunks of chunk_texels
    const uint32_t* owned;  // pack only
    int n_owned;
    int planes;
    size_t plane, chunk_texels;
    int tiles_x, tw, th, tex_w;
    int G, self, B, n_probes, max_owned;
};
cudaError_t launch_pack_tiles(const TilePack& T, bool unpack, cudaStream_t s, int* launches);

cudaError_t launch_probe_blend_octahedral(const FrameParams& P, const OctJob& J, cudaStream_t s, int* launches);

struct PixelJob {
    const uint32_t* albedo;    // probe texture
    const uint32_t* distance;  // distance texture (read only by the restored Chebyshev weight)
    int tex_w;
    int group_row0, group_rows;  // rows of 16x16 workgroups this context renders (all of them by default)
    uint32_t* frame;        // w*h RGBA8
    float4* frame_f32;      // debug or nullptr
    uint32_t* lookups;      // debug or nullptr
};

// grid_limit > 0 caps the resident blocks per SM the persistent kernel launches (tuning)
cudaError_t launch_probe_update(const FrameParams& P, const ProbeJob& J, int variant, uint32_t* counter,
                                int march_min, int grid_limit, cudaStream_t s, int* launches);
uint32_t wavefront_warps(uint32_t n_rays, int grid_limit, int* lanes = nullptr, int* threads = nullptr);
cudaError_t launch_render_frame(const FrameParams& P, const PixelJob& J, cudaStream_t s, int* launches);
cudaError_t launch_bake_scene(int scene, const int dims[3], const int org[3], uint8_t* types,
                              cudaStream_t s, int* launches);
cudaError_t launch_bake_synthetic(const int dims[3], const int org[3], int permille, uint32_t seed,
                                  uint8_t* types, cudaStream_t s, int* launches);
// Rebuilds the occupancy words of the brick box [b0, b0 + bn) from the block types.
cudaError_t launch_build_occupancy(const int dims[3], const int shift[3], const int nb[3], const int b0[3], const int bn[3],
                                   const uint8_t* types, uint32_t* occ, cudaStream_t s, int* launches);
// Copies a box of block types (device staging buffer, x fastest) into the field at grid cell `at`.
cudaError_t launch_edit_voxels(const int dims[3], const int at[3], const int ext[3], const uint8_t* src, uint8_t* types,
                               cudaStream_t s, int* launches);

}  // namespace ddgi
