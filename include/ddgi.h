/*
 * ddgi.h — C ABI of the B200-native DDGI probe-field engine (libddgi_b200.so).
 *
 * This is the drop-in boundary for the hot path of
 * helenl9098/Dynamic-Diffuse-Global-Illumination-Minecraft: the two Vulkan compute
 * dispatches recorded by RVPT::record_compute_command_buffer (src/rvpt/rvpt.cpp:1096-1143)
 * and the per-frame uniform / storage-buffer uploads of RVPT::update
 * (src/rvpt/rvpt.cpp:265-290).  Each entry point cites the reference interface it
 * replaces.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *  - Every function returns DDGI_OK (0) or a negative DDGI_E_* code; it never aborts or
 *    throws (the reference asserts on VkResult < 0, src/rvpt/vk_util.h:18-27).
 *    ddgi_last_error(ctx) gives a human-readable message for the last failure.
 *  - Host pointers are copied during the call; the caller keeps ownership (as
 *    VK::Buffer::copy_to memcpys, src/rvpt/vk_util.cpp:1141-1146).
 *  - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *    ddgi_probe_update / ddgi_render_frame are asynchronous on it; ddgi_sync waits.
 *  - A context is bound to one CUDA device and is not thread-safe (the reference is a
 *    single-threaded caller, src/rvpt/main.cpp:80-96).
 *  - There is no CPU fallback: without a CUDA device ddgi_create fails with
 *    DDGI_E_CUDA.
 */
#ifndef DDGI_H
#define DDGI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DDGI_API __attribute__((visibility("default")))
#else
#define DDGI_API
#endif

#define DDGI_OK 0
#define DDGI_E_INVALID (-1) /* bad argument / call order */
#define DDGI_E_CUDA (-2)    /* CUDA runtime error or no device */
#define DDGI_E_STATE (-3)   /* missing voxels / rays / field before a dispatch */

#define DDGI_MAX_LIGHTS 8

typedef struct ddgi_ctx ddgi_ctx;

/* RVPT::RenderSettings, src/rvpt/rvpt.h:70-80 (32 bytes, uniform binding 0 of both shaders) */
typedef struct ddgi_render_settings {
    int32_t screen_width;
    int32_t screen_height;
    int32_t max_bounces;
    int32_t camera_mode; /* only 0 (pinhole) is on the path */
    int32_t render_mode; /* eval_integrator, compute_pass.comp:58-87: 0 DDGI, 1 direct, 2 indirect,
                            3 colour, 4 normal, 5 depth; any other value = DDGI (the default branch) */
    int32_t scene;       /* 0 cave, 1 Cornell, 2 house: selects the built-in baker / lights */
    float time;
    int32_t visualize_probes; /* != 0: probe markers over render modes 0 and 2 (integrators.glsl:45-67) */
} ddgi_render_settings;

/* RVPT::IrradianceField, src/rvpt/rvpt.h:82-90 (48 bytes, std140) */
typedef struct ddgi_irradiance_field {
    int32_t probe_count[3];      /* @0  */
    int32_t side_length;         /* @12 */
    float hysteresis;            /* @16 plumbed, unused by the reference (probe_pass.comp:298-299) */
    int32_t sqrt_rays_per_probe; /* @20 */
    int32_t _pad0[2];            /* @24 */
    float field_origin[3];       /* @32 */
    int32_t visualize;           /* @44 host-only bool */
} ddgi_irradiance_field;

/* struct ProbeRay, src/rvpt/probe.h:5-20 == assets/shaders/structs.glsl:22-27 (48 bytes) */
typedef struct ddgi_probe_ray {
    float origin[3];
    float _p0;
    float direction[3];
    float _p1;
    float probe_info[3]; /* (probe index, tile x, tile y) stored as floats */
    float _p2;
} ddgi_probe_ray;

/* struct Light, assets/shaders/structs.glsl:54-59 */
typedef struct ddgi_light {
    float intensity;
    float col[3];
    float pos[3];
} ddgi_light;

/* ---- lifetime: RVPT::RVPT + RVPT::initialize (rvpt.cpp:212-264), RVPT::shutdown ---- */
DDGI_API int ddgi_create(ddgi_ctx** out, int device);
DDGI_API void ddgi_destroy(ddgi_ctx* ctx);
DDGI_API const char* ddgi_last_error(const ddgi_ctx* ctx);
/* "major.minor sm_XX" of the build */
DDGI_API const char* ddgi_version(void);

/* ---- per-frame uniforms: RVPT::update, rvpt.cpp:281-287 ---- */
/* settings_uniform.copy_to (rvpt.cpp:282) */
DDGI_API int ddgi_set_render_settings(ddgi_ctx* ctx, const ddgi_render_settings* rs);
/* irradiance_field_uniform.copy_to (rvpt.cpp:287); a change of probe_count or
   sqrt_rays_per_probe re-creates the probe textures as recreate_probe_textures does
   (rvpt.cpp:661-755, 873-890) and invalidates the ray set. */
DDGI_API int ddgi_set_irradiance_field(ddgi_ctx* ctx, const ddgi_irradiance_field* field);
/* Extension: rectangular rx x ry ray tile for non-square rays/probe (128 = 8x16).
   Call after ddgi_set_irradiance_field; rx == ry == sqrt_rays_per_probe is the default. */
DDGI_API int ddgi_set_ray_tile(ddgi_ctx* ctx, int32_t rx, int32_t ry);
/* camera_uniform.copy_to (rvpt.cpp:283): Camera::get_data(), src/rvpt/camera.cpp:100-111 —
   16 floats column-major camera-to-world + (aspect, hfov_rad, scale, 0) */
DDGI_API int ddgi_set_camera(ddgi_ctx* ctx, const float cam[20]);

/* ---- scene: compiled into the reference's GLSL, explicit inputs here ---- */
/* Light table (assets/shaders/structs.glsl:61-89, get_light intersection.glsl:1131-1150). */
DDGI_API int ddgi_set_lights(ddgi_ctx* ctx, int32_t n, const ddgi_light* lights);
/* The reference's table for `scene` (n = 1,1,2) */
DDGI_API int ddgi_default_lights(int32_t scene, ddgi_light* out, int32_t* n);
/* update_lights (probe_pass.comp:217-250 == compute_pass.comp:126-160): the light animation whose
   call the reference has commented out in both main()s (probe_pass.comp:254, compute_pass.comp:174),
   as a function of render_settings.time applied to a table of n lights of `scene` (0 cave, 1 Cornell,
   2 house).  Pass the result to ddgi_set_lights each frame for dynamic lights. */
DDGI_API int ddgi_update_lights(int32_t scene, float time, const ddgi_light* base, int32_t n, ddgi_light* out);
/* The commented 4-light cave table (structs.glsl:65-69) moved by update_lights' cave
   branch (probe_pass.comp:219-235) for `time`. */
DDGI_API int ddgi_cave_lights4(float time, ddgi_light* out);
/* Stored voxel field replacing getBlockAt (intersection.glsl:699-826): `types` is
   dims[0]*dims[1]*dims[2] block types, x fastest; cell (0,0,0) is voxel id `origin`
   (voxel id c covers (c-1, c] per axis).  palette: 256 rgb triples or NULL for the
   flat-colour table of the reference's block types. */
DDGI_API int ddgi_upload_voxels(ddgi_ctx* ctx, const int32_t dims[3], const int32_t origin[3],
                       const uint8_t* types, const float* palette);
/* Built-in baker: evaluates the reference's procedural block function for `scene` on
   the device over the given box. */
DDGI_API int ddgi_bake_scene(ddgi_ctx* ctx, int32_t scene, const int32_t dims[3], const int32_t origin[3]);
/* Synthetic cave-like field for the 32^3-probe benchmark (SURVEY.md 8d cfg 4): the
   reference cave's four-sphere cavity scaled to the grid, `solid_permille` of the cavity
   cells filled at random (xorshift32, seed), six flat block types. */
DDGI_API int ddgi_bake_synthetic(ddgi_ctx* ctx, const int32_t dims[3], const int32_t origin[3],
                        int32_t solid_permille, uint32_t seed);
/* Albedo of a voxel hit (getColorAt, intersection.glsl:872-1047).
   DDGI_COLOR_PALETTE (default): the flat colour of the block type from the palette — exactly
   the reference for its flat types 2-5 (Cornell), the README's "flat colors" variant for the
   cave.  DDGI_COLOR_LITERAL: the reference's procedural textures (worley / fbm / dots / hash of
   the continuous hit point), evaluated on the device; with the scene baked by ddgi_bake_scene
   over a box that holds every probe this reproduces the reference's textured cave. */
#define DDGI_COLOR_PALETTE 0
#define DDGI_COLOR_LITERAL 1
DDGI_API int ddgi_set_color_mode(ddgi_ctx* ctx, int32_t mode);
/* How a new probe-ray colour meets the old texel.  DDGI_BLEND_OVERWRITE (default) is the
   reference as shipped: the texel is overwritten every frame.  DDGI_BLEND_HYSTERESIS restores
   the blend the reference has commented out (probe_pass.comp:298-299):
   color = mix(imageLoad(albedo, texel).rgb, color, irradiance_field.hysteresis) — note that
   the reference weights the NEW colour by `hysteresis`.  The texture then carries state from
   frame to frame (ddgi_write_probe_texture / ddgi_read_probe_texture checkpoint it). */
#define DDGI_BLEND_OVERWRITE 0
#define DDGI_BLEND_HYSTERESIS 1
DDGI_API int ddgi_set_blend_mode(ddgi_ctx* ctx, int32_t mode);
/* Cage-sample weight.  DDGI_WEIGHT_LITERAL (default) is the reference as shipped: the Chebyshev
   visibility term is computed and discarded (intersection.glsl:1367-1383).  DDGI_WEIGHT_CHEBYSHEV
   restores the commented-out `weight *= chebyshevWeight;` (:1382), reading mean / mean^2 from the
   distance texture exactly as that code does (sample_probe(.., 1).rg). */
#define DDGI_WEIGHT_LITERAL 0
#define DDGI_WEIGHT_CHEBYSHEV 1
DDGI_API int ddgi_set_weight_mode(ddgi_ctx* ctx, int32_t mode);
/* Distance texture.  DDGI_DISTANCE_ZERO (default) is the reference as shipped: `distances =
   vec2(0)` (probe_pass.comp:276,302).  DDGI_DISTANCE_MOMENTS (extension; the reference has no
   code for it) stores (d, d*d) with d = t / scale of the probe ray's first intersect_scene (INF on
   a miss: saturates to 1 in the UNORM store); the Chebyshev weight then compares
   length(pos - probe_pos) / scale with it.  scale = 1 leaves the restored reference text
   unchanged (x / 1 is exact); scale ~ the probe spacing makes the RGBA8 moments useful. */
#define DDGI_DISTANCE_ZERO 0
#define DDGI_DISTANCE_MOMENTS 1
DDGI_API int ddgi_set_distance_mode(ddgi_ctx* ctx, int32_t mode, float scale);
/* Per-frame scene edit (dynamic scenes; the reference's scene is compiled into its shaders and
   cannot change): overwrites the box of voxel ids [origin, origin + dims) with `types` (x fastest)
   and rebuilds only the occupancy bricks it touches.  The box must lie inside the uploaded field.
   Ordered on `stream` with the dispatches and asynchronous: pageable `types` may be reused on return,
   pinned `types` once `stream` has passed the copy.  The cost-ordered schedule keeps its last
   calibration (results never depend on it). */
DDGI_API int ddgi_edit_voxels(ddgi_ctx* ctx, const int32_t origin[3], const int32_t dims[3], const uint8_t* types, void* stream);
/* Probe-texture layout.  DDGI_LAYOUT_RAY_TILE (default) is the reference: one texel per ray, tile
   rx x ry (probe_pass.comp:269-271).  DDGI_LAYOUT_OCTAHEDRAL is the textbook layout the north star
   names: an oct x oct octahedral tile per probe (mapping = the reference's unused
   assets/shaders/octahedral.glsl:16-35); texel = cosine-weighted mean over ALL the probe's rays of
   the ray radiance (albedo plane) and of the first-hit distance moments (d, d^2), d = min(t /
   distance_scale, 1) (distance plane), reduced with warp shuffles, blended by the hysteresis rule
   when DDGI_BLEND_HYSTERESIS is on; the pixel pass fetches tiles bilinearly at octEncode(dir).
   The reference has no code that fills or reads such a tile, so this mode has no reference
   output: parity for it is engine vs oracle only.  Re-creates the probe textures
   (W = X*Z*oct, H = Y*oct); needs a generated ray set. */
#define DDGI_LAYOUT_RAY_TILE 0
#define DDGI_LAYOUT_OCTAHEDRAL 1
DDGI_API int ddgi_set_layout(ddgi_ctx* ctx, int32_t layout, int32_t oct);
/* Copies the block types back (dims product bytes). */
DDGI_API int ddgi_read_voxels(ddgi_ctx* ctx, uint8_t* dst, size_t bytes);

/* ---- probe rays: RVPT::generate_probe_rays (rvpt.h:63, rvpt.cpp:1177-1224) ---- */
/* Generates the stratified sample set with libc rand() as generate_samples (rvpt.cpp:1147-1173)
   and keeps only the per-probe direction table on the device; the kernel derives origin / direction /
   tile offset of ray k itself (no 48 B/ray read).  reseed != 0 calls srand(1) first (the state of
   a fresh process).  The reference draws a sample's two jitters as constructor arguments
   (rvpt.cpp:1161-1162), whose evaluation order C++ leaves open: ddgi_set_sample_order(ctx, 0)
   (default) draws x then y (SURVEY.md 8c-5's pin, what the committed fixtures hold),
   ddgi_set_sample_order(ctx, 1) y then x - what g++, the reference's Linux toolchain, compiles that
   line to; with it the ray list equals the one a g++ build of the reference generates, bit for bit. */
DDGI_API int ddgi_set_sample_order(ddgi_ctx* ctx, int32_t y_first);
DDGI_API int ddgi_generate_probe_rays(ddgi_ctx* ctx, int32_t reseed);
/* Spherical-Fibonacci set of rx*ry directions (north-star ray generator; no reference counterpart —
   the reference draws the stratified rand() set above): cos(theta_i) = 1 - (2i+1)/n, azimuth
   2 pi frac(i (phi - 1)).  Deterministic, no rand(). */
DDGI_API int ddgi_generate_fibonacci_rays(ddgi_ctx* ctx);
/* Same, from a caller-supplied table of rx*ry un-normalised sphere samples (xyz). */
DDGI_API int ddgi_set_ray_samples(ddgi_ctx* ctx, const float* samples, size_t count);
/* Reads the current rx*ry sample table back (3 floats each). */
DDGI_API int ddgi_get_ray_samples(ddgi_ctx* ctx, float* dst, size_t count);
/* probe_buffer.copy_to(probe_rays) (rvpt.cpp:285): literal storage-buffer mode, the
   kernel reads the caller's 48-byte records. */
DDGI_API int ddgi_set_probe_rays(ddgi_ctx* ctx, const ddgi_probe_ray* rays, size_t count);
/* Writes the full ray list (what the reference's std::vector<ProbeRay> holds). */
DDGI_API int ddgi_get_probe_rays(ddgi_ctx* ctx, ddgi_probe_ray* dst, size_t count);
DDGI_API size_t ddgi_num_probe_rays(const ddgi_ctx* ctx);

/* ---- probe sharding across GPUs (no reference counterpart) ---- */
/* This context updates only probe rows [y0, y1) (texture rows [y0*ry, y1*ry)). */
DDGI_API int ddgi_set_probe_rows(ddgi_ctx* ctx, int32_t y0, int32_t y1);
/* Block-cyclic ownership instead: this context updates the probe rows y with
   (y / block) % world == rank.  Spreads the expensive (open-cavity) rows over all ranks;
   each owned block is still a contiguous byte range of the texture. */
DDGI_API int ddgi_set_probe_rows_cyclic(ddgi_ctx* ctx, int32_t rank, int32_t world, int32_t block);
/* The same at probe granularity: this context updates the probes p with
   (p / block) % world == rank.  Finest balance; a rank's texels are then scattered tiles,
   so pair it with the fused exchange (ddgi_open_peers), not an all-gather. */
DDGI_API int ddgi_set_probes_cyclic(ddgi_ctx* ctx, int32_t rank, int32_t world, int32_t block);
/* Device address and size of probe texture `which` (0 albedo, 1 distance); both live in
   one allocation, albedo first, so one collective can move both. */
DDGI_API int ddgi_probe_texture_device_ptr(ddgi_ctx* ctx, int32_t which, void** ptr, size_t* bytes);
/* Fused exchange: 64-byte CUDA IPC handles of the texture allocation(s) / open the peers'.
   After ddgi_open_peers, ddgi_probe_update stores every texel into all replicas.
   ddgi_export_texture_handles writes *count = 1 handle, or 2 for a double-buffered context
   (ddgi_set_double_buffer before exporting; ddgi_export_texture_handle is the single-buffer form);
   ddgi_open_peers takes, rank by rank, as many handles per rank as this context exported - every rank
   must be configured alike.  Double-buffered replicas are what lets ddgi_read_probe_texture_async
   overlap the next update under the fused exchange: update i stores into allocation i & 1 of every
   rank while frame i-1 is still being rendered or read from the other one. */
DDGI_API int ddgi_export_texture_handle(ddgi_ctx* ctx, void* handle64);
DDGI_API int ddgi_export_texture_handles(ddgi_ctx* ctx, void* handles, int32_t* count);
DDGI_API int ddgi_open_peers(ddgi_ctx* ctx, int32_t n_peers, const void* handles64, int32_t self_index);
DDGI_API int ddgi_close_peers(ddgi_ctx* ctx);
/* Completion barrier of the fused exchange, on `stream` after ddgi_probe_update: publishes this
   rank's frame epoch in every peer's replica and waits (on the device) until every peer's epoch
   has arrived here, i.e. until their texels are in the local replica.  One small kernel, no
   collective library.  Every rank must call it once per frame.  A peer that does not arrive
   within 5 s ends the wait (never a hang); ddgi_exchange_status reports it (synchronises).
   Frames never mix: with single-buffered replicas ddgi_probe_update itself issues a second barrier
   in front of the kernel, so no rank stores frame i+1 into a replica whose owner has not yet issued
   its own update i+1 - which its ddgi_render_frame / reads of frame i precede in stream order; with
   double-buffered replicas frame i+1 goes to the other allocation and that barrier is not needed.
   Either way render and read frame i on the SAME stream as the update, or through
   ddgi_read_probe_texture_async (this barrier waits for the last asynchronous read of the
   allocation the next update will write). */
DDGI_API int ddgi_exchange_barrier(ddgi_ctx* ctx, void* stream);
DDGI_API int ddgi_exchange_status(ddgi_ctx* ctx);

/* The same exchange over NCCL (the collective SURVEY.md 8e names; what a deployment without CUDA IPC
   between its processes uses).  libnccl.so.2 is bound at run time (dlopen), a single-GPU host does
   not need it.  ddgi_comm_unique_id fills a 128-byte ncclUniqueId on one rank; the host passes it to
   the others by its own means (MPI, a file, torch.distributed) and every rank calls ddgi_comm_init.
   ddgi_exchange_allgather, on `stream` after ddgi_probe_update, moves the probe rows in place:
   ONE ncclAllGather per plane for equal contiguous slabs (ddgi_set_probe_rows(rank*Y/world,
   (rank+1)*Y/world), Y % world == 0: the send buffer is the rank's slab inside the receive buffer),
   one group of in-place ncclBroadcasts for block-cyclic rows (ddgi_set_probe_rows_cyclic), or - probe-cyclic
   ownership (ddgi_set_probes_cyclic: the balanced one, scattered tiles) - the rank's tiles packed into one
   chunk, ONE ncclAllGather of the chunks, and a scatter of the other ranks' tiles (two small kernels around
   the collective).  The distance plane travels only once something other than the reference's zeros has
   been stored in it.  The ownership must match the communicator's rank and size: DDGI_E_STATE otherwise. */
DDGI_API int ddgi_comm_unique_id(void* id128);
DDGI_API int ddgi_comm_init(ddgi_ctx* ctx, const void* id128, int32_t rank, int32_t world);
DDGI_API int ddgi_comm_destroy(ddgi_ctx* ctx);
DDGI_API int ddgi_exchange_allgather(ddgi_ctx* ctx, void* stream);

/* Pixel pass across GPUs: every replica holds the whole probe texture after the exchange, so the
   frame splits into `world` bands of 16-pixel workgroup rows with no further collective; this
   context's ddgi_render_frame writes only band `rank` (pixel rows [y0, y1) of
   ddgi_frame_band_rows) of its frame buffer.  Default (0, 1): the whole frame. */
DDGI_API int ddgi_set_frame_band(ddgi_ctx* ctx, int32_t rank, int32_t world);
DDGI_API int ddgi_frame_band_rows(const ddgi_ctx* ctx, int32_t* y0, int32_t* y1);

/* ---- the two dispatches ---- */
/* vkCmdDispatch #1, probe_pass.comp (rvpt.cpp:1121-1129) */
DDGI_API int ddgi_probe_update(ddgi_ctx* ctx, void* stream);
/* vkCmdDispatch #2, compute_pass.comp (rvpt.cpp:1133-1140) */
DDGI_API int ddgi_render_frame(ddgi_ctx* ctx, void* stream);
/* raytrace_work_fence.wait (rvpt.cpp:277): waits for the stream of this context's last dispatch and
   for its asynchronous reads - not for the device (the caller's other streams keep running) */
DDGI_API int ddgi_sync(ddgi_ctx* ctx);

/* ---- results ---- */
#define DDGI_FMT_RGBA8 0 /* what the reference stores (VK_FORMAT_R8G8B8A8_UNORM) */
#define DDGI_FMT_F32 1   /* the fp32 value before the UNORM store, 4 floats per texel;
                            only kept when ddgi_set_debug(ctx, 1) */
DDGI_API int ddgi_probe_texture_size(const ddgi_ctx* ctx, int32_t* width, int32_t* height);
DDGI_API int ddgi_read_probe_texture(ddgi_ctx* ctx, int32_t which, int32_t fmt, void* dst, size_t bytes);
/* Double buffering (default off): probe updates alternate between two texture allocations, the
   hysteresis blend reads the previous frame's, ddgi_render_frame and the reads use the latest.
   With it frame i can be copied to the host while frame i+1 is traced:
   ddgi_read_probe_texture_async enqueues the copy of the latest frame on the engine's own copy
   stream, ordered after the update that produced it (dst should be pinned; it is valid after
   ddgi_read_wait); an update that is about to overwrite a buffer first waits, on the device, for
   the last asynchronous read of it (the next update but one; without double buffering the very
   next update, so the asynchronous read is then correct but overlaps nothing).  Under the fused
   exchange turn it on BEFORE exporting the handles (ddgi_export_texture_handles lists both
   allocations).  With partial probe ownership and no exchange the texels a context does not own
   are one frame older in every other buffer. */
DDGI_API int ddgi_set_double_buffer(ddgi_ctx* ctx, int32_t on);
/* Frames in flight: 1 (default) - every dispatch runs on the caller's stream.  2 (needs the double-buffered
   texture) - the reference keeps two frames in flight as well (MAX_FRAMES_IN_FLIGHT, src/rvpt/rvpt.h:23): the
   probe update of frame i, and the exchange that completes it, run on the engine's own stream i & 1, ordered
   behind everything the caller's stream held when ddgi_probe_update was called.  Consecutive updates write
   different allocations and do not depend on each other (unless the hysteresis blend is on: then they
   serialise), so the first blocks of update i+1 fill the SMs the persistent kernel of update i leaves idle
   while it drains - what separates 8 GPUs from 8x on a 1/8 share.  ddgi_render_frame and the reads wait for the
   frame they use; ddgi_frame_fence makes a stream of the caller's wait for every frame in flight (e.g. before an
   event that times a run); under the fused exchange every update is preceded by a barrier that keeps a rank
   from storing frame i+2 into a replica whose owner still reads frame i (see ddgi_exchange_barrier). */
DDGI_API int ddgi_set_frames_in_flight(ddgi_ctx* ctx, int32_t n);
DDGI_API int ddgi_frame_fence(ddgi_ctx* ctx, void* stream);
/* Device time of the latest probe-update kernel (CUDA events on the stream it ran on); synchronises with it. */
DDGI_API int ddgi_last_update_ms(ddgi_ctx* ctx, float* ms);
DDGI_API int ddgi_read_probe_texture_async(ddgi_ctx* ctx, int32_t which, void* dst, size_t bytes);
/* The same for texture rows [row0, row1) only (a rank's share of the replica it holds after an exchange);
   ordered after everything given to the dispatch stream so far, i.e. also after ddgi_exchange_barrier /
   ddgi_exchange_allgather of this frame. */
DDGI_API int ddgi_read_probe_texture_rows_async(ddgi_ctx* ctx, int32_t which, int32_t row0, int32_t row1, void* dst, size_t bytes);
DDGI_API int ddgi_read_wait(ddgi_ctx* ctx);
/* Uploads texture contents (checkpoint / resume, and pixel-pass tests). */
DDGI_API int ddgi_write_probe_texture(ddgi_ctx* ctx, int32_t which, const void* src, size_t bytes);
/* On-disk formats (the reference has none: its scene is compiled in and its texture
   recomputed every frame).  Little-endian.
   Voxel file: "DDGIVOX1", int32 dims[3] (x, y, z), int32 origin[3], dims product block types, x fastest.
   ddgi_load_voxels uploads it as ddgi_upload_voxels would (palette may be NULL) and reports dims / origin
   (either may be NULL).
   Checkpoint: "DDGIPTX1", int32 W, int32 H, float time, albedo plane, distance plane (RGBA8 rows): the
   state a hysteresis-blended texture carries from frame to frame, with the caller's render_settings.time.
   ddgi_load_checkpoint needs the field of the dumped shape to be set already. */
DDGI_API int ddgi_save_voxels(ddgi_ctx* ctx, const char* path);
DDGI_API int ddgi_load_voxels(ddgi_ctx* ctx, const char* path, const float* palette, int32_t dims_out[3], int32_t origin_out[3]);
DDGI_API int ddgi_save_checkpoint(ddgi_ctx* ctx, const char* path, float time);
DDGI_API int ddgi_load_checkpoint(ddgi_ctx* ctx, const char* path, float* time_out);
DDGI_API int ddgi_read_frame(ddgi_ctx* ctx, int32_t fmt, void* dst, size_t bytes);

/* ---- instrumentation ---- */
/* debug >= 1: keep fp32 copies of the outputs and per-invocation voxel-lookup counts.
   debug == 2: kernel variants 1 and 2 also record, per warp, the %globaltimer (ns) at which it started,
   took its last ray and exited (ddgi_read_warp_times) — how long the persistent kernel's tail is. */
DDGI_API int ddgi_set_debug(ddgi_ctx* ctx, int32_t debug);
/* *n_warps = warps of the last probe update that recorded times; dst (may be NULL to query the
   count) receives 3 values per warp. */
DDGI_API int ddgi_read_warp_times(ddgi_ctx* ctx, uint64_t* dst, size_t count, size_t* n_warps);
/* Per-ray (which = 0) or per-pixel (which = 1) voxel lookups of the last dispatch. */
DDGI_API int ddgi_read_lookup_counts(ddgi_ctx* ctx, int32_t which, uint32_t* dst, size_t count);
/* Kernel variant: 0 = one thread per ray, reference loop order; 1 = regrouped state-machine
   kernel performing exactly the voxel lookups of the reference algorithm; 2 (default) = the same
   kernel with its result-preserving early-outs: the march of a shadow feeler ends once the feeler
   has left its light behind (the reference marches on to the next block or 125 cells and then
   discards that hit, probe_pass.comp:186-207 with intersection.glsl:1262-1295).  Texels are
   identical bit for bit in all three; the per-ray lookup counts of ddgi_read_lookup_counts are the
   reference algorithm's in variants 0 and 1 and at most those in variant 2. */
DDGI_API int ddgi_set_kernel_variant(ddgi_ctx* ctx, int32_t variant);
/* Scheduling knob of variant 1: a warp keeps stepping its marches while at least
   march_min/32 of the lanes that hold a ray are marching (1..32, default 16).  Results do
   not depend on it. */
DDGI_API int ddgi_set_tuning(ddgi_ctx* ctx, int32_t march_min);
/* Granularity of the cost-ordered schedule: slots of `rays` consecutive rays of a probe (any count
   that divides rays/probe; default 32 = one warp's fetch, two rows of a 16x16 ray tile; 1 = every
   ray ranked on its own), 0 = whole probes.  Results do not depend on it. */
DDGI_API int ddgi_set_schedule_slot(ddgi_ctx* ctx, int32_t rays);
/* Caps the resident blocks per SM variant 1 launches (0 = as many as fit, the default).  Results
   do not depend on it. */
DDGI_API int ddgi_set_grid_limit(ddgi_ctx* ctx, int32_t blocks_per_sm);
/* Cost-ordered scheduling (default on): the first ddgi_probe_update after the voxels, the
   field shape or the probe ownership changed also sums the voxel lookups per probe and
   synchronises once to read them; later updates trace the owned probes most expensive
   first, which shortens the tail of the persistent kernel.  Results do not depend on it.
   0 = natural probe order, never synchronises. */
DDGI_API int ddgi_set_auto_schedule(ddgi_ctx* ctx, int32_t on);
/* Number of kernel launches issued by this context so far. */
DDGI_API uint64_t ddgi_launch_count(const ddgi_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* DDGI_H */
