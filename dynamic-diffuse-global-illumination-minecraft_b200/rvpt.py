"""Host-side mirror of the reference's ``class RVPT`` for the DDGI hot path.

Same member names and call order as src/rvpt/rvpt.h:33-92 / src/rvpt/main.cpp:80-96:

    rvpt = RVPT(width, height)           # RVPT::RVPT + initialize()
    rvpt.ir.probe_count = (9, 7, 9)      # public POD members
    rvpt.generate_probe_rays()           # rvpt.cpp:1177
    rvpt.update(); rvpt.draw()           # per frame; rvpt.cpp:265, :372

Everything below is a thin veneer over the C-ABI (capi.py / include/ddgi.h); all
computation happens in libddgi_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import capi


class DDGIError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ddgi error {code}: {msg}")
        self.code = code


class Camera:
    """src/rvpt/camera.{h,cpp}: FPS camera; get_data() is the 80-byte camera block."""

    def __init__(self, aspect: float, origin=(1.5, 2.0, -2.0), rotation=(-38.0, 36.0, 0.0)):
        self.aspect = float(aspect)
        self.translation = np.asarray(origin, dtype=np.float32)
        self.rotation = np.asarray(rotation, dtype=np.float32)
        self.fov = 75.0
        self.scale = 4.0
        self.mode = 0

    @staticmethod
    def _rot(axis: int, deg: float) -> np.ndarray:
        a = math.radians(float(np.float32(deg)))
        c, s = math.cos(a), math.sin(a)
        m = np.eye(4, dtype=np.float64)
        i, j = [(1, 2), (2, 0), (0, 1)][axis]
        m[i, i] = c
        m[j, j] = c
        m[i, j] = -s
        m[j, i] = s
        return m

    def camera_matrix(self) -> np.ndarray:
        """construct_camera_matrix, camera.cpp:18-26: T * R_y(rot.x) * R_x(rot.y) * R_z(rot.z)."""
        t = np.eye(4, dtype=np.float64)
        t[:3, 3] = self.translation
        m = t @ self._rot(1, self.rotation[0]) @ self._rot(0, self.rotation[1]) @ self._rot(2, self.rotation[2])
        return m.astype(np.float32)

    def get_data(self) -> np.ndarray:
        """camera.cpp:100-111: 4 matrix columns + (aspect, radians(fov), scale, 0)."""
        m = self.camera_matrix()
        out = np.zeros(20, dtype=np.float32)
        out[:16] = m.T.reshape(-1)  # column major
        out[16] = np.float32(self.aspect)
        out[17] = np.float32(math.radians(self.fov))
        out[18] = np.float32(self.scale)
        return out


class RVPT:
    """The probe-field part of the reference's RVPT class, backed by the CUDA engine."""

    def __init__(self, width: int = 1600, height: int = 900, device: int = 0):
        self._lib = capi.load()
        self._ctx = C.c_void_p()
        rc = self._lib.ddgi_create(C.byref(self._ctx), device)
        if rc != capi.OK:
            self._ctx = C.c_void_p()
            raise DDGIError(rc, "ddgi_create failed (no sm_100 CUDA device? the engine has no CPU fallback)")
        self.device = device
        # RVPT::RenderSettings defaults, rvpt.h:70-80
        self.render_settings = capi.RenderSettings(width, height, 8, 0, 0, 0, 0.0, 0)
        # RVPT::IrradianceField defaults, rvpt.h:82-90
        self.ir = capi.IrradianceField((9, 7, 9), 11, 0.9, 20, (0, 0), (1.4, 0.0, 1.0), 1)
        self.scene_camera = Camera(width / float(height))
        self.ray_tile = None  # (rx, ry) extension; None = square sqrt_rays_per_probe
        self.stream = None    # cudaStream_t as int, None = default stream
        self.lights = None    # None = the reference's table for render_settings.scene
        self.animate_lights = False  # True = update_lights() restored: lights move with render_settings.time
        self._applied_field = None
        self.kernel_variant = 2  # the engine's default (ddgi_set_kernel_variant)

    # -- plumbing -------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != capi.OK:
            raise DDGIError(rc, self._lib.ddgi_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.ddgi_destroy(self._ctx)
            self._ctx = C.c_void_p()

    shutdown = close  # RVPT::shutdown

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _apply_field(self):
        self._check(self._lib.ddgi_set_irradiance_field(self._ctx, C.byref(self.ir)))
        if self.ray_tile is not None:
            self._check(self._lib.ddgi_set_ray_tile(self._ctx, int(self.ray_tile[0]), int(self.ray_tile[1])))

    def _apply_lights(self):
        arr = (capi.Light * capi.MAX_LIGHTS)()
        n = C.c_int32()
        if self.lights is None:
            self._check(self._lib.ddgi_default_lights(self.render_settings.scene, arr, C.byref(n)))
        else:
            n.value = len(self.lights)
            for i, l in enumerate(self.lights):
                arr[i] = l
        if self.animate_lights:
            # the `update_lights();` call both shaders have commented out (probe_pass.comp:254,
            # compute_pass.comp:174): positions as a function of render_settings.time
            moved = (capi.Light * capi.MAX_LIGHTS)()
            self._check(self._lib.ddgi_update_lights(self.render_settings.scene, self.render_settings.time, arr, n.value, moved))
            arr = moved
        self._check(self._lib.ddgi_set_lights(self._ctx, n.value, arr))

    # -- scene ----------------------------------------------------------------------
    def bake_scene(self, dims, origin, scene=None):
        """Bakes the reference's procedural block function (intersection.glsl:699-826)."""
        scene = self.render_settings.scene if scene is None else scene
        d = (C.c_int32 * 3)(*dims)
        o = (C.c_int32 * 3)(*origin)
        self._check(self._lib.ddgi_bake_scene(self._ctx, scene, d, o))

    def bake_synthetic(self, dims, origin, solid_permille=50, seed=0x9E3779B9):
        d = (C.c_int32 * 3)(*dims)
        o = (C.c_int32 * 3)(*origin)
        self._check(self._lib.ddgi_bake_synthetic(self._ctx, d, o, solid_permille, seed))

    def upload_voxels(self, types: np.ndarray, origin, palette: np.ndarray | None = None):
        """types: uint8 array indexed [z, y, x]."""
        types = np.ascontiguousarray(types, dtype=np.uint8)
        dz, dy, dx = types.shape
        d = (C.c_int32 * 3)(dx, dy, dz)
        o = (C.c_int32 * 3)(*origin)
        pal = None
        if palette is not None:
            palette = np.ascontiguousarray(palette, dtype=np.float32)
            assert palette.size == 256 * 3
            pal = palette.ctypes.data
        self._check(self._lib.ddgi_upload_voxels(self._ctx, d, o, types.ctypes.data, pal))

    def edit_voxels(self, types: np.ndarray, origin):
        """Overwrites the box of voxel ids starting at `origin` with `types` ([z, y, x] uint8) and
        rebuilds the occupancy bricks it touches (per-frame scene edits)."""
        types = np.ascontiguousarray(types, dtype=np.uint8)
        dz, dy, dx = types.shape
        d = (C.c_int32 * 3)(dx, dy, dz)
        o = (C.c_int32 * 3)(*origin)
        self._check(self._lib.ddgi_edit_voxels(self._ctx, o, d, types.ctypes.data, self.stream))

    def save_voxels(self, path: str, dims=None, origin=None):
        """Raw voxel file (ddgi_save_voxels): magic, dims (x, y, z), origin, then dims product block types, x fastest."""
        self._check(self._lib.ddgi_save_voxels(self._ctx, str(path).encode()))

    def load_voxels(self, path: str, palette: np.ndarray | None = None):
        """Uploads a file written by save_voxels (ddgi_load_voxels); returns (dims, origin)."""
        d, o = (C.c_int32 * 3)(), (C.c_int32 * 3)()
        pal = None if palette is None else np.ascontiguousarray(palette, dtype=np.float32)
        self._check(self._lib.ddgi_load_voxels(self._ctx, str(path).encode(), None if pal is None else pal.ctypes.data, d, o))
        return tuple(d), tuple(o)

    def set_color_mode(self, mode: int):
        """capi.COLOR_PALETTE (flat colours) or capi.COLOR_LITERAL (the reference's procedural textures)."""
        self._check(self._lib.ddgi_set_color_mode(self._ctx, mode))

    def set_blend_mode(self, mode: int):
        """capi.BLEND_OVERWRITE (the reference as shipped) or capi.BLEND_HYSTERESIS (probe_pass.comp:298-299 restored)."""
        self._check(self._lib.ddgi_set_blend_mode(self._ctx, mode))

    def set_weight_mode(self, mode: int):
        """capi.WEIGHT_LITERAL (as shipped) or capi.WEIGHT_CHEBYSHEV (intersection.glsl:1382 restored)."""
        self._check(self._lib.ddgi_set_weight_mode(self._ctx, mode))

    def set_distance_mode(self, mode: int, scale: float = 1.0):
        """capi.DISTANCE_ZERO (as shipped) or capi.DISTANCE_MOMENTS (first-hit (d, d^2) / scale)."""
        self._check(self._lib.ddgi_set_distance_mode(self._ctx, mode, float(scale)))

    def read_voxels(self, dims) -> np.ndarray:
        out = np.empty((dims[2], dims[1], dims[0]), dtype=np.uint8)
        self._check(self._lib.ddgi_read_voxels(self._ctx, out.ctypes.data, out.nbytes))
        return out

    # -- rays -----------------------------------------------------------------------
    def generate_probe_rays(self, reseed: bool = True):
        """RVPT::generate_probe_rays (rvpt.cpp:1177-1224)."""
        self._apply_field()
        self._check(self._lib.ddgi_generate_probe_rays(self._ctx, 1 if reseed else 0))

    def generate_fibonacci_rays(self):
        """Spherical-Fibonacci ray set (north-star generator; deterministic)."""
        self._apply_field()
        self._check(self._lib.ddgi_generate_fibonacci_rays(self._ctx))

    def set_layout(self, layout: int, oct: int = 8):
        """capi.LAYOUT_RAY_TILE (the reference) or capi.LAYOUT_OCTAHEDRAL (oct x oct tile per probe)."""
        self._apply_field()
        self._check(self._lib.ddgi_set_layout(self._ctx, layout, oct))

    def set_ray_samples(self, samples: np.ndarray):
        self._apply_field()
        s = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1, 3)
        self._check(self._lib.ddgi_set_ray_samples(self._ctx, s.ctypes.data, s.shape[0]))

    @property
    def ray_samples(self) -> np.ndarray:
        n = self.num_probe_rays // (self.ir.probe_count[0] * self.ir.probe_count[1] * self.ir.probe_count[2])
        out = np.empty((n, 3), dtype=np.float32)
        self._check(self._lib.ddgi_get_ray_samples(self._ctx, out.ctypes.data, n))
        return out

    def set_probe_rays(self, rays: np.ndarray):
        """probe_buffer.copy_to(probe_rays), rvpt.cpp:285.  rays: float32 [R, 12]."""
        self._apply_field()
        r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 12)
        self._check(self._lib.ddgi_set_probe_rays(self._ctx, r.ctypes.data, r.shape[0]))

    @property
    def num_probe_rays(self) -> int:
        return int(self._lib.ddgi_num_probe_rays(self._ctx))

    @property
    def probe_rays(self) -> np.ndarray:
        n = self.num_probe_rays
        out = np.empty((n, 12), dtype=np.float32)
        self._check(self._lib.ddgi_get_probe_rays(self._ctx, out.ctypes.data, n))
        return out

    # -- per frame ------------------------------------------------------------------
    def update(self, advance_time: bool = True):
        """RVPT::update (rvpt.cpp:265-290): time += 2, upload the uniforms."""
        if advance_time:
            self.render_settings.time += 2
        self._check(self._lib.ddgi_set_render_settings(self._ctx, C.byref(self.render_settings)))
        cam = self.scene_camera.get_data()
        self._check(self._lib.ddgi_set_camera(self._ctx, cam.ctypes.data_as(C.POINTER(C.c_float))))
        self._apply_field()
        self._apply_lights()
        return True

    def probe_update(self):
        """Dispatch #1 (rvpt.cpp:1121-1129)."""
        self._check(self._lib.ddgi_probe_update(self._ctx, self.stream))

    def render_frame(self):
        """Dispatch #2 (rvpt.cpp:1133-1140)."""
        self._check(self._lib.ddgi_render_frame(self._ctx, self.stream))

    def draw(self):
        """RVPT::draw -> record_compute_command_buffer (rvpt.cpp:1096-1143): both dispatches."""
        self.probe_update()
        self.render_frame()

    def sync(self):
        self._check(self._lib.ddgi_sync(self._ctx))

    # -- results --------------------------------------------------------------------
    @property
    def probe_texture_size(self):
        w, h = C.c_int32(), C.c_int32()
        self._check(self._lib.ddgi_probe_texture_size(self._ctx, C.byref(w), C.byref(h)))
        return w.value, h.value

    def read_probe_texture(self, which: int = 0, fmt: int = capi.FMT_RGBA8, out: np.ndarray | None = None):
        w, h = self.probe_texture_size
        if fmt == capi.FMT_RGBA8:
            if out is None:
                out = np.empty((h, w), dtype=np.uint32)
        else:
            if out is None:
                out = np.empty((h, w, 4), dtype=np.float32)
        self._check(self._lib.ddgi_read_probe_texture(self._ctx, which, fmt, out.ctypes.data, out.nbytes))
        return out

    def set_double_buffer(self, on: bool):
        self._check(self._lib.ddgi_set_double_buffer(self._ctx, 1 if on else 0))

    def set_frames_in_flight(self, n: int):
        """2: updates run on the engine's own streams so that consecutive frames overlap (needs set_double_buffer)."""
        self._check(self._lib.ddgi_set_frames_in_flight(self._ctx, n))

    def frame_fence(self):
        """Makes self.stream wait for every frame in flight."""
        self._check(self._lib.ddgi_frame_fence(self._ctx, self.stream))

    def last_update_ms(self) -> float:
        ms = C.c_float()
        self._check(self._lib.ddgi_last_update_ms(self._ctx, C.byref(ms)))
        return ms.value

    def read_probe_texture_async(self, out_ptr: int, nbytes: int, which: int = 0):
        """Enqueues the copy of the latest frame into (pinned) host memory at out_ptr; valid after read_wait()."""
        self._check(self._lib.ddgi_read_probe_texture_async(self._ctx, which, out_ptr, nbytes))

    def read_probe_texture_rows_async(self, out_ptr: int, row0: int, row1: int, nbytes: int, which: int = 0):
        """The same for texture rows [row0, row1), ordered after this frame's exchange too."""
        self._check(self._lib.ddgi_read_probe_texture_rows_async(self._ctx, which, row0, row1, out_ptr, nbytes))

    def read_wait(self):
        self._check(self._lib.ddgi_read_wait(self._ctx))

    def write_probe_texture(self, tex: np.ndarray, which: int = 0):
        t = np.ascontiguousarray(tex, dtype=np.uint32)
        self._check(self._lib.ddgi_write_probe_texture(self._ctx, which, t.ctypes.data, t.nbytes))

    def read_frame(self, fmt: int = capi.FMT_RGBA8, out: np.ndarray | None = None):
        w, h = self.render_settings.screen_width, self.render_settings.screen_height
        if out is None:
            out = np.empty((h, w), dtype=np.uint32) if fmt == capi.FMT_RGBA8 else np.empty((h, w, 4), dtype=np.float32)
        self._check(self._lib.ddgi_read_frame(self._ctx, fmt, out.ctypes.data, out.nbytes))
        return out

    # -- checkpoint / resume ----------------------------------------------------------
    # The reference recomputes the probe texture from scratch every frame and has nothing to
    # save; with the hysteresis blend the texture carries state from frame to frame.
    def save_checkpoint(self, path: str):
        """Probe-texture dump (ddgi_save_checkpoint): magic, (W, H, time) header, albedo plane, distance plane."""
        self._check(self._lib.ddgi_save_checkpoint(self._ctx, str(path).encode(), C.c_float(self.render_settings.time)))

    def load_checkpoint(self, path: str):
        """Restores both texture planes and render_settings.time; the field must already have the dumped shape."""
        t = C.c_float()
        self._check(self._lib.ddgi_load_checkpoint(self._ctx, str(path).encode(), C.byref(t)))
        self.render_settings.time = t.value

    # -- instrumentation / multi-GPU ------------------------------------------------
    def set_debug(self, on):
        """False / True, or 2 to also record per-warp start / last-fetch / exit times (read_warp_times)."""
        self._check(self._lib.ddgi_set_debug(self._ctx, int(on)))

    def read_warp_times(self) -> np.ndarray:
        """uint64 [warps, 3]: %globaltimer ns at start, last ray taken, exit (debug level 2, variant 1)."""
        n = C.c_size_t()
        self._check(self._lib.ddgi_read_warp_times(self._ctx, None, 0, C.byref(n)))
        out = np.zeros((n.value, 3), dtype=np.uint64)
        if n.value:
            self._check(self._lib.ddgi_read_warp_times(self._ctx, out.ctypes.data, out.size, C.byref(n)))
        return out

    def set_schedule_slot(self, rays: int):
        self._check(self._lib.ddgi_set_schedule_slot(self._ctx, rays))

    def set_grid_limit(self, blocks_per_sm: int):
        self._check(self._lib.ddgi_set_grid_limit(self._ctx, blocks_per_sm))

    def set_kernel_variant(self, v: int):
        self._check(self._lib.ddgi_set_kernel_variant(self._ctx, v))
        self.kernel_variant = v

    def set_tuning(self, march_min: int):
        self._check(self._lib.ddgi_set_tuning(self._ctx, march_min))

    def set_auto_schedule(self, on: bool):
        self._check(self._lib.ddgi_set_auto_schedule(self._ctx, 1 if on else 0))

    def read_lookup_counts(self, which: int = 0) -> np.ndarray:
        if which == 0:
            out = np.empty(self.num_probe_rays, dtype=np.uint32)
        else:
            out = np.empty(self.render_settings.screen_width * self.render_settings.screen_height, dtype=np.uint32)
        self._check(self._lib.ddgi_read_lookup_counts(self._ctx, which, out.ctypes.data, out.size))
        return out

    @property
    def launch_count(self) -> int:
        return int(self._lib.ddgi_launch_count(self._ctx))

    def set_probe_rows(self, y0: int, y1: int):
        self._check(self._lib.ddgi_set_probe_rows(self._ctx, y0, y1))

    def set_probe_rows_cyclic(self, rank: int, world: int, block: int = 1):
        self._check(self._lib.ddgi_set_probe_rows_cyclic(self._ctx, rank, world, block))

    def set_probes_cyclic(self, rank: int, world: int, block: int = 1):
        self._check(self._lib.ddgi_set_probes_cyclic(self._ctx, rank, world, block))

    def probe_texture_device_ptr(self, which: int = 0):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._lib.ddgi_probe_texture_device_ptr(self._ctx, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def export_texture_handle(self) -> bytes:
        """CUDA IPC handle(s) of the texture allocation(s): 64 bytes, or 128 for a double-buffered context."""
        buf = C.create_string_buffer(128)
        n = C.c_int32()
        self._check(self._lib.ddgi_export_texture_handles(self._ctx, buf, C.byref(n)))
        return buf.raw[:64 * n.value]

    def open_peers(self, handles: list[bytes], self_index: int):
        """handles[g] = what rank g's export_texture_handle returned (every rank configured alike)."""
        blob = b"".join(handles)
        self._check(self._lib.ddgi_open_peers(self._ctx, len(handles), blob, self_index))

    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte ncclUniqueId (one rank creates it, the host passes it to the others)."""
        buf = C.create_string_buffer(128)
        rc = capi.load().ddgi_comm_unique_id(buf)
        if rc != capi.OK:
            raise DDGIError(rc, "ddgi_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        self._check(self._lib.ddgi_comm_init(self._ctx, unique_id, rank, world))

    def comm_destroy(self):
        self._check(self._lib.ddgi_comm_destroy(self._ctx))

    def exchange_allgather(self):
        """In-place NCCL exchange of the probe rows after probe_update (ddgi_exchange_allgather)."""
        self._check(self._lib.ddgi_exchange_allgather(self._ctx, self.stream))

    def set_sample_order(self, y_first: bool):
        """False (default): x jitter drawn first; True: y first, what g++ makes of rvpt.cpp:1161-1162."""
        self._check(self._lib.ddgi_set_sample_order(self._ctx, 1 if y_first else 0))

    def set_frame_band(self, rank: int, world: int):
        """This context renders band `rank` of `world` bands of 16-pixel rows; returns its pixel rows (y0, y1)."""
        self._check(self._lib.ddgi_set_frame_band(self._ctx, rank, world))
        y0, y1 = C.c_int32(), C.c_int32()
        self._check(self._lib.ddgi_frame_band_rows(self._ctx, C.byref(y0), C.byref(y1)))
        return y0.value, y1.value

    def exchange_barrier(self):
        """Device-side completion barrier of the fused exchange (after probe_update, every rank)."""
        self._check(self._lib.ddgi_exchange_barrier(self._ctx, self.stream))

    def exchange_status(self):
        """Raises if a peer missed an exchange barrier (synchronises)."""
        self._check(self._lib.ddgi_exchange_status(self._ctx))

    def close_peers(self):
        self._check(self._lib.ddgi_close_peers(self._ctx))


from .sharding import probe_row_shard  # noqa: E402  (kept importable from here)
