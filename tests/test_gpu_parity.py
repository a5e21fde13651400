"""GPU parity: the CUDA engine, called through the C-ABI, against the CPU oracle on the
same inputs.  Bit-exact is the bar: RGBA8 bytes, the fp32 value before quantisation
(compared as bit patterns) and the per-ray voxel-lookup counts.

north_star's stated tolerance is 1e-3 relative L-inf on the fp32 texture / frame; the
tests assert the stronger max |diff| == 0 and print the tolerance figure alongside.
"""
import numpy as np
import pytest

import ddgi_b200
import util
from oracle import oracle

pytestmark = pytest.mark.gpu

CFG = util.configs.CONFIGS
TOL_REL_LINF = 1e-3  # north_star tolerance; we require exact equality below


def make_engine(cfg, *, debug=True, time=0.0):
    r = ddgi_b200.RVPT(*cfg["screen"])
    r.set_debug(debug)
    util.configs.apply(r, cfg, time=time)
    r.generate_probe_rays(reseed=True)
    r.update(advance_time=False)
    return r


def oracle_rays(sc, cfg):
    rx, ry = cfg["tile"]
    return oracle.generate_probe_rays(sc, oracle.generate_samples(rx, ry, reseed=True))


def rel_linf(a, b):
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / denom


@pytest.mark.parametrize("name", ["cornell_2x2x2", "cornell_3x3x3", "cave_64", "field_8"])
def test_bake_matches_oracle(name):
    cfg = CFG[name]
    with ddgi_b200.RVPT(64, 64) as r:
        util.configs.apply(r, cfg)
        got = r.read_voxels(cfg["voxels"][1])
    want, _ = util.oracle_voxels(cfg)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", ["cornell_2x2x2", "cave_64", "field_8"])
def test_generated_rays_match_reference_generator(name):
    cfg = CFG[name]
    sc = util.oracle_scene(cfg)
    want = oracle_rays(sc, cfg)
    with make_engine(cfg) as r:
        got = r.probe_rays
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("ray_mode", ["generated", "ssbo"])
@pytest.mark.parametrize("name", ["cornell_2x2x2", "cornell_3x3x3", "cave_64", "field_8"])
def test_probe_texture_parity(name, ray_mode, variant):
    cfg = CFG[name]
    sc = util.oracle_scene(cfg)
    rays = oracle_rays(sc, cfg)
    alb, dist, f32, steps, _ = oracle.probe_update(sc, rays)
    with make_engine(cfg) as r:
        if ray_mode == "ssbo":
            r.set_probe_rays(rays)
        r.set_kernel_variant(variant)
        r.probe_update()
        r.sync()
        got = r.read_probe_texture(0)
        got_dist = r.read_probe_texture(1)
        got_f32 = r.read_probe_texture(0, ddgi_b200.capi.FMT_F32)
        got_steps = r.read_lookup_counts(0)
    print(f"{name}: rel Linf fp32 = {rel_linf(got_f32[..., :3], f32[..., :3]):.3e} (tolerance {TOL_REL_LINF})")
    util.assert_lookups(got_steps, steps, variant, name)
    assert np.array_equal(got_f32.view(np.uint32), f32.view(np.uint32))
    assert np.array_equal(got, alb)
    assert np.array_equal(got_dist, dist)
    assert (got_dist == 0).all()


@pytest.mark.parametrize("name", ["cornell_2x2x2", "cornell_3x3x3", "field_8"])
def test_frame_parity(name):
    cfg = CFG[name]
    sc = util.oracle_scene(cfg)
    rays = oracle_rays(sc, cfg)
    alb, *_ = oracle.probe_update(sc, rays)
    cam = util.camera_block(cfg)
    frame, f32, steps = oracle.render_frame(sc, cam, alb)
    with make_engine(cfg) as r:
        r.draw()
        r.sync()
        got = r.read_frame()
        got_f32 = r.read_frame(ddgi_b200.capi.FMT_F32)
        got_steps = r.read_lookup_counts(1).reshape(got.shape)
    w, h = cfg["screen"]
    wx, hy = (w // 16) * 16, (h // 16) * 16
    print(f"{name}: frame rel Linf fp32 = {rel_linf(got_f32[:hy, :wx, :3], f32[:hy, :wx, :3]):.3e}")
    assert np.array_equal(got_steps[:hy, :wx], steps[:hy, :wx])
    assert np.array_equal(got_f32[:hy, :wx].view(np.uint32), f32[:hy, :wx].view(np.uint32))
    assert np.array_equal(got[:hy, :wx], frame[:hy, :wx])
    # pixels outside the reference's truncated dispatch are never written
    assert (got[hy:, :] == 0).all() and (got[:, wx:] == 0).all()


def test_cave_1080p_frame_rows_sample():
    """cfg 2 at its full 1920x1080: oracle on a 64-row band (the full frame takes too long on CPU)."""
    cfg = CFG["cave_64"]
    sc = util.oracle_scene(cfg)
    rays = oracle_rays(sc, cfg)
    alb, *_ = oracle.probe_update(sc, rays)
    with make_engine(cfg, debug=False) as r:
        r.draw()
        r.sync()
        got = r.read_frame()
    # oracle renders a 1920 x 64 window whose rows are rows [512, 576) of the full frame:
    # the pixel's y coordinate only enters through gy / h, so render full-height rows on
    # a band by asking for the full frame lazily is not possible; instead compare a
    # reduced-height render of identical per-pixel maths: same w, h but only check rows
    frame, _, _ = oracle.render_frame(sc, util.camera_block(cfg), alb)
    hy = (1080 // 16) * 16
    assert np.array_equal(got[:hy], frame[:hy])
    assert (got[hy:] == 0).all()


def test_dynamic_lights_change_the_texture_and_stay_in_parity():
    cfg = CFG["field_8"]
    tex = []
    for time in (0.0, 40.0):
        sc = util.oracle_scene(cfg, time=time)
        rays = oracle_rays(sc, cfg)
        alb, *_ = oracle.probe_update(sc, rays)
        with make_engine(cfg, time=time) as r:
            r.probe_update()
            r.sync()
            got = r.read_probe_texture(0)
        assert np.array_equal(got, alb)
        tex.append(got)
    assert not np.array_equal(tex[0], tex[1])


def test_probe_row_shards_compose_to_the_full_texture():
    """Two contexts each update half the probe rows; the halves tile the full result."""
    cfg = CFG["field_8"]
    with make_engine(cfg, debug=False) as r:
        r.probe_update()
        r.sync()
        full = r.read_probe_texture(0)
    rows = cfg["probe_count"][1]
    ry = cfg["tile"][1]
    parts = []
    for rank in range(2):
        y0, y1 = ddgi_b200.probe_row_shard(rows, rank, 2)
        with make_engine(cfg, debug=False) as r:
            r.set_probe_rows(y0, y1)
            r.probe_update()
            r.sync()
            t = r.read_probe_texture(0)
            assert (t[: y0 * ry] == 0).all() and (t[y1 * ry :] == 0).all()
            parts.append(t[y0 * ry : y1 * ry])
    assert np.array_equal(np.concatenate(parts, axis=0), full)


def test_block_cyclic_shards_compose_to_the_full_texture():
    """ddgi_set_probe_rows_cyclic: three ranks, block 1 and 2, ragged (8 rows / 3)."""
    cfg = CFG["field_8"]
    with make_engine(cfg, debug=False) as r:
        r.probe_update()
        r.sync()
        full = r.read_probe_texture(0)
    rows = cfg["probe_count"][1]
    ry = cfg["tile"][1]
    for block in (1, 2):
        acc = np.zeros_like(full)
        for rank in range(3):
            owned = ddgi_b200.sharding.probe_row_blocks(rows, rank, 3, block)
            for variant in (0, 1, 2):
                with make_engine(cfg, debug=False) as r:
                    r.set_kernel_variant(variant)
                    r.set_probe_rows_cyclic(rank, 3, block)
                    r.probe_update()
                    r.sync()
                    t = r.read_probe_texture(0)
                mask = np.zeros(t.shape[0], dtype=bool)
                for a, b in owned:
                    mask[a * ry:b * ry] = True
                assert (t[~mask] == 0).all()
                assert np.array_equal(t[mask], full[mask])
            acc[mask] = t[mask]
        assert np.array_equal(acc, full)


def test_probe_cyclic_shards_compose_to_the_full_texture():
    """ddgi_set_probes_cyclic: 3 ranks over 512 probes (ragged), blocks of 1 and 5 probes."""
    cfg = CFG["field_8"]
    with make_engine(cfg, debug=False) as r:
        r.probe_update()
        r.sync()
        full = r.read_probe_texture(0)
    X, Y, Z = cfg["probe_count"]
    rx, ry = cfg["tile"]
    tiles = full.reshape(Y, ry, X * Z, rx).transpose(0, 2, 1, 3).reshape(X * Y * Z, ry, rx)  # [probe, ty, tx]
    for block in (1, 5):
        owner = (np.arange(X * Y * Z) // block) % 3
        for rank in range(3):
            with make_engine(cfg, debug=False) as r:
                r.set_probes_cyclic(rank, 3, block)
                r.probe_update()
                r.sync()
                t = r.read_probe_texture(0)
            tt = t.reshape(Y, ry, X * Z, rx).transpose(0, 2, 1, 3).reshape(X * Y * Z, ry, rx)
            assert np.array_equal(tt[owner == rank], tiles[owner == rank])
            assert (tt[owner != rank] == 0).all()


def test_cost_ordered_schedule_does_not_change_results():
    """The first update measures per-probe costs, later ones trace expensive probes first:
    same bytes, same per-ray lookup counts, with the schedule on or off, on every variant."""
    cfg = CFG["field_8"]
    sc = util.oracle_scene(cfg)
    alb, _, _, steps, _ = oracle.probe_update(sc, oracle_rays(sc, cfg))
    for variant in (0, 1, 2):
        for on in (True, False):
            with make_engine(cfg) as r:
                r.set_kernel_variant(variant)
                r.set_auto_schedule(on)
                for _ in range(3):  # calibrating update, then two scheduled ones
                    r.write_probe_texture(np.zeros_like(alb))
                    r.probe_update()
                    r.sync()
                    assert np.array_equal(r.read_probe_texture(0), alb)
                    util.assert_lookups(r.read_lookup_counts(0), steps, variant)


def test_idempotent_and_tuning_independent():
    cfg = CFG["field_8"]
    with make_engine(cfg, debug=False) as r:
        r.probe_update()
        r.sync()
        a = r.read_probe_texture(0)
        for m in (1, 8, 24, 32):
            r.set_tuning(m)
            r.probe_update()
            r.sync()
            assert np.array_equal(r.read_probe_texture(0), a)
        n0 = r.launch_count
        r.probe_update()
        assert r.launch_count == n0 + 1


def test_errors_are_reported_not_thrown():
    with ddgi_b200.RVPT(64, 64) as r:
        with pytest.raises(ddgi_b200.DDGIError) as e:
            r.probe_update()
        assert e.value.code == ddgi_b200.capi.E_STATE
        r.render_settings.camera_mode = 1  # ortho / spherical cameras are out of scope
        with pytest.raises(ddgi_b200.DDGIError) as e:
            r.update()
        assert e.value.code == ddgi_b200.capi.E_INVALID


def test_empty_screen_and_zero_bounces():
    cfg = util.small(CFG["cornell_2x2x2"], screen=(8, 8), max_bounces=0)
    sc = util.oracle_scene(cfg)
    rays = oracle_rays(sc, cfg)
    alb, *_ = oracle.probe_update(sc, rays)
    with make_engine(cfg) as r:
        r.draw()  # 8x8 screen: floor(8/16) = 0 workgroups -> nothing is written
        r.sync()
        assert (r.read_frame() == 0).all()
        assert np.array_equal(r.read_probe_texture(0), alb)


def test_frame_bands_compose_to_the_whole_frame():
    """ddgi_set_frame_band: 3 bands of 16-pixel workgroup rows rendered one after another into
    the same frame buffer equal the frame rendered at once (the multi-GPU pixel-pass split)."""
    cfg = util.small(CFG["cornell_3x3x3"], screen=(96, 112))  # 7 workgroup rows: uneven bands
    with make_engine(cfg) as r:
        r.draw()
        r.sync()
        whole = r.read_frame().copy()
        assert (whole != 0).any()
        rows = []
        for band in range(3):
            y0, y1 = r.set_frame_band(band, 3)
            assert (y0, y1) == ddgi_b200.sharding.frame_band_rows(112, band, 3)   # the host-side mirror agrees
            rows.append((y0, y1))
            r.render_frame()
            r.sync()
            got = r.read_frame()
            assert np.array_equal(got[y0:y1], whole[y0:y1])
        assert rows[0][0] == 0 and rows[-1][1] == 112 and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        # a band writes nothing outside its rows
        r.set_frame_band(1, 3)
        # (no frame upload entry point: re-create the frame buffer by resizing the screen back and forth)
        r.render_settings.screen_width, r.render_settings.screen_height = 64, 64
        r.update(advance_time=False)
        r.render_settings.screen_width, r.render_settings.screen_height = 96, 112
        r.update(advance_time=False)
        r.render_frame()
        r.sync()
        got = r.read_frame()
        y0, y1 = rows[1]
        assert np.array_equal(got[y0:y1], whole[y0:y1]) and (got[:y0] == 0).all() and (got[y1:] == 0).all()


def test_double_buffered_async_reads_match_single_buffered_frames():
    """ddgi_set_double_buffer + ddgi_read_probe_texture_async: four frames with moving lights and
    the hysteresis blend (which must read the PREVIOUS frame's buffer) give the same textures as
    the single-buffered engine, each read back while the next frame is being traced."""
    cfg = CFG["field_8"]
    frames = 4

    def run(double):
        out = []
        with make_engine(cfg, debug=False) as r:
            r.ir.hysteresis = 0.6
            r.set_blend_mode(ddgi_b200.capi.BLEND_HYSTERESIS)
            r.set_double_buffer(double)
            W, H = r.probe_texture_size
            bufs = [np.zeros((H, W), dtype=np.uint32) for _ in range(frames)]
            for f in range(frames):
                r.render_settings.time = 2.0 * (f + 1)
                r.lights = util.configs.lights_for(cfg, r.render_settings.time)
                r.update(advance_time=False)
                r.probe_update()
                if double:
                    r.read_probe_texture_async(bufs[f].ctypes.data, bufs[f].nbytes, 0)
                else:
                    r.sync()
                    bufs[f][:] = r.read_probe_texture(0)
            r.read_wait()
            r.sync()
            out = [b.copy() for b in bufs]
            last = r.read_probe_texture(0)
            assert np.array_equal(last, out[-1])   # the synchronous read sees the latest buffer
            r.render_frame()
            r.sync()
            return out, r.read_frame().copy()

    single, frame_s = run(False)
    double, frame_d = run(True)
    for f in range(frames):
        assert np.array_equal(single[f], double[f]), f"frame {f}"
    assert not np.array_equal(single[0], single[-1])
    assert np.array_equal(frame_s, frame_d)


@pytest.mark.parametrize("blend", [False, True])
def test_two_frames_in_flight_give_the_same_frames(blend):
    """ddgi_set_frames_in_flight(2): updates run on the engine's own two streams and overlap; six frames with moving
    lights (and, with the hysteresis blend, a dependency between consecutive frames), each rendered and read back
    asynchronously, and a voxel edit in the middle, equal the frames of the plain single-stream engine."""
    cfg = CFG["field_8"]
    frames = 6
    edit_at = 3
    box = np.full((6, 6, 6), 4, dtype=np.uint8)

    def run(in_flight):
        with make_engine(cfg, debug=False) as r:
            if blend:
                r.ir.hysteresis = 0.6
                r.set_blend_mode(ddgi_b200.capi.BLEND_HYSTERESIS)
            r.set_double_buffer(True)
            r.set_frames_in_flight(in_flight)
            W, H = r.probe_texture_size
            w, h = cfg["screen"]
            tex = [np.zeros((H, W), dtype=np.uint32) for _ in range(frames)]
            img = []
            for f in range(frames):
                r.render_settings.time = 2.0 * (f + 1)
                r.lights = util.configs.lights_for(cfg, r.render_settings.time)
                r.update(advance_time=False)
                if f == edit_at:
                    r.edit_voxels(box, (2, 3, -4))   # must follow the updates in flight, precede this one
                r.probe_update()
                r.read_probe_texture_async(tex[f].ctypes.data, tex[f].nbytes, 0)
                r.render_frame()
                img.append(r.read_frame().copy())    # (synchronous: waits for the frame's render)
            r.read_wait()
            r.sync()
            assert np.array_equal(r.read_probe_texture(0), tex[-1])
            return [t.copy() for t in tex], img

    tex1, img1 = run(1)
    tex2, img2 = run(2)
    for f in range(frames):
        assert np.array_equal(tex1[f], tex2[f]), f"texture of frame {f}"
        assert np.array_equal(img1[f], img2[f]), f"image of frame {f}"
    assert not np.array_equal(tex1[edit_at - 1], tex1[edit_at])


def test_field_32_full_size_sampled_parity_and_properties():
    """BASELINE configs[3] at its FULL size (32^3 probes x 256 rays = 8 388 608 probe rays, 512^3
    voxels, 4 moving lights).  The oracle cannot trace the whole field in seconds, so:
      * 48 probes drawn at random (plus the field's corners and centre) are traced by the oracle
        on the engine's own baked voxels and compared texel for texel and lookup for lookup;
      * idempotence: a second update gives the same bytes; schedule / slot independence;
      * composition: 4 round-robin probe shards tile the full texture (what 4 GPUs compute);
      * a checksum of per-row checksums over all 16 384 x 512 texels pins the frame for the record.
    The baker itself is checked against its numpy restatement on field_8 (test_bake_matches_oracle)."""
    cfg = CFG["field_32"]
    X, Y, Z = cfg["probe_count"]
    rx, ry = cfg["tile"]
    n = rx * ry
    with make_engine(cfg, debug=True, time=6.0) as r:
        r.set_kernel_variant(1)   # performs exactly the reference algorithm's voxel lookups
        r.probe_update()
        r.sync()
        full = r.read_probe_texture(0).copy()
        lk = r.read_lookup_counts(0).reshape(X * Y * Z, n)
        r.set_kernel_variant(2)   # the default: result-preserving early-outs, never more lookups
        r.write_probe_texture(np.zeros_like(full))
        r.probe_update()
        r.sync()
        assert np.array_equal(r.read_probe_texture(0), full)
        assert (r.read_lookup_counts(0).reshape(X * Y * Z, n) <= lk).all()
        vox = r.read_voxels(cfg["voxels"][1])
        # --- sampled oracle parity on the engine's voxels
        sc = util.oracle_scene(cfg, time=6.0, voxels=vox)
        rays = oracle_rays(sc, cfg)
        rng = np.random.default_rng(32)
        probes = sorted(set(rng.integers(0, X * Y * Z, size=48).tolist()) | {0, X * Y * Z - 1, (Y // 2 * Z + Z // 2) * X + X // 2})
        tiles = full.reshape(Y, ry, X * Z, rx).transpose(0, 2, 1, 3).reshape(X * Y * Z, ry, rx)
        W, H = sc.tex_size
        want_tex = np.zeros((H, W), dtype=np.uint32)
        open_probes = 0
        for p in probes:
            _, _, _, steps, _ = oracle.probe_update(sc, rays, p * n, (p + 1) * n, tex=want_tex)
            assert np.array_equal(lk[p], steps[p * n:(p + 1) * n]), f"probe {p}: lookup counts differ"
            open_probes += int(steps[p * n:(p + 1) * n].max() > 16)
        want_tiles = want_tex.reshape(Y, ry, X * Z, rx).transpose(0, 2, 1, 3).reshape(X * Y * Z, ry, rx)
        assert np.array_equal(tiles[probes], want_tiles[probes])
        assert open_probes >= 5, "the sample must include probes in the open cavity, not only rock"
        mean_lookups = lk.mean()
        assert 100.0 < mean_lookups < 115.0   # the figure bench.py's algorithmic bytes are built on (~106.4)
        # --- idempotence, schedule independence
        r.set_debug(False)
        for slot, sched in ((32, True), (0, True), (32, False)):
            r.set_schedule_slot(slot)
            r.set_auto_schedule(sched)
            for _ in range(2):
                r.probe_update()
            r.sync()
            assert np.array_equal(r.read_probe_texture(0), full)
        # --- 4 round-robin shards compose to the full texture
        acc = np.zeros_like(tiles)
        owner = np.arange(X * Y * Z) % 4
        for rank in range(4):
            r.write_probe_texture(np.zeros_like(full))
            r.set_probes_cyclic(rank, 4, 1)
            r.probe_update()
            r.sync()
            t = r.read_probe_texture(0).reshape(Y, ry, X * Z, rx).transpose(0, 2, 1, 3).reshape(X * Y * Z, ry, rx)
            assert (t[owner != rank] == 0).all()
            acc[owner == rank] = t[owner == rank]
        assert np.array_equal(acc, tiles)
    # --- checksum of checksums (size-independent record of the frame)
    row_sums = full.astype(np.uint64).sum(axis=1)
    total = int((row_sums * (np.arange(H, dtype=np.uint64) + 1)).sum() % (1 << 61))
    import json, os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "field_32.json")) as f:
        recorded = json.load(f)
    assert total == recorded["albedo_row_checksum"], "checksum of row checksums differs from the full-size oracle run"
    print(f"field_32 t=6: mean lookups/ray {mean_lookups:.3f}, checksum of row checksums {total}")
    assert (full >> 24 == 255).all()   # every texel was written (alpha = 1)


@pytest.mark.parametrize("name", ["cornell_3x3x3", "cave_64"])
def test_early_out_variant_saves_lookups_and_changes_nothing(name):
    """Kernel variant 2 (the default) ends a shadow feeler's march once it has left its light behind
    (csrc/ddgi_wavefront.cuh: wf_resolve_hit): same bytes and fp32 values as the reference algorithm,
    strictly fewer voxel lookups wherever a light is visible."""
    cfg = CFG[name]
    sc = util.oracle_scene(cfg)
    alb, _, f32, steps, _ = oracle.probe_update(sc, oracle_rays(sc, cfg))
    with make_engine(cfg) as r:
        assert r.kernel_variant == 2
        r.probe_update()
        r.sync()
        lk = r.read_lookup_counts(0)
        assert np.array_equal(r.read_probe_texture(0, ddgi_b200.capi.FMT_F32).view(np.uint32), f32.view(np.uint32))
        assert np.array_equal(r.read_probe_texture(0), alb)
    assert (lk <= steps).all() and int(lk.sum()) < int(steps.sum())
    print(f"{name}: {steps.mean():.1f} voxel lookups per ray in the reference algorithm, {lk.mean():.1f} performed")
