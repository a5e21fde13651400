"""N > 1 host logic on CPU: two gloo ranks each update their slab of probe rows (the
engine's per-ray headers on the host, tests/hostsim) and exchange the texture with the same
in-place all-gather bench.py uses on NCCL; the result must equal the single-rank texture.
Even (2 rows / 2) and ragged (3 rows / 2) splits, contiguous slabs and block-cyclic ownership."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ddgi_b200
import util
from oracle import oracle

CFG = util.configs.CONFIGS


def test_probe_row_shard_partitions_exactly():
    for rows in (1, 2, 3, 7, 8, 32):
        for world in (1, 2, 3, 4, 8):
            spans = [ddgi_b200.probe_row_shard(rows, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    rng = ddgi_b200.sharding.shard_byte_ranges(32, 8, 16384 * 4 * 16)
    assert rng[3] == (3 * 4 * 16384 * 64, 4 * 4 * 16384 * 64)
    with pytest.raises(ValueError):
        ddgi_b200.probe_row_shard(8, 2, 2)


def test_block_cyclic_ownership_partitions_exactly():
    sh = ddgi_b200.sharding
    for rows in (1, 3, 8, 9, 32):
        for world in (1, 2, 3, 8):
            for block in (1, 2, 4):
                seen = []
                for r in range(world):
                    for a, b in sh.probe_row_blocks(rows, r, world, block):
                        assert 0 <= a < b <= rows and b - a <= block
                        assert all((y // block) % world == r for y in range(a, b))
                        seen += list(range(a, b))
                assert sorted(seen) == list(range(rows))
    assert sh.cyclic_block(32, 8) == 1 and sh.cyclic_block(32, 2) == 2 and sh.cyclic_block(8, 2) == 1


def test_probe_cyclic_ownership_and_frame_bands_partition_exactly():
    sh = ddgi_b200.sharding
    for world in (1, 2, 3, 8):
        for block in (1, 5):
            owners = [sh.probe_owner(p, world, block) for p in range(100)]
            assert set(owners) == set(range(min(world, -(-100 // block))))
            assert all(owners[p] == owners[(p // block) * block] for p in range(100))
            counts = np.bincount(owners, minlength=world)
            assert counts.max() - counts.min() <= block
    for h in (8, 16, 112, 900, 1080):
        for world in (1, 2, 3, 8):
            bands = [sh.frame_band_rows(h, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == (h // 16) * 16
            assert all(a[1] == b[0] and a[0] % 16 == 0 for a, b in zip(bands, bands[1:]))
            sizes = [(b - a) // 16 for a, b in bands]
            assert max(sizes) - min(sizes) <= 1


def _fused_worker(rank, world, port, name, out_path):
    """The fused exchange's host logic on CPU: probes dealt round-robin, every rank's texels "stored
    into every replica" (here: one sum all-reduce of the disjoint planes), then each rank renders
    its frame band from the complete replica."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg, sc, rays = _scene_and_rays(name)
        X, Y, Z = cfg["probe_count"]
        rx, ry = cfg["tile"]
        n = rx * ry
        W, H = sc.tex_size
        sh = ddgi_b200.sharding
        alb = np.zeros((H, W), dtype=np.uint32)
        hs = util.hostsim()
        for p in range(X * Y * Z):
            if sh.probe_owner(p, world) == rank:
                hs.sim_probe_update(C.byref(sc.p), rays.ctypes.data, p * n, (p + 1) * n, 1, alb.ctypes.data, None, None, None)
        plane = torch.from_numpy(alb.view(np.int32).reshape(-1))
        dist.all_reduce(plane)  # disjoint non-zero texels: the sum is the union
        screen = (64, 112)      # 7 workgroup rows over 2 ranks: uneven bands
        sc.p.screen_width, sc.p.screen_height = screen
        cam = util.camera_block(util.small(cfg, screen=screen))
        frame = np.zeros((screen[1], screen[0]), dtype=np.uint32)
        hs.sim_render_frame(C.byref(sc.p), cam.ctypes.data, alb.ctypes.data, None, frame.ctypes.data, None, None)
        y0, y1 = sh.frame_band_rows(screen[1], rank, world)
        band = torch.from_numpy(np.ascontiguousarray(frame[y0:y1]).view(np.int32))
        np.save(f"{out_path}.tex.{rank}.npy", alb)
        np.save(f"{out_path}.band.{rank}.npy", band.numpy().view(np.uint32))
    finally:
        dist.destroy_process_group()


def test_two_rank_fused_exchange_and_frame_bands(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "fused")
    name = "cornell_3x3x3"
    mp.spawn(_fused_worker, args=(2, port, name, out), nprocs=2, join=True)
    cfg, sc, rays = _scene_and_rays(name)
    want = oracle.probe_update(sc, rays)[0]
    screen = (64, 112)
    sc.p.screen_width, sc.p.screen_height = screen
    frame = oracle.render_frame(sc, util.camera_block(util.small(cfg, screen=screen)), want)[0]
    bands = []
    for rank in range(2):
        assert np.array_equal(np.load(f"{out}.tex.{rank}.npy"), want), f"rank {rank}: replica differs"
        bands.append(np.load(f"{out}.band.{rank}.npy"))
    assert np.array_equal(np.concatenate(bands, axis=0), frame[:112])


def _scene_and_rays(name):
    cfg = CFG[name]
    sc = util.oracle_scene(cfg)
    rx, ry = cfg["tile"]
    return cfg, sc, oracle.generate_probe_rays(sc, oracle.generate_samples(rx, ry, reseed=True))


def _worker(rank, world, port, name, out_path, block=0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg, sc, rays = _scene_and_rays(name)
        X, Y, Z = cfg["probe_count"]
        rx, ry = cfg["tile"]
        W, H = sc.tex_size
        sh = ddgi_b200.sharding
        owned = sh.probe_row_blocks(Y, rank, world, block) if block else [sh.probe_row_shard(Y, rank, world)]
        per_row = X * Z * rx * ry
        alb = np.zeros((H, W), dtype=np.uint32)
        hs = util.hostsim()
        for y0, y1 in owned:
            hs.sim_probe_update(C.byref(sc.p), rays.ctypes.data, y0 * per_row, y1 * per_row, 1, alb.ctypes.data, None, None, None)
        mask = np.zeros(H, dtype=bool)
        for y0, y1 in owned:
            mask[y0 * ry:y1 * ry] = True
        assert (alb[~mask] == 0).all()
        plane = torch.from_numpy(alb.view(np.uint8).reshape(-1))
        if block:
            sh.allgather_probe_rows_cyclic(plane, Y, W * 4 * ry, rank, world, block)
        else:
            sh.allgather_probe_rows(plane, Y, W * 4 * ry, rank, world)
        np.save(f"{out_path}.{rank}.npy", alb)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,block", [("cornell_2x2x2", 0), ("cornell_3x3x3", 0), ("cornell_2x2x2", 1), ("cornell_3x3x3", 1)])
def test_two_rank_exchange_equals_single_rank(name, block, tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "tex")
    mp.spawn(_worker, args=(2, port, name, out, block), nprocs=2, join=True)
    cfg, sc, rays = _scene_and_rays(name)
    want = oracle.probe_update(sc, rays)[0]
    for rank in range(2):
        got = np.load(f"{out}.{rank}.npy")
        assert np.array_equal(got, want), f"rank {rank} texture differs after the exchange"
