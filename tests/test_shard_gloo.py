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
