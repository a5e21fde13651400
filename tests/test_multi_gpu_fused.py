"""Two GPUs, two processes: the fused exchange (texels stored into the peers' replicas through CUDA IPC mappings, device
epoch barriers) under the access pattern that can mix frames - moving lights, a render and a read of the whole replica
every frame, one rank lagging on the host - for single- and double-buffered replicas and with two frames in flight.
Every texture and every rendered band any rank ever read must equal the single-GPU engine's for that frame (round 1's
ADVICE: `bench --verify` only compares replicas at a quiescent point).  The same loop also runs over the NCCL exchange of probe-cyclic ownership (tiles packed into one chunk per rank, ONE
ncclAllGather, unpacked: ddgi_exchange_allgather).  Skipped on a box with one GPU (gpurun --gpus 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import util

ddgi_b200 = util.ddgi_b200
CFG = util.configs.CONFIGS
HERE = os.path.dirname(os.path.abspath(__file__))
FRAMES = 8


def reference_frames():
    cfg = CFG["field_8"]
    tex, img = [], []
    with ddgi_b200.RVPT(*cfg["screen"]) as r:
        util.configs.apply(r, cfg)
        r.generate_probe_rays(reseed=True)
        for f in range(FRAMES):
            r.render_settings.time = 2.0 * (f + 1)
            r.lights = util.configs.lights_for(cfg, r.render_settings.time)
            r.update(advance_time=False)
            r.draw()
            r.sync()
            tex.append(r.read_probe_texture(0).copy())
            img.append(r.read_frame().copy())
    return tex, img


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("in_flight,double_buffer,exchange", [(1, 0, "fused"), (1, 1, "fused"), (2, 1, "fused"), (1, 0, "nccl"), (1, 1, "nccl")])
def test_fused_exchange_never_mixes_frames(tmp_path, in_flight, double_buffer, exchange):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = 2
    tex, img = reference_frames()
    assert not np.array_equal(tex[0], tex[1]), "the lights must move between frames for the test to mean anything"
    outs = [str(tmp_path / f"rank{r}.npz") for r in range(world)]
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "multi", "fused_worker.py"), str(r), str(world), str(in_flight),
                               str(double_buffer), str(tmp_path), outs[r], exchange], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(world)]
    for r, p in enumerate(procs):
        so, se = p.communicate(timeout=240)
        assert p.returncode == 0, f"rank {r}: {se[-2000:]}"
    for r in range(world):
        d = np.load(outs[r])
        b0, b1 = (int(v) for v in d["band"])
        for f in range(FRAMES):
            assert np.array_equal(d["tex"][f], tex[f]), f"rank {r}: replica read in frame {f} is not frame {f}'s texture"
        for i, f in enumerate(range(1, FRAMES, 2)):
            assert np.array_equal(d["frames"][i], img[f][b0:b1]), f"rank {r}: band rendered in frame {f}"
