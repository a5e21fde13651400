"""The C++ host mirror of the reference's class RVPT / Camera (include/rvpt_ddgi.hpp) — the
reference's host is compiled C++ (src/rvpt/rvpt.cpp, main.cpp), so the drop-in is exercised from
compiled code too: tests/host_cpp/host_main.cpp drives generate_probe_rays / initialize / update /
draw through the C-ABI exactly as src/rvpt/main.cpp:37-96 drives the reference."""
import os
import subprocess

import numpy as np
import pytest

import ddgi_b200

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "host_cpp", "host_main")


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "host_cpp")], stdout=subprocess.DEVNULL)
    return EXE


def camera_block(aspect, o, r):
    out = subprocess.check_output([build(), "camera", repr(float(aspect)), *(repr(float(v)) for v in (*o, *r))], text=True)
    return np.array([int(w, 16) for w in out.split()], dtype=np.uint32).view(np.float32)


def test_cpp_camera_block_matches_the_python_mirror():
    """Camera::get_data (camera.cpp:100-111) restated twice: glm-style fp32 in C++, fp64 rounded once
    in Python.  Identical for axis-aligned views, within an fp32 rounding otherwise."""
    a = camera_block(1.0, (0.0, 0.0, -5.0), (0.0, 0.0, 0.0))
    b = ddgi_b200.Camera(1.0, (0.0, 0.0, -5.0), (0.0, 0.0, 0.0)).get_data()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for aspect, o, r in ((1600 / 900.0, (1.5, 2.0, -2.0), (-38.0, 36.0, 0.0)), (2.0, (3.0, -1.0, 7.5), (200.0, -80.0, 15.0))):
        a = camera_block(aspect, o, r)
        b = ddgi_b200.Camera(aspect, o, r).get_data()
        assert a.shape == b.shape == (20,)
        assert np.abs(a - b).max() <= 2.5e-7 * max(1.0, float(np.abs(b).max()))
        assert a[16] == np.float32(aspect) and a[18] == 4.0 and a[19] == 0.0


def test_cpp_host_reports_a_missing_device_instead_of_falling_back():
    import torch

    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    p = subprocess.run([build(), "frame", "1", "3", "3", "3", "11", "8", "0", "0", "15", "128", "128", "0", "0", "-5", "0", "0", "0", "1",
                        "/tmp/never_written.bin"], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("in_flight", [None, 2])
def test_cpp_host_main_loop_reproduces_the_reference_fixture(tmp_path, in_flight):
    """generate_probe_rays() / initialize() / update() / draw() from C++ on Cornell 3x3x3: probe texture
    and frame must equal the reference-shader fixture (the texture does not depend on the frame
    number: static lights, frame-invariant RNG) - one frame at a time and with two frames in flight
    (ddgi_set_frames_in_flight, as the reference's MAX_FRAMES_IN_FLIGHT = 2)."""
    g = np.load(os.path.join(HERE, "golden", "cornell_3x3x3.npz"))
    w, h = (int(v) for v in g["screen"])
    out = str(tmp_path / "frame.bin")
    args = ["frame", "1", "3", "3", "3", "11", "8", "0", "0", "15", str(w), str(h), "0", "0", "-5", "0", "0", "0", "3", out]
    env = dict(os.environ)
    env.pop("DDGI_FRAMES_IN_FLIGHT", None)
    if in_flight:
        env["DDGI_FRAMES_IN_FLIGHT"] = str(in_flight)
    msg = subprocess.check_output([build(), *args], text=True, env=env)
    assert msg.startswith("ok 1728 probe rays, time 6.0")
    raw = np.fromfile(out, dtype=np.uint32)
    W, H, fw, fh = (int(v) for v in raw[:4].view(np.int32))
    assert (H, W) == g["albedo"].shape and (fw, fh) == (w, h)
    tex = raw[4:4 + W * H].reshape(H, W)
    frame = raw[4 + W * H:].reshape(h, w)
    assert np.array_equal(tex, g["albedo"])
    assert np.array_equal(frame, g["frame"])


def _run_shards(tmp_path, world):
    idfile = str(tmp_path / "nccl.id")
    outs = [str(tmp_path / f"rank{r}.bin") for r in range(world)]
    procs = [subprocess.Popen([build(), "shard", str(r), str(world), str(r), idfile, outs[r]], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(world)]
    for r, p in enumerate(procs):
        so, se = p.communicate(timeout=240)
        assert p.returncode == 0, f"rank {r}: rc {p.returncode}\n{so}\n{se}"
        assert f"ok rank {r} of {world}, 1024 probe rays" in so   # (NCCL prints its version to stdout first)
    import util
    from oracle import oracle
    cfg = dict(util.configs.CONFIGS["cornell_3x3x3"], probe_count=(2, 4, 2), side_length=7, screen=(64, 64))
    sc = util.oracle_scene(cfg)
    rays = oracle.generate_probe_rays(sc, oracle.generate_samples(8, 8, reseed=True))
    want, *_ = oracle.probe_update(sc, rays)
    for r in range(world):
        raw = np.fromfile(outs[r], dtype=np.uint32)
        W, H, rank, w = (int(v) for v in raw[:4].view(np.int32))
        assert (H, W, rank, w) == (*want.shape, r, world)
        assert np.array_equal(raw[4:].reshape(H, W), want), f"rank {r}: the exchanged texture differs from a full update"


@pytest.mark.gpu
def test_cpp_host_nccl_exchange_single_rank(tmp_path):
    """ddgi_comm_unique_id / ddgi_comm_init / ddgi_exchange_allgather from compiled C++ with one rank (the in-place
    all-gather of a 1-rank communicator), plus the checkpoint file round trip through the C entry points."""
    _run_shards(tmp_path, 1)


@pytest.mark.gpu
def test_cpp_host_nccl_exchange_two_gpus(tmp_path):
    """Two processes, two GPUs: each updates its slab of probe rows, ONE in-place ncclAllGather, every
    rank ends with the texture a full update gives."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    _run_shards(tmp_path, 2)
