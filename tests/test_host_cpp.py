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
def test_cpp_host_main_loop_reproduces_the_reference_fixture(tmp_path):
    """generate_probe_rays() / initialize() / update() / draw() from C++ on Cornell 3x3x3: probe texture
    and frame must equal the reference-shader fixture (the texture does not depend on the frame
    number: static lights, frame-invariant RNG)."""
    g = np.load(os.path.join(HERE, "golden", "cornell_3x3x3.npz"))
    w, h = (int(v) for v in g["screen"])
    out = str(tmp_path / "frame.bin")
    args = ["frame", "1", "3", "3", "3", "11", "8", "0", "0", "15", str(w), str(h), "0", "0", "-5", "0", "0", "0", "3", out]
    msg = subprocess.check_output([build(), *args], text=True)
    assert msg.startswith("ok 1728 probe rays, time 6.0")
    raw = np.fromfile(out, dtype=np.uint32)
    W, H, fw, fh = (int(v) for v in raw[:4].view(np.int32))
    assert (H, W) == g["albedo"].shape and (fw, fh) == (w, h)
    tex = raw[4:4 + W * H].reshape(H, W)
    frame = raw[4 + W * H:].reshape(h, w)
    assert np.array_equal(tex, g["albedo"])
    assert np.array_equal(frame, g["frame"])
