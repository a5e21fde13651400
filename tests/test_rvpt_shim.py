"""SURVEY.md 8f-3, the RVPT shim, compiled: oracle/_ref/rvpt_shim_main is the reference's OWN text of RVPT::update()
(src/rvpt/rvpt.cpp:265-290), RVPT::record_compute_command_buffer() (:1096-1143) and generate_probe_rays (:1145-1224),
with struct RenderSettings / IrradianceField / ProbeRay as the reference declares them, built unmodified against a
stub of the Vulkan layer underneath (oracle/ref_glsl/vk_shim/vk_stub.h) that forwards the four copy_to() uploads and
the two vkCmdDispatch calls to include/ddgi.h (oracle/ref_glsl/build_shim.py; needs /root/reference, so the binary is
built in the build container and travels to the GPU box)."""
import os
import subprocess

import numpy as np
import pytest

import util
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "rvpt_shim_main")
needs_shim = pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/rvpt_shim_main not built (oracle/ref_glsl/build_shim.py needs /root/reference)")

ARGS = ["1", "3", "3", "3", "11", "8", "0", "0", "15", "128", "128", "0", "0", "-5", "0", "0", "0"]   # Cornell 3x3x3, 128x128


@needs_shim
def test_shim_fails_like_a_failed_device_creation_without_a_gpu(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    p = subprocess.run([EXE, *ARGS, "1", str(tmp_path / "never.bin")], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU fallback" in p.stderr


@needs_shim
@pytest.mark.gpu
def test_reference_frame_loop_drives_the_engine_through_the_stubbed_vulkan_layer(tmp_path):
    """Three frames of the reference's update() / record_compute_command_buffer() on Cornell 3x3x3: every upload and
    both dispatches land on the C-ABI (0 stub failures), the ray list is the one the reference's own generator makes
    under g++ (y jitter drawn first), and probe texture + frame equal the oracle's for exactly those rays."""
    out = str(tmp_path / "shim.bin")
    msg = subprocess.check_output([EXE, *ARGS, "3", out], text=True)
    assert msg.startswith("ok 1728 probe rays, time 6.0, 0 stub failures"), msg
    raw = np.fromfile(out, dtype=np.uint32)
    W, H, w, h = (int(v) for v in raw[:4].view(np.int32))
    tex = raw[4:4 + W * H].reshape(H, W)
    frame = raw[4 + W * H:].reshape(h, w)
    cfg = util.small(util.configs.CONFIGS["cornell_3x3x3"], screen=(128, 128))
    sc = util.oracle_scene(cfg)
    rays = oracle.generate_probe_rays(sc, oracle.generate_samples(8, 8, reseed=True, y_first=True))
    want_tex, *_ = oracle.probe_update(sc, rays)
    want_frame, *_ = oracle.render_frame(sc, util.camera_block(cfg), want_tex)
    assert (H, W) == want_tex.shape and (h, w) == want_frame.shape
    assert np.array_equal(tex, want_tex)
    assert np.array_equal(frame, want_frame)
    # and the pinned order gives a different ray set: the comparison above is not vacuous
    other = oracle.generate_probe_rays(sc, oracle.generate_samples(8, 8, reseed=True, y_first=False))
    assert not np.array_equal(rays, other)
