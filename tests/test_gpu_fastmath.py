"""GPU self-test of the exact-division shortcuts of csrc/ddgi_fastmath.cuh
(tests/selftest_div.cu, built by tests/Makefile / __graft_entry__.build()):
div_tenth(x) == x / 0.1f for ALL 2^32 bit patterns, and div_markstein(a, d, 1/d) == a / d on
2^32 random operand pairs of the DDA step's ranges (adversarial significands included), and
rcp_regular(x) == 1 / x for every float with |x| in [2^-60, 2]."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_exact_division_shortcuts_on_the_device():
    exe = os.path.join(HERE, "selftest_div")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", HERE, "selftest_div"])
    out = subprocess.run([exe, "32"], capture_output=True, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "tenth_mismatch=0 markstein_mismatch=0 rcp_mismatch=0" in out.stdout
