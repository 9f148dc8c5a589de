"""csrc/nrldpc_bp_math.cuh on the host: the sum-product kernel's own tanh(q/2) and 2 atanh(x) (py5gphy/ldpc/
nr_ldpc_decode.py:158-163 use np.tanh / np.arctanh) stay within 3 ulp of the correctly rounded value.  The header is
host-compilable on purpose, so this needs g++ only (the device build differs in the division sequence, which the GPU
parity tests of algo='BP' cover)."""
import os
import re
import subprocess

from .conftest import ROOT


def test_bp_math_within_3_ulp(tmp_path):
    exe = str(tmp_path / "bp_math_check")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "python_5gtoolbox_b200", "csrc"),
                    "-o", exe, os.path.join(ROOT, "tests", "helpers", "bp_math_check.cpp")], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    t = float(re.search(r"tanh_half max ulp ([0-9.]+)", out).group(1))
    a = float(re.search(r"atanh_twice max ulp ([0-9.]+)", out).group(1))
    assert t <= 3.0 and a <= 3.0, out
    # tanh_half(0) = 0, tanh_half(huge) = 1, atanh_twice(0) = 0
    assert out.strip().splitlines()[-1].split() == ["0", "1", "0"], out
