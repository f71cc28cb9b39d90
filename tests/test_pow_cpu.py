"""CPU check of the table-driven pow (lgar-py_b200/csrc/lgar_pow.cuh): the same source compiles as plain C++ and every
operation is an IEEE fp64 add / mul / fma, so the host results are the device results.  Against mpmath (160 bit) on the
argument ranges of the van Genuchten closures: at most 0.51 ulp, and no further from the correctly rounded result
than glibc's pow, which the reference runs on ("Why a custom pow")."""
import math
import os
import subprocess

import pytest

mp = pytest.importorskip("mpmath")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pow_core_accuracy(tmp_path):
    exe = str(tmp_path / "pow_accuracy")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-w",
                           os.path.join(ROOT, "tools", "pow_accuracy.cpp"), "-o", exe])
    n = 12000
    out = subprocess.run([exe, str(n)], capture_output=True, text=True, check=True)
    assert out.stderr.strip().startswith("fallbacks: 0 of")  # every closure-range argument takes the fast path
    mp.mp.prec = 160
    worst_f = worst_g = 0.0
    wrong_f = wrong_g = differ = 0
    for line in out.stdout.splitlines():
        x, y, rf, rg = (float.fromhex(t) for t in line.split())
        t = mp.power(mp.mpf(x), mp.mpf(y))
        ulp = mp.mpf(math.ulp(rg))
        ef, eg = float(abs((mp.mpf(rf) - t) / ulp)), float(abs((mp.mpf(rg) - t) / ulp))
        worst_f, worst_g = max(worst_f, ef), max(worst_g, eg)
        wrong_f += ef > 0.5
        wrong_g += eg > 0.5
        differ += rf != rg
    assert worst_f <= 0.51, f"max error {worst_f:.4f} ulp"
    assert wrong_f <= max(wrong_g, 1) * 2 + 3          # not correctly rounded about as rarely as glibc (~0.05 %)
    assert differ <= 0.003 * n                         # differs from glibc in well under 0.3 % of the calls
