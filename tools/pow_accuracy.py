"""ULP error of lgar::pow_fast (and of glibc pow) against mpmath, on the harness output."""
import subprocess, sys, os
import mpmath as mp
mp.mp.prec = 160
here = os.path.dirname(os.path.abspath(__file__))
exe = os.path.join(here, "pow_accuracy")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", os.path.join(here, "pow_accuracy.cpp"), "-o", exe])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
out = subprocess.run([exe, str(n)], capture_output=True, text=True)
print(out.stderr.strip())
import math
worst_f = worst_g = 0.0; bad_f = bad_g = differ = 0
for line in out.stdout.splitlines():
    x, y, rf, rg = (float.fromhex(t) for t in line.split())
    t = mp.power(mp.mpf(x), mp.mpf(y))
    ulp = mp.mpf(math.ulp(rg))
    ef = abs((mp.mpf(rf) - t) / ulp); eg = abs((mp.mpf(rg) - t) / ulp)
    worst_f = max(worst_f, float(ef)); worst_g = max(worst_g, float(eg))
    bad_f += ef > 0.5; bad_g += eg > 0.5; differ += rf != rg
    if ef > 0.6: print("BAD", x.hex(), y.hex(), float(ef))
print(f"n={n}  fast: max {worst_f:.4f} ulp, not-correctly-rounded {bad_f} ({100*bad_f/n:.3f}%) | glibc: max {worst_g:.4f} ulp, ncr {bad_g} ({100*bad_g/n:.3f}%) | fast != glibc: {differ} ({100*differ/n:.3f}%)")
