"""ULP error of lgar::pow_inverse_root (the shortcut for Se^(1/m) right after Se = 1 / (1+p)^m) against mpmath,
next to the full table-driven pow and glibc's pow on the same arguments."""
import math, os, subprocess, sys
import mpmath as mp
mp.mp.prec = 200
here = os.path.dirname(os.path.abspath(__file__))
exe = os.path.join(here, "pow_inverse_root_accuracy")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-w", exe + ".cpp", "-o", exe])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
out = subprocess.run([exe, str(n)], capture_output=True, text=True)
print(out.stderr.strip())
worst = {"fast": 0.0, "full": 0.0, "glibc": 0.0}
wrong = {"fast": 0, "full": 0, "glibc": 0}
worst_scaled = 0.0
cnt = 0
for line in out.stdout.splitlines():
    u, m, se, fast, full, gl = (float.fromhex(t) for t in line.split())
    inv_m = 1.0 / m
    t = mp.power(mp.mpf(se), mp.mpf(inv_m))
    ulp = mp.mpf(math.ulp(gl))
    cnt += 1
    for k, v in (("fast", fast), ("full", full), ("glibc", gl)):
        e = float(abs((mp.mpf(v) - t) / ulp))
        worst[k] = max(worst[k], e)
        wrong[k] += e > 0.5
        if k == "fast":
            worst_scaled = max(worst_scaled, (e - 0.5) * m)
print(f"n={cnt}  max error [ulp]: shortcut {worst['fast']:.4f}, full pow {worst['full']:.4f}, glibc {worst['glibc']:.4f};  "
      f"not correctly rounded: shortcut {100*wrong['fast']/cnt:.3f} %, full pow {100*wrong['full']/cnt:.3f} %, glibc {100*wrong['glibc']/cnt:.3f} %;  "
      f"max (error - 0.5) * m = {worst_scaled:.5f}")
