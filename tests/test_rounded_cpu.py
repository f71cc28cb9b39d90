"""CPU check of advance_rounded() (lgar-py_b200/csrc/lgar_rounded.cuh): the exact result of k rounded additions
x = fl(x + s) in O(number of binades), which lets the CUDA root finders jump along runs of equal steps while visiting
exactly the values the reference's one-step-at-a-time loops visit.  The header compiles as plain C++ with the same
IEEE operations as on the device; tools/rounded_check.cpp compares it with the literal chain on psi runs, Geff node
chains, binade edges, half-ulp ties, negative values and long check_column_mass runs."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_advance_rounded_equals_literal_chain(tmp_path):
    exe = str(tmp_path / "rounded_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-w",
                           os.path.join(ROOT, "tools", "rounded_check.cpp"), "-o", exe])
    out = subprocess.run([exe, "120000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout
    n_exact = int(out.stdout.split("exact comparisons")[1].split(",")[0])
    assert n_exact > 100000
