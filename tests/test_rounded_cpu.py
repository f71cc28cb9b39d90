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


def test_column_mass_search_with_jumps_equals_literal_loop(tmp_path):
    """tools/colmass_jump_check.cpp: a host restatement of the jumping search of Column::check_column_mass
    (lgar_device.cuh; probes on depths from the real advance_rounded header) against the literal loop of the reference
    (Layer.py:681-701) on column-mass functions of every slope, FLAT ones included (theta of the free-drainage front
    equal to theta of the front below: the reference never leaves the loop, the library reports ITER_CAP): capped or
    not, final depth bit for bit, iteration count -- and a capped search must not walk its iterations one by one."""
    exe = str(tmp_path / "colmass_jump_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-w",
                           os.path.join(ROOT, "tools", "colmass_jump_check.cpp"), "-o", exe])
    out = subprocess.run([exe, "6000", "100000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout
    lit = int(out.stdout.split("literal ")[1].split(" ")[0])
    jmp = int(out.stdout.split("jumping ")[1].split(",")[0])
    assert int(out.stdout.split("capped ")[1].split(",")[0]) > 1000   # the cap is exercised
    assert jmp * 10 < lit, out.stdout                                  # the jumps do the work
