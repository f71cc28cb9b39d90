"""CPU: the C-ABI library loads, exports every symbol include/lgar_b200.h declares, the ctypes
structs match the C layout, and compute entry points fail loudly without a GPU (no fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT
import lgar_b200
from lgar_b200 import _capi

HEADER = os.path.join(ROOT, "include", "lgar_b200.h")


def test_library_exports_every_declared_symbol():
    lib = _capi.lib()
    text = open(HEADER).read()
    declared = set(re.findall(r"^\s*(?:int|size_t|double|const char\*)\s+(lgar_\w+)\s*\(", text, flags=re.M))
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert lib.lgar_abi_version() == _capi.ABI_VERSION


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lgar_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(lgar_problem),sizeof(lgar_outputs),offsetof(lgar_problem,alpha),offsetof(lgar_problem,iter_cap),'
                   'offsetof(lgar_outputs,sums));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sp, so, oa, oi, os_ = (int(x) for x in subprocess.check_output([str(exe)]).split())
    assert C.sizeof(_capi.Problem) == sp
    assert C.sizeof(_capi.Outputs) == so
    assert _capi.Problem.alpha.offset == oa
    assert _capi.Problem.iter_cap.offset == oi
    assert _capi.Outputs.sums.offset == os_


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "h.c"
    src.write_text('#include "lgar_b200.h"\nint main(void){return LGAR_ABI_VERSION - 1;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(tmp_path / "h")])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _capi.lib()
    assert lib.lgar_device_check() == -3  # LGAR_E_NO_DEVICE
    assert b"no CUDA device" in lib.lgar_last_error_string() or b"not sm_100" in lib.lgar_last_error_string()
    with pytest.raises(_capi.LGARLibraryError):
        lgar_b200.ColumnEnsemble(theta_r=[[0.06]], theta_e=[[0.45]], thickness=[[44.0]],
                                 forcing=[[0.0, 0.0]], device="cpu")
    p = _capi.Problem()
    p.abi_version = _capi.ABI_VERSION
    p.num_columns, p.num_layers, p.num_steps, p.num_subcycles, p.num_sites, p.nint, p.num_giuh = 4, 3, 8, 1, 1, 120, 5
    assert lib.lgar_workspace_bytes(C.byref(p), 0) > 0
    assert lib.lgar_workspace_bytes(C.byref(p), 1) > lib.lgar_workspace_bytes(C.byref(p), 0)
    o = _capi.Outputs()
    dummy = (C.c_double * 64)()
    for f in ("alpha", "n", "ksat", "theta_r", "theta_e", "thickness", "initial_psi", "ponded_depth_max", "forcing"):
        setattr(p, f, C.addressof(dummy))
    rc = lib.lgar_forward_host(C.byref(p), C.byref(o))
    assert rc == -3, "compute entry point must fail without a GPU"
    assert lib.lgar_measure_fp64_flops(16) == 0.0


def test_invalid_arguments_are_reported():
    lib = _capi.lib()
    p = _capi.Problem()
    p.abi_version = 99
    assert lib.lgar_workspace_bytes(C.byref(p), 0) == 0
    assert b"abi_version" in lib.lgar_last_error_string()
    p.abi_version = _capi.ABI_VERSION
    p.num_columns, p.num_layers, p.num_steps, p.num_subcycles, p.nint, p.num_giuh = 1, 9, 1, 1, 120, 5
    assert lib.lgar_workspace_bytes(C.byref(p), 0) == 0
    assert b"num_layers" in lib.lgar_last_error_string()
