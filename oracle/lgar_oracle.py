"""ctypes binding of the CPU oracle (oracle/lgar_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblgar_oracle.so")

LMAX, FMAX, NOUT, NGIUH = 8, 16, 10, 8
OUT_NAMES = ("runoff", "percolation", "AET", "infiltration", "ending_volume",
             "ponded_water", "giuh_runoff", "precip", "PET", "discharge")
STATUS_NAMES = ("OK", "NEG_POW", "NAN", "THETA_ORDER", "BOTTOM_REACHED", "NULL_NEIGHBOUR",
                "FRONT_OVERFLOW", "ITER_CAP", "INDEX_ERROR")


class OracleCfg(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int32), ("nint", C.c_int32), ("num_subcycles", C.c_int32),
        ("num_giuh", C.c_int32),
        ("dt_h", C.c_double), ("initial_psi", C.c_double), ("wilting_point_psi", C.c_double),
        ("ponded_depth_max", C.c_double), ("frozen_factor", C.c_double),
        ("thickness", C.c_double * LMAX), ("theta_r", C.c_double * LMAX),
        ("theta_e", C.c_double * LMAX), ("alpha", C.c_double * LMAX), ("n", C.c_double * LMAX),
        ("ksat", C.c_double * LMAX), ("giuh", C.c_double * NGIUH), ("iter_cap", C.c_int64),
        ("use_closed_form_G", C.c_int64),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lgar_oracle.cpp")
    dev = os.path.join(_HERE, "liblgar_oracle_devpow.so")
    if (force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src)
            or not os.path.exists(dev) or os.path.getmtime(dev) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_libs = {}
_variant = "glibc"


class device_pow:
    """Context manager: inside it every oracle call runs liblgar_oracle_devpow.so, the same restatement with the CUDA
    path's table-driven pow (lgar-py_b200/csrc/lgar_pow.cuh, compiled for the host) instead of glibc's.  Purpose: the two
    pows differ in the last bit in ~0.035 % of calls, which an ill-conditioned column can amplify into another
    trajectory; with the same pow on both sides any remaining difference is a logic defect of the CUDA path."""

    def __enter__(self):
        global _variant
        self._old, _variant = _variant, "devpow"
        return self

    def __exit__(self, *exc):
        global _variant
        _variant = self._old
        return False


def lib():
    if _variant not in _libs:
        build()
        so = _SO if _variant == "glibc" else os.path.join(_HERE, "liblgar_oracle_devpow.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        l = C.CDLL(so)
        assert l.lgar_oracle_sizeof_cfg() == C.sizeof(OracleCfg)
        l.lgar_oracle_forward.restype = C.c_int
        l.lgar_oracle_forward_tangent.restype = C.c_int
        l.lgar_oracle_forward_batch.restype = C.c_int
        l.lgar_oracle_forward_batch_ex.restype = C.c_int
        _libs[_variant] = l
    return _libs[_variant]


def make_cfg(alpha, n, ksat, theta_r, theta_e, thickness=(44.0, 131.0, 25.0), dt_h=1.0,
             num_subcycles=1, initial_psi=2000.0, wilting_point_psi=15495.0, ponded_depth_max=0.0,
             frozen_factor=1.0, nint=120, giuh=(0.06, 0.51, 0.28, 0.12, 0.03), iter_cap=0,
             use_closed_form_G=False) -> OracleCfg:
    c = OracleCfg()
    L = len(alpha)
    c.num_layers, c.nint, c.num_subcycles, c.num_giuh = L, int(nint), int(num_subcycles), len(giuh)
    c.dt_h, c.initial_psi, c.wilting_point_psi = float(dt_h), float(initial_psi), float(wilting_point_psi)
    c.ponded_depth_max, c.frozen_factor, c.iter_cap = float(ponded_depth_max), float(frozen_factor), int(iter_cap)
    c.use_closed_form_G = 1 if use_closed_form_G else 0
    for l in range(L):
        c.thickness[l] = float(thickness[l]); c.theta_r[l] = float(theta_r[l]); c.theta_e[l] = float(theta_e[l])
        c.alpha[l] = float(alpha[l]); c.n[l] = float(n[l]); c.ksat[l] = float(ksat[l])
    for i, g in enumerate(giuh):
        c.giuh[i] = float(g)
    return c


def cfg_from_golden(g, **over) -> OracleCfg:
    """Build the oracle configuration from a tests/golden/*.npz record."""
    kw = dict(
        alpha=g["alpha"], n=g["n"], ksat=g["ksat"], theta_r=g["theta_r"], theta_e=g["theta_e"],
        thickness=g["layer_thickness"], dt_h=float(g["subcycle_length_h"]),
        num_subcycles=int(g["num_subcycles"]), initial_psi=float(g["initial_psi"]),
        wilting_point_psi=float(g["wilting_point_psi"]), ponded_depth_max=float(g["ponded_depth_max"]),
        frozen_factor=float(g["frozen_factor"]), nint=int(g["nint"]), giuh=g["giuh_ordinates"],
        use_closed_form_G=bool(g["use_closed_form_G"]) if "use_closed_form_G" in g else False)
    kw.update(over)
    return make_cfg(**kw)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def forward(cfg: OracleCfg, forcing: np.ndarray, fronts: bool = True):
    """One column through all forcing steps.  Returns dict with per-step outputs
    `out[T,NOUT]`, front dumps, `status`, `crash_step`, `counters`."""
    forcing = np.ascontiguousarray(forcing, dtype=np.float64)
    T = forcing.shape[0]
    out = np.zeros((T, NOUT))
    fr = np.zeros((T, FMAX, 5)) if fronts else None
    fl = np.full((T, FMAX), -1, dtype=np.int8) if fronts else None
    ftb = np.zeros((T, FMAX), dtype=np.int8) if fronts else None
    nf = np.zeros(T, dtype=np.int32)
    crash = C.c_int32(-1)
    cnt = np.zeros(12, dtype=np.int64)  # 0-6 work counters, 8-11 branch coverage (lgar_oracle.cpp g_cnt)
    st = lib().lgar_oracle_forward(
        C.byref(cfg), _p(forcing, C.c_double), C.c_int(T), _p(out, C.c_double), _p(fr, C.c_double),
        _p(fl, C.c_int8), _p(ftb, C.c_int8), _p(nf, C.c_int32), C.byref(crash), _p(cnt, C.c_longlong))
    r = {k: out[:, i] for i, k in enumerate(OUT_NAMES)}
    r.update(out=out, fronts=fr, front_layer=fl, front_to_bottom=ftb, nfronts=nf, status=int(st),
             crash_step=int(crash.value), counters=cnt)
    return r


def forward_tangent(cfg: OracleCfg, forcing: np.ndarray):
    """Forward-mode tangents: dout[T,NOUT,3L] ordered (alpha[L], n[L], ksat[L])."""
    forcing = np.ascontiguousarray(forcing, dtype=np.float64)
    T = forcing.shape[0]
    L = cfg.num_layers
    out = np.zeros((T, NOUT))
    dout = np.zeros((T, NOUT, 3 * L))
    crash = C.c_int32(-1)
    st = lib().lgar_oracle_forward_tangent(
        C.byref(cfg), _p(forcing, C.c_double), C.c_int(T), _p(out, C.c_double), _p(dout, C.c_double),
        C.byref(crash))
    return dict(out=out, dout=dout, status=int(st), crash_step=int(crash.value))


def forward_batch(cfgs, forcing: np.ndarray, nthreads: int = 1):
    """B columns sharing one forcing record; returns (sums[B,NOUT], status[B])."""
    forcing = np.ascontiguousarray(forcing, dtype=np.float64)
    B = len(cfgs)
    arr = (OracleCfg * B)(*cfgs)
    sums = np.zeros((B, NOUT))
    status = np.zeros(B, dtype=np.int32)
    lib().lgar_oracle_forward_batch(arr, C.c_int(B), _p(forcing, C.c_double), C.c_int(forcing.shape[0]),
                                    _p(sums, C.c_double), _p(status, C.c_int32), C.c_int(nthreads))
    return sums, status


def forward_batch_ex(cfgs, forcing: np.ndarray, site=None, nthreads: int = 1, tangents: bool = False):
    """B columns; `forcing` is [T,2] (shared) or [sites,T,2] with `site[B]` choosing the record of each column.
    Returns dict(sums[B,NOUT], status[B], crash_step[B]) and, with tangents=True, dsums[B,NOUT,3L]: forward-mode
    derivatives of the sums w.r.t. (alpha[L], n[L], ksat[L]) with the reference's autograd conventions."""
    forcing = np.ascontiguousarray(forcing, dtype=np.float64)
    T = forcing.shape[-2]
    B = len(cfgs)
    arr = (OracleCfg * B)(*cfgs)
    sums = np.zeros((B, NOUT))
    status = np.zeros(B, dtype=np.int32)
    crash = np.zeros(B, dtype=np.int32)
    L = cfgs[0].num_layers if B else 0
    dsums = np.zeros((B, NOUT, 3 * L)) if tangents else None
    site_arr = None
    if forcing.ndim == 3:
        site_arr = np.ascontiguousarray(site if site is not None else np.zeros(B), dtype=np.int32)
    lib().lgar_oracle_forward_batch_ex(arr, C.c_int(B), _p(forcing, C.c_double), _p(site_arr, C.c_int32), C.c_int(T),
                                       _p(sums, C.c_double), _p(dsums, C.c_double), _p(status, C.c_int32),
                                       _p(crash, C.c_int32), C.c_int(nthreads))
    r = dict(sums=sums, status=status, crash_step=crash)
    if tangents:
        r["dsums"] = dsums
    return r
