"""Run the UNMODIFIED Python reference (/root/reference/dpLGAR) on one soil column.

TEST INFRASTRUCTURE ONLY (oracle/).  This module exists to (a) generate the
golden vectors committed under tests/golden/ and (b) validate the C++ oracle
(oracle/lgar_oracle.cpp).  It reads the reference from /root/reference (build container) or, where that does not exist (the GPU
box), from the staged copy oracle/_ref/ (oracle/make_ref.sh: unmodified files, git-ignored).  The `-m gpu`
tests and smoke() never import it; bench.py's `--impl reference` arm runs oracle/time_reference.py (which
uses this module) in sub-processes to time the real Python reference on the box's host cores.

Recipe:
  * a stub `omegaconf` (oracle/pyref_stub) makes the reference importable;
  * default dtype must be float64 before anything is built
    (reference: dpLGAR/agents/DifferentiableLGAR.py:32);
  * cfg.models.* derived fields as in dpLGAR/agents/DifferentiableLGAR.py:35-52;
  * after every forward() the accumulators are read and then zeroed, which is what
    MassBalance.change_mass does (dpLGAR/models/physics/MassBalance.py:31-53).
"""
from __future__ import annotations

import os
import sys
import traceback

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_ref_root():
    for cand in (os.environ.get("LGAR_REF_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "dpLGAR")):
            return cand
    return "/root/reference"


REF_ROOT = _find_ref_root()
FMAX = 16

OUT_KEYS = (
    "runoff", "percolation", "AET", "infiltration", "ending_volume",
    "ponded_water", "giuh_runoff", "precip", "PET", "discharge",
)


def _import_reference():
    stub = os.path.join(_HERE, "pyref_stub")
    if stub not in sys.path:
        sys.path.insert(0, stub)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import torch

    torch.set_default_dtype(torch.float64)
    from omegaconf import DictConfig
    from dpLGAR.models.dpLGAR import dpLGAR

    return torch, DictConfig, dpLGAR


def build_cfg(
    forcing_file=None,
    soil_params_file=f"{REF_ROOT}/data/vG_default_params.dat",
    layer_thickness=(44.0, 131.0, 25.0),
    layer_soil_type=(12, 13, 14),
    initial_psi=2000.0,
    ponded_depth_max=0.0,
    subcycle_length=3600.0,
    forcing_resolution=3600.0,
    endtime_h=8760.0,
    nint=120,
    frozen_factor=1,
    use_closed_form_G=False,
    wilting_point_psi=15495.0,
    giuh_ordinates=(0.06, 0.51, 0.28, 0.12, 0.03),
):
    _, DictConfig, _ = _import_reference()
    hr_to_sec = 3600.0
    models = dict(
        subcycle_length=subcycle_length,
        forcing_resolution=forcing_resolution,
        endtime=endtime_h,
        hyperparameters=dict(warmup=168, epochs=1, learning_rate=1e-3, minibatch=0.04166666667),
    )
    # derived exactly as dpLGAR/agents/DifferentiableLGAR.py:35-52
    models["endtime_s"] = models["endtime"] * hr_to_sec
    models["subcycle_length_h"] = models["subcycle_length"] * (1 / hr_to_sec)
    models["forcing_resolution_h"] = models["forcing_resolution"] / hr_to_sec
    models["time_per_step"] = models["forcing_resolution_h"] * hr_to_sec
    models["nsteps"] = int(models["endtime_s"] / models["time_per_step"])
    models["num_subcycles"] = int(models["forcing_resolution_h"] / models["subcycle_length_h"])
    cfg = DictConfig(
        device="cpu",
        constants=dict(frozen_factor=frozen_factor, nint=nint),
        conversions=dict(cm_to_mm=10.0, mm_to_cm=0.1, cm_to_m=0.01, hr_to_sec=hr_to_sec),
        data=dict(
            forcing_file=forcing_file,
            soil_params_file=soil_params_file,
            layer_thickness=list(layer_thickness),
            initial_psi=float(initial_psi),
            ponded_depth_max=float(ponded_depth_max),
            use_closed_form_G=use_closed_form_G,
            layer_soil_type=list(layer_soil_type),
            wilting_point_psi=float(wilting_point_psi),
            max_soil_types=25,
            giuh_ordinates=list(giuh_ordinates),
            soil_index=None,
        ),
        models=models,
    )
    return cfg


def read_forcing_cm_per_h(path, nrows=None):
    """[T,2] (P, PET) in cm/h. Handles the `Time` csv and the `#Time` .txt files."""
    import pandas as pd

    df = pd.read_csv(path)
    df.columns = [c.lstrip("#") for c in df.columns]
    x = np.stack([df["P(mm/h)"].values, df["PET(mm/h)"].values], axis=1).astype(np.float64)
    if nrows is not None:
        x = x[:nrows]
    # Data.py:40 does `x_tr * cfg.conversions.mm_to_cm` in float64
    return x * 0.1


def dump_fronts(model):
    """Walk top_layer..next_layer and record every wetting front (Appendix B step 7)."""
    rows = []
    layer = model.top_layer
    while layer is not None:
        for wf in layer.wetting_fronts:
            rows.append(
                (
                    float(wf.depth), float(wf.theta), float(wf.psi_cm), float(wf.k_cm_per_h),
                    float(wf.dzdt), int(wf.layer_num), int(bool(wf.to_bottom)), int(layer.layer_num),
                )
            )
        layer = layer.next_layer
    return rows


def run_reference(
    forcing,                      # ndarray [T,2] cm/h
    cfg_kwargs=None,
    alpha=None, n=None, ksat=None,  # optional per-layer overrides
    record_fronts=True,
    grad_losses=None,             # e.g. ("AET","infiltration","runoff","final_volume")
    verbose=False,
    pdm_leaf=False,               # ponded_depth_max as nn.Parameter: what the commented-out line dpLGAR.py:48 does
):
    """Returns dict of numpy arrays.  Per-step outputs are the model accumulators read
    after each forward() and then zeroed (== MassBalance.change_mass semantics)."""
    torch, _, dpLGAR = _import_reference()
    cfg = build_cfg(**(cfg_kwargs or {}))
    model = dpLGAR(cfg)
    if pdm_leaf:  # set on the instance (the reference source is not touched); set_internal_states() hands a clone of
        # it to GlobalParams (dpLGAR.py:104, GlobalParams.py:22), so autograd reaches it through every use
        model.ponded_depth_max = torch.nn.Parameter(model.ponded_depth_max.detach().clone())
        model.set_internal_states()
    if alpha is not None or n is not None or ksat is not None:
        with torch.no_grad():
            for i in range(len(model.alpha)):
                if alpha is not None:
                    model.alpha[i].fill_(float(alpha[i]))
                if n is not None:
                    model.n[i].fill_(float(n[i]))
                if ksat is not None:
                    model.ksat[i].fill_(float(ksat[i]))
        model.set_internal_states()
    L = len(model.alpha)
    T = forcing.shape[0]
    out = {k: np.zeros(T) for k in OUT_KEYS}
    out["nfronts"] = np.zeros(T, dtype=np.int32)
    out["start_volume"] = float(model.ending_volume)
    out["c"] = model.c.detach().numpy().copy()
    out["alpha"] = np.array([float(a) for a in model.alpha])
    out["n"] = np.array([float(a) for a in model.n])
    out["ksat"] = np.array([float(a) for a in model.ksat])
    out["forcing"] = np.asarray(forcing, dtype=np.float64)
    if record_fronts:
        out["fronts"] = np.zeros((T, FMAX, 5))
        out["front_layer"] = np.full((T, FMAX), -1, dtype=np.int8)
        out["front_to_bottom"] = np.zeros((T, FMAX), dtype=np.int8)
    out["crash_step"] = -1
    out["crash_type"] = ""
    x = torch.tensor(forcing, dtype=torch.float64)
    want_grad = bool(grad_losses)
    series = {k: [] for k in ("AET", "infiltration", "runoff", "percolation")}
    ctx = torch.enable_grad() if want_grad else torch.no_grad()
    zero = lambda: torch.tensor(0.0)
    with ctx:
        for t in range(T):
            try:
                model(x[t])
            except Exception as e:  # the reference uses exceptions as status (Q9-Q11)
                out["crash_step"] = t
                out["crash_type"] = type(e).__name__
                if verbose:
                    traceback.print_exc()
                break
            for k in OUT_KEYS:
                out[k][t] = float(getattr(model, k))
            if want_grad:
                for k in series:
                    series[k].append(getattr(model, k))
            fr = dump_fronts(model)
            out["nfronts"][t] = len(fr)
            if record_fronts:
                for j, r in enumerate(fr[:FMAX]):
                    out["fronts"][t, j] = r[:5]
                    out["front_layer"][t, j] = r[5]
                    out["front_to_bottom"][t, j] = r[6]
            # MassBalance.change_mass: zero the accumulators (MassBalance.py:45-53)
            model.precip = zero(); model.PET = zero(); model.AET = zero()
            model.infiltration = zero(); model.runoff = zero(); model.percolation = zero()
            model.giuh_runoff = zero(); model.discharge = zero()
            model.groundwater_discharge = zero()
        if want_grad and out["crash_step"] < 0:
            params = list(model.alpha) + list(model.n) + list(model.ksat)
            for lname in grad_losses:
                if lname == "final_volume":
                    loss = model.ending_volume
                else:
                    loss = torch.stack([torch.as_tensor(v) for v in series[lname]]).sum()
                if not getattr(loss, "requires_grad", False):
                    g = [None] * len(params)
                else:
                    g = torch.autograd.grad(loss, params, retain_graph=True, allow_unused=True)
                out[f"grad_{lname}"] = np.array(
                    [0.0 if gi is None else float(gi) for gi in g]
                ).reshape(3, L)  # rows: alpha, n, ksat
                out[f"loss_{lname}"] = float(loss)
                if pdm_leaf:
                    gp = torch.autograd.grad(loss, [model.ponded_depth_max], retain_graph=True, allow_unused=True)[0] \
                        if getattr(loss, "requires_grad", False) else None
                    out[f"grad_{lname}_pdm"] = 0.0 if gp is None else float(gp)
    return out


if __name__ == "__main__":
    f = read_forcing_cm_per_h(f"{REF_ROOT}/data/forcing_data_resampled_uniform_Phillipsburg.csv", 80)
    r = run_reference(f, verbose=True)
    print("start", r["start_volume"], "c0", r["c"][0])
    print("sumAET", r["AET"].sum(), "end", r["ending_volume"][-1], "nf", r["nfronts"][-10:])
