"""Helpers shared by the -m gpu tests: run golden cases through the C ABI (CUDA path)."""
import numpy as np
import torch

import lgar_b200
from lgar_b200 import ColumnEnsemble, forward_raw, OUT_NAMES


def ensemble_from_golden(g, copies=1, **over):
    L = len(g["alpha"])
    B = copies
    rep = lambda v: np.repeat(np.asarray(v, dtype=np.float64).reshape(L, 1), B, axis=1)
    kw = dict(
        theta_r=rep(g["theta_r"]), theta_e=rep(g["theta_e"]), thickness=rep(g["layer_thickness"]),
        forcing=np.asarray(g["forcing"]), initial_psi=float(g["initial_psi"]),
        ponded_depth_max=float(g["ponded_depth_max"]), subcycle_length_h=float(g["subcycle_length_h"]),
        num_subcycles=int(g["num_subcycles"]), nint=int(g["nint"]),
        wilting_point_psi=float(g["wilting_point_psi"]), frozen_factor=float(g["frozen_factor"]),
        giuh_ordinates=tuple(float(x) for x in g["giuh_ordinates"]),
        use_closed_form_G=bool(g["use_closed_form_G"]) if "use_closed_form_G" in g else False)
    kw.update(over)
    ens = ColumnEnsemble(**kw)
    return ens, rep(g["alpha"]), rep(g["n"]), rep(g["ksat"])


def run_golden_on_gpu(g, copies=1, dump=True, **over):
    ens, a, n, k = ensemble_from_golden(g, copies=copies, **over)
    res, _ = forward_raw(ens, a, n, k, outputs=OUT_NAMES, dump_fronts=dump, num_fronts=True)
    torch.cuda.synchronize()
    out = {name: res[name].cpu().numpy() for name in OUT_NAMES}  # [T,B]
    out["status"] = res.status.cpu().numpy()
    out["crash_step"] = res.crash_step.cpu().numpy()
    out["nfronts"] = res.num_fronts.cpu().numpy()
    out["start_volume"] = res.start_volume.cpu().numpy()
    out["sums"] = res.sums.cpu().numpy()
    if dump:
        out["fronts"] = res.fronts.cpu().numpy()             # [T,16,5,B]
        out["front_layer"] = res.front_layer.cpu().numpy()   # [T,16,B]
        out["front_to_bottom"] = res.front_to_bottom.cpu().numpy()
        out["counters"] = res.counters.cpu().numpy()
    return out
