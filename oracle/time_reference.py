"""Time the UNMODIFIED Python reference on ONE column (one process, one thread).

TEST INFRASTRUCTURE ONLY.  Spawned by bench.py's `--impl reference` arm, one process per host core
(SURVEY 8d "Timing the CPU reference beside it"): runs `dpLGAR.forward(x[t])` row by row through the
reference's own module (oracle/ref_harness.py builds the cfg exactly like agents/DifferentiableLGAR.py:35-52),
either under torch.no_grad() (forward figure) or recording the autograd graph and calling backward() on
loss = sum_t (runoff_t + AET_t) (forward+gradient figure, agents/DifferentiableLGAR.py:163).  Only the step
loop (+ backward) is timed; import and model construction are not.  Prints one JSON line.

  python oracle/time_reference.py --spec spec.npz [--grad] [--warm-rows W] [--budget-s 25]
spec.npz: forcing[T,2] cm/h, alpha[L], n[L], ksat[L], soil_rows[L] (rows of data/vG_default_params.dat).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spec", required=True)
    ap.add_argument("--grad", action="store_true")
    ap.add_argument("--warm-rows", type=int, default=0, help="leading forcing rows stepped before the clock starts")
    ap.add_argument("--budget-s", type=float, default=1e9, help="stop stepping after this many seconds")
    args = ap.parse_args()
    from oracle import ref_harness as H
    torch, _, dpLGAR = H._import_reference()
    torch.set_num_threads(1)
    spec = np.load(args.spec)
    forcing = spec["forcing"]
    cfg = H.build_cfg(layer_soil_type=tuple(int(r) for r in spec["soil_rows"]),
                      soil_params_file=os.path.join(H.REF_ROOT, "data", "vG_default_params.dat"))
    model = dpLGAR(cfg)
    with torch.no_grad():
        for i in range(len(model.alpha)):
            model.alpha[i].fill_(float(spec["alpha"][i]))
            model.n[i].fill_(float(spec["n"][i]))
            model.ksat[i].fill_(float(spec["ksat"][i]))
    model.set_internal_states()
    x = torch.tensor(forcing, dtype=torch.float64)
    zero = lambda: torch.tensor(0.0)
    steps, crash, terms = 0, "", []
    t0 = time.perf_counter()
    with (torch.enable_grad() if args.grad else torch.no_grad()):
        for t in range(x.shape[0]):
            if t == args.warm_rows:  # warm-up rows are simulated (and stay in the autograd graph) but not timed
                t0 = time.perf_counter()
                steps = 0
            try:
                model(x[t])
            except Exception as e:  # the reference's exceptions end the column (status codes of the CUDA path)
                crash = type(e).__name__
                break
            steps += 1
            if args.grad:
                terms.append(model.runoff + model.AET)
            sums_t = float(model.runoff) + float(model.AET)  # the per-step read MassBalance.change_mass does
            model.precip = zero(); model.PET = zero(); model.AET = zero(); model.infiltration = zero()
            model.runoff = zero(); model.percolation = zero(); model.giuh_runoff = zero(); model.discharge = zero()
            model.groundwater_discharge = zero()
            if time.perf_counter() - t0 > args.budget_s:
                break
        t_fwd = time.perf_counter() - t0
        grad_ok = None
        if args.grad and steps > 0:
            loss = torch.stack([torch.as_tensor(v) for v in terms]).sum()
            if loss.requires_grad:
                loss.backward()
                grad_ok = all(p.grad is None or bool(torch.isfinite(p.grad)) for p in model.parameters())
    total = time.perf_counter() - t0
    print(json.dumps({"steps": steps, "seconds": total, "forward_seconds": t_fwd, "grad": bool(args.grad),
                      "crash": crash, "grad_finite": grad_ok, "ref_root": H.REF_ROOT}), flush=True)


if __name__ == "__main__":
    main()
