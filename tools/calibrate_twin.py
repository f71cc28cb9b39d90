"""Twin-experiment calibration on N GPUs (BASELINE.json config C5: parameter learning with batched gradient steps and
an NCCL all-reduce of the shared-parameter gradients).

Every rank holds `--sites` forcing records (windows of the synthetic C4 site records).  "Observations" are the runoff
series of a run with the TRUE parameters (the reference's Phillipsburg soils); training starts from perturbed
parameters and uses the reference agent's loop (lgar_b200.agent.DifferentiableLGAR: MSE + RangeBoundLoss, Adam), one
forward + one reverse launch per epoch and ONE all-reduce of [loss, d alpha, d n, d ksat].

  python tools/calibrate_twin.py --sites 32 --epochs 30
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/calibrate_twin.py --sites 32
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=32, help="forcing records per rank")
    ap.add_argument("--start", type=int, default=4300)
    ap.add_argument("--hours", type=int, default=1000)
    ap.add_argument("--epochs", type=int, default=30)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--learn", default="all", choices=["all", "ksat"],
                    help="all: alpha, n, ksat of every layer learnable, top layer perturbed (the reference's set-up); "
                         "ksat: only ksat learnable and perturbed -- the one parameter for which the reference's "
                         "straight-through autograd gradient agrees with finite differences (SURVEY Q14)")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import lgar_b200
    from lgar_b200 import workloads, dpLGAR
    from lgar_b200.agent import DifferentiableLGAR, derive_time_config

    we = workloads.synthetic_sites_ensemble(B=6 * a.sites, T=8760, sites=6 * a.sites, rank=rank)
    x_all = we.forcing[0::2, a.start:a.start + a.hours]  # the Phillipsburg-based records of this rank's shard
    thr, the = workloads.SOILS["phil"]
    cfg = dict(device="cuda", conversions=dict(hr_to_sec=3600.0, mm_to_cm=0.1), constants=dict(frozen_factor=1.0, nint=120),
               data=dict(layer_soil_type=[12, 13, 14], layer_thickness=[44.0, 131.0, 25.0], initial_psi=2000.0,
                         ponded_depth_max=0.0, wilting_point_psi=15495.0, giuh_ordinates=[0.06, 0.51, 0.28, 0.12, 0.03]),
               models=dict(endtime=float(a.hours), subcycle_length=3600.0, forcing_resolution=3600.0,
                           hyperparameters=dict(warmup=24, epochs=a.epochs, learning_rate=a.lr,
                                                lb=[0.0015, 1.0, 1e-6, 0.0], ub=[0.015, 5.0, 30, 10.0])))
    derive_time_config(cfg)
    dev = torch.device("cuda", local)
    # alpha, n, ksat of the top layer (the one runoff is sensitive to)
    perturb = (1.25, 0.95, 1.5) if a.learn == "all" else (1.0, 1.0, 1.5)
    # the reference aborts a run whose wetting front reaches the bottom of the column: keep the records
    # that survive both the true and the perturbed parameters
    probe = dpLGAR(cfg, theta_r=thr, theta_e=the, columns=x_all.shape[0], device=dev)
    with torch.no_grad():
        ok = probe.forward_record(x_all, outputs=("runoff",), on_status="ignore")["status"] == 0
        probe.alpha[0].mul_(perturb[0]); probe.n[0].mul_(perturb[1]); probe.ksat[0].mul_(perturb[2])
        ok &= probe.forward_record(x_all, outputs=("runoff",), on_status="ignore")["status"] == 0
    keep = ok.nonzero().flatten().cpu().numpy()[:a.sites]
    assert len(keep) == a.sites, f"only {len(keep)} usable records"
    x = np.ascontiguousarray(x_all[keep])
    truth = dpLGAR(cfg, theta_r=thr, theta_e=the, columns=a.sites, device=dev)
    true_params = [[float(p) for p in pl] for pl in (truth.alpha, truth.n, truth.ksat)]
    with torch.no_grad():
        y = truth.forward_record(x, outputs=("runoff",), on_status="ignore")["runoff"].transpose(0, 1).cpu()  # [sites, T]
    model = dpLGAR(cfg, theta_r=thr, theta_e=the, columns=a.sites, device=dev)
    with torch.no_grad():  # perturbed start (top layer matters for runoff)
        model.alpha[0].mul_(perturb[0]); model.n[0].mul_(perturb[1]); model.ksat[0].mul_(perturb[2])
    if a.learn == "ksat":
        for pl in (model.alpha, model.n):
            for p_ in pl:
                p_.requires_grad_(False)
    ag = DifferentiableLGAR(cfg, model=model, x=x, y=y.numpy(), device=dev, on_column_error="mask")
    log = []
    t0 = time.time()
    for ep in range(a.epochs):
        ag.train_one_epoch()
        ag.current_epoch += 1
        model.set_internal_states()
        _, loss, nse = ag.history[-1]
        cur = [[float(p) for p in pl] for pl in (model.alpha, model.n, model.ksat)]
        if rank == 0:
            log.append(dict(epoch=ep, loss=loss, nse_rank0=nse, alpha0=cur[0][0], n0=cur[1][0], ksat0=cur[2][0]))
            print(json.dumps(log[-1]), flush=True)
    torch.cuda.synchronize()
    if world > 1:  # every rank must hold the same parameters after the same all-reduced steps
        flat = torch.tensor([v for pl in cur for v in pl], dtype=torch.float64, device=dev)
        mx, mn = flat.clone(), flat.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        assert torch.equal(mx, mn), "ranks diverged"
    if rank == 0:
        print(json.dumps(dict(summary=True, learn=a.learn, lr=a.lr, n_gpus=world, sites_per_gpu=a.sites, hours=a.hours, epochs=a.epochs,
                              seconds=time.time() - t0, column_steps_per_epoch=world * a.sites * a.hours,
                              true=dict(alpha0=true_params[0][0], n0=true_params[1][0], ksat0=true_params[2][0]),
                              first_loss=log[0]["loss"], last_loss=log[-1]["loss"])), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
