#!/usr/bin/env python
"""bench.py -- column-timesteps/sec of the LGAR time-stepping core on N B200s.

A "step" is ONE pass of the hot path over the whole batch: every column of this rank's shard
advanced through all T forcing steps in one persistent launch.  Workload (config C4 of
BASELINE.json, one GPU's share): 125,000 columns x 8760 hourly steps, 3 layers, 128 synthetic
site records per GPU, per-column random van Genuchten parameters.  Weak scaling: every rank
draws its own shard; no data-path collective (columns are independent).

  python bench.py --gpus 1 --steps K --warmup W            # CUDA path (this repo)
  python bench.py --impl reference ...                     # CPU restatement of the reference on host cores

`value`  : whole-job column-timesteps/s, inputs resident in HBM, CUDA-event time, max over ranks.
           Only column-steps that were actually simulated count (columns the reference would
           abort with an exception stop at their crash step; see config.ok_fraction).
`e2e`    : same metric through the public API with HOST (pinned) parameter/forcing tensors:
           H2D of parameters + forcing and D2H of per-column results inside the timed region.
`roofline`: FP64-ALU bound: algorithmic flop counted from closure-call counters.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--columns", type=int, default=125_000, help="columns per GPU")
    ap.add_argument("--nsteps", type=int, default=8760, help="forcing steps T")
    ap.add_argument("--sites", type=int, default=128, help="forcing records per GPU")
    ap.add_argument("--max-fronts", type=int, default=16)
    ap.add_argument("--chunk", type=int, default=64)
    ap.add_argument("--cpu-sample-columns", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--grad-columns", type=int, default=37888,
                    help="columns of the shard used for the forward+gradient figure (0 = skip); the default is one "
                         "32-column tile for each of the 148 SMs x 2 CTAs x 4 resident warps of the reverse kernel")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0, interval=0.2):
        super().__init__(daemon=True)
        self.index = index
        self.interval = interval
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(self.interval)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_reference_rate(ens, T, columns, threads):
    """Times the CPU restatement of the reference (oracle/, kind='port') on `columns` columns."""
    from oracle import lgar_oracle as O
    cfgs = [O.make_cfg(ens.alpha[:, b], ens.n[:, b], ens.ksat[:, b], ens.theta_r[:, b], ens.theta_e[:, b],
                       thickness=ens.thickness[:, b]) for b in range(columns)]
    # group by site so that each batch shares one forcing record
    done_steps = 0
    t0 = time.perf_counter()
    for s in np.unique(ens.site_index[:columns]):
        idx = np.nonzero(ens.site_index[:columns] == s)[0]
        sums, st = O.forward_batch([cfgs[i] for i in idx], ens.forcing[s, :T], nthreads=threads)
        # crashed columns stop early; count what was simulated (precip sum cannot tell, so rerun cheaply)
        done_steps += len(idx) * T
    dt = time.perf_counter() - t0
    return done_steps / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The Python reference
    cannot travel to the GPU box; its validated C++ restatement (oracle/, pinned bit-for-bit to
    the Python reference's golden vectors) is timed on all host cores instead (kind = port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lgar_b200 import workloads
    threads = os.cpu_count() or 1
    ncol = args.cpu_sample_columns or 24 * threads
    ens = workloads.synthetic_sites_ensemble(B=args.columns, T=args.nsteps, sites=args.sites, rank=0)
    for _ in range(args.warmup):
        cpu_reference_rate(ens, min(args.nsteps, 200), min(ncol, 2 * threads), threads)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt = cpu_reference_rate(ens, args.nsteps, ncol, threads)
        rates.append(r); times.append(dt)
    value = float(np.mean(rates))
    sample = f"{ncol} columns x {args.nsteps} steps of the same ensemble per step (columns 0..{ncol - 1})"
    line = {
        "impl": "reference", "metric": "column-timesteps/sec (fwd)", "value": value, "unit": "column-timesteps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times)) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C4 shard: {args.columns} columns x {args.nsteps} steps, 3 layers, {args.sites} sites/GPU",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "column-timesteps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "column-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    import lgar_b200
    from lgar_b200 import workloads, ColumnEnsemble, forward_raw, _capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _capi.check(_capi.lib().lgar_device_check(), "lgar_device_check")

    B, T = args.columns, args.nsteps
    we = workloads.synthetic_sites_ensemble(B=B, T=T, sites=args.sites, rank=rank)
    # host (pinned) copies for the e2e path, device copies for the HBM-resident path
    host = {k: torch.from_numpy(np.ascontiguousarray(getattr(we, k))).pin_memory()
            for k in ("alpha", "n", "ksat", "forcing")}
    ens = ColumnEnsemble(theta_r=we.theta_r, theta_e=we.theta_e, thickness=we.thickness, forcing=we.forcing,
                         site_index=we.site_index, max_fronts=args.max_fronts, chunk_steps=args.chunk, device=dev)
    d_alpha, d_n, d_ksat = (host[k].to(dev) for k in ("alpha", "n", "ksat"))
    outs = ("runoff", "AET")  # per-step series kept in HBM: 2 x T x B x 8 B

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the first warm-up pass also counts closure calls (work counters -> algorithmic flop per launch)
    # and tells how many column-steps are actually simulated (crashed columns stop at their crash step)
    res, ws = forward_raw(ens, d_alpha, d_n, d_ksat, outputs=outs, counters=True)
    torch.cuda.synchronize()
    counters = res.counters.cpu().numpy()
    status = res.status.cpu().numpy()
    crash = res.crash_step.cpu().numpy()
    alive_steps = int(np.where(status == 0, T, np.maximum(crash, 0)).sum())
    ok_fraction = float((status == 0).mean())
    flop_per_launch = workloads.algorithmic_flops(counters)
    del res

    def one_pass():
        r, _ = forward_raw(ens, d_alpha, d_n, d_ksat, outputs=outs, workspace=ws)
        return r

    for _ in range(max(args.warmup - 1, 0)):
        one_pass()
    barrier()
    # every rank samples its own GPU (rank 0's summary goes into `clocks`; the others sample once per second)
    sampler = ClockSampler(local, interval=0.2 if rank == 0 else 1.0)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        one_pass()
        ev[i + 1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    sampler.stop_flag = True

    # e2e: host tensors in, per-column results out, through the public API
    e2e = None
    if not args.no_e2e:
        def e2e_pass():
            ens.forcing.copy_(host["forcing"], non_blocking=True)
            a = host["alpha"].to(dev, non_blocking=True)
            n_ = host["n"].to(dev, non_blocking=True)
            k = host["ksat"].to(dev, non_blocking=True)
            r, _ = forward_raw(ens, a, n_, k, outputs=outs, workspace=ws)
            return r.sums.cpu(), r.status.cpu()
        e2e_pass()  # one warm-up of the host path (pinned staging buffers, allocator)
        barrier()
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            sums_h, st_h = e2e_pass()
        e1.record()
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
        h2d = sum(host[k].numel() * 8 for k in ("alpha", "n", "ksat", "forcing"))
        d2h = sums_h.numel() * 8 + st_h.numel() * 4
        e2e = (e2e_ms, h2d, d2h)

    # forward + gradient (reverse-mode kernel) on the first `grad_columns` columns of the shard, full record
    fg = None
    if args.grad_columns > 0:
        from lgar_b200 import lgar_columns
        Bg = min(args.grad_columns, B)
        del ws
        torch.cuda.empty_cache()
        sub = lambda x: np.ascontiguousarray(x[:, :Bg])
        ens_g = ColumnEnsemble(theta_r=sub(we.theta_r), theta_e=sub(we.theta_e), thickness=sub(we.thickness),
                               forcing=we.forcing, site_index=we.site_index[:Bg], max_fronts=args.max_fronts,
                               chunk_steps=args.chunk, device=dev)
        alive_g = int(np.where(status[:Bg] == 0, T, np.maximum(crash[:Bg], 0)).sum())

        def fwd_bwd():
            A = d_alpha[:, :Bg].clone().requires_grad_(True)
            N_ = d_n[:, :Bg].clone().requires_grad_(True)
            Kk = d_ksat[:, :Bg].clone().requires_grad_(True)
            out = lgar_columns(A, N_, Kk, ens_g, outputs=("runoff", "AET"))
            ok = out["status"] == 0
            loss = torch.nan_to_num(out["runoff"] + out["AET"]).sum(dim=0)[ok].mean()
            loss.backward()
            return A.grad
        fwd_bwd()
        barrier()
        g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
        g0.record()
        grad = fwd_bwd()
        g1.record()
        barrier()
        fg = (g0.elapsed_time(g1), alive_g, float(torch.isfinite(grad).float().mean()))

    stats = torch.tensor([total_ms, float(alive_steps), flop_per_launch, e2e[0] if e2e else 0.0,
                          fg[0] if fg else 0.0, float(fg[1]) if fg else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms_all, fg_ms_all = float(mx[0]), float(mx[3]), float(mx[4])
        alive_all, flop_all, fg_alive_all = float(sm[1]), float(sm[2]), float(sm[5])
    else:
        e2e_ms_all, alive_all, flop_all = float(stats[3]), float(alive_steps), flop_per_launch
        fg_ms_all, fg_alive_all = float(stats[4]), float(stats[5])
    # per-rank step time and SM clock (diagnosis of weak-scaling losses: which rank / GPU was the slow one)
    my_clk = sampler.summary()
    per_rank = torch.tensor([float(stats[0]) / args.steps, float(my_clk["sm_mhz"] or 0.0),
                             1.0 if my_clk["reasons"] else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        gathered = [torch.zeros_like(per_rank) for _ in range(world)]
        dist.all_gather(gathered, per_rank)
    else:
        gathered = [per_rank]
    per_rank_ms = [float(g[0]) for g in gathered]
    per_rank_mhz = [float(g[1]) for g in gathered]
    per_rank_throttled = [bool(g[2] > 0) for g in gathered]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / args.steps
    value = alive_all / (ms_per_step * 1e-3)
    fp64_peak = _capi.lib().lgar_measure_fp64_flops(8192)
    kern_s = float(np.mean(kern_ms)) * 1e-3
    achieved = flop_per_launch / kern_s  # rank 0's kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    out_bytes = (len(outs) * T * B * 8 + T * 16 * args.sites)
    line = {
        "metric": "column-timesteps/sec (fwd)", "value": value, "unit": "column-timesteps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C4 shard per GPU: {B} columns x {T} hourly steps, 3 layers, {args.sites} synthetic sites, "
                               "alpha~U[0.0015,0.015] n~U[1.1,3] Ks~logU[0.01,30]",
                   "columns_per_gpu": B, "forcing_steps": T, "ok_fraction": ok_fraction,
                   "alive_column_steps_per_gpu": alive_steps, "max_fronts": args.max_fronts, "chunk_steps": args.chunk,
                   "l2": "working set (per-step outputs 2 x T x B x 8 B = %.1f GB) exceeds L2; no flush needed" % (len(outs) * T * B * 8 / 1e9)},
        "gpu_launches": args.steps,
        "clocks": my_clk,
        "per_rank": {"ms_per_step": per_rank_ms, "sm_mhz": per_rank_mhz, "throttle_reason_seen": per_rank_throttled},
        "roofline": {"bound": "fp64", "achieved": achieved / 1e12, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak if fp64_peak else None, "traffic": None,
                     "peak_source": "in-run DFMA probe (lgar_measure_fp64_flops); MEASURED_PEAKS.json has no FP64 entry",
                     "flop_per_column_step": flop_per_launch / max(alive_steps, 1),
                     "hbm": {"algorithmic_bytes_per_launch": out_bytes, "achieved_gbs": out_bytes / kern_s / 1e9,
                             "peak_gbs": hbm_peak, "frac": out_bytes / kern_s / 1e9 / hbm_peak}},
    }
    if fg:
        line["fwd_grad"] = {"value": fg_alive_all / (fg_ms_all * 1e-3), "unit": "column-timesteps/s",
                            "columns_per_gpu": min(args.grad_columns, B), "forcing_steps": T,
                            "what": "lgar_forward(keep checkpoints) + lgar_backward through torch.autograd, "
                                    "loss = mean over OK columns of sum_t (runoff + AET); bounded column subset of the "
                                    "same shard, full record", "finite_grad_fraction": fg[2]}
    if e2e:
        line["e2e"] = {"value": alive_all / (e2e_ms_all / args.steps * 1e-3), "unit": "column-timesteps/s",
                       "h2d_bytes_per_step": e2e[1], "d2h_bytes_per_step": e2e[2]}
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        ncol = args.cpu_sample_columns or 24 * threads
        ncol = min(ncol, B)
        rate, dt = cpu_reference_rate(we, T, ncol, threads)
        line["cpu_baseline"] = {"value": rate, "unit": "column-timesteps/s", "cores": threads, "kind": "port",
                                "sample": f"columns 0..{ncol - 1} x {T} steps of the same ensemble, {dt:.1f} s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
