"""Experiment: how much would placing columns by their (future) crash step / by a dynamic work measure gain?
usage: python tools/exp_order.py [T]"""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lgar_b200
from lgar_b200 import workloads, ColumnEnsemble, forward_raw
B = 125000; T = int(sys.argv[1]) if len(sys.argv) > 1 else 8760
we = workloads.synthetic_sites_ensemble(B=B, T=T, sites=128, rank=0)
ens = ColumnEnsemble(theta_r=we.theta_r, theta_e=we.theta_e, thickness=we.thickness, forcing=we.forcing, site_index=we.site_index)
ens.balance(we.ksat)
def run(tag, nf=False):
    torch.cuda.synchronize()
    res, ws = forward_raw(ens, we.alpha, we.n, we.ksat, outputs=("runoff", "AET"), num_fronts=nf)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); res, ws = forward_raw(ens, we.alpha, we.n, we.ksat, outputs=("runoff", "AET"), workspace=ws, num_fronts=nf); e1.record()
    torch.cuda.synchronize()
    st = res.status.cpu().numpy(); cr = res.crash_step.cpu().numpy()
    alive = np.where(st == 0, T, np.maximum(cr, 0))
    ms = e0.elapsed_time(e1)
    print(f"{tag}: {ms:.0f} ms -> {alive.sum() / ms * 1e3:.4g} col-steps/s", flush=True)
    return res, alive
res, alive = run("balanced (site, ksat0)", nf=True)
nfm = res.num_fronts.double().mean(dim=0).cpu().numpy()   # mean front count of every column over the record
site = we.site_index.astype(np.int64)
def set_order(keys):
    order = np.lexsort(keys[::-1])
    ens.column_order = torch.as_tensor(order.astype(np.int32), device=ens.device)
set_order((site, alive, we.ksat[0])); run("(site, alive steps, ksat0)  [uses the future]")
set_order((-alive, site, we.ksat[0])); run("(alive steps desc, site, ksat0)  [uses the future]")
set_order((site, np.round(nfm * 2), we.ksat[0])); run("(site, mean front count, ksat0)  [uses the future]")
set_order((site, alive, np.round(nfm * 2), we.ksat[0])); run("(site, alive, mean front count, ksat0)  [uses the future]")
