"""Generate the golden vectors in tests/golden/*.npz by running the UNMODIFIED Python
reference (/root/reference) through oracle/ref_harness.py.

Only runnable in the build container (needs /root/reference).  The .npz files are the
fixtures that travel; tests never import the reference.

    python tests/golden/make_golden.py            # all cases, 7 worker processes
    python tests/golden/make_golden.py synth_1    # selected cases
"""
from __future__ import annotations

import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

D = "/root/reference/data"
PHIL = f"{D}/forcing_data_resampled_uniform_Phillipsburg.csv"
BUSH = f"{D}/forcing_data_resampled_uniform_Bushland.csv"

# rows 12..14 of the synth .dat tables (alpha, n, Ks) -- Q1: the hard-coded table of
# read_test_params does not contain them, so they are injected as parameter overrides.
SYNTH1 = dict(alpha=[0.036, 0.019, 0.010], n=[1.56, 1.31, 1.23], ksat=[3.12, 0.26, 0.07])
SYNTH2 = dict(alpha=[0.124, 0.036, 0.020], n=[2.28, 1.56, 1.41], ksat=[14.59, 1.04, 0.45])


def random_members(count, seed):
    """C3-style ensemble members: alpha~U[0.0015,0.015], n~U[1.1,3.0],
    Ks~logU[0.01,30] per layer."""
    rng = np.random.default_rng(seed)
    alpha = rng.uniform(0.0015, 0.015, size=(count, 3))
    n = rng.uniform(1.1, 3.0, size=(count, 3))
    ks = np.exp(rng.uniform(np.log(0.01), np.log(30.0), size=(count, 3)))
    return alpha, n, ks


def cases():
    c = {}
    c["phil_0_1000"] = dict(forcing=(PHIL, 0, 1000))
    c["phil_4500_600"] = dict(forcing=(PHIL, 4500, 600))
    c["phil_4500_400_pdm2"] = dict(forcing=(PHIL, 4500, 400), cfg=dict(ponded_depth_max=2.0))
    c["bush_0_1000"] = dict(forcing=(BUSH, 0, 1000), cfg=dict(layer_soil_type=(15, 16, 17)))
    c["bush_5500_600"] = dict(forcing=(BUSH, 5500, 600), cfg=dict(layer_soil_type=(15, 16, 17)))
    s5 = dict(subcycle_length=300.0, forcing_resolution=300.0)
    c["synth_1"] = dict(
        forcing=(f"{D}/forcing_data_synth_1.txt", 0, None),
        cfg=dict(soil_params_file=f"{D}/vG_default_params_synth_1.dat", **s5), **SYNTH1)
    c["synth_2"] = dict(
        forcing=(f"{D}/forcing_data_synth_2.txt", 0, None),
        cfg=dict(soil_params_file=f"{D}/vG_default_params_synth_2.dat", **s5), **SYNTH2)
    c["synth_1_hard"] = dict(forcing=(f"{D}/forcing_data_synth_1.txt", 0, None), cfg=dict(**s5))
    c["thin_crash"] = dict(
        forcing=(f"{D}/forcing_data_synth_0.csv", 0, 40),
        cfg=dict(layer_thickness=(10.0, 10.0, 10.0)))
    c["phil_dt300_200"] = dict(forcing=(PHIL, 0, 200), cfg=dict(subcycle_length=300.0))
    c["phil_dt300_4560_60_pdm2"] = dict(
        forcing=(PHIL, 4560, 60), cfg=dict(subcycle_length=300.0, ponded_depth_max=2.0))
    # random C3-style members on 500 h windows that contain storms
    al, nn, ks = random_members(16, seed=0)
    starts_b = [5500, 5600, 3300, 5500, 5600, 3300, 5500, 5600]
    starts_p = [4500, 4560, 600, 7000, 4500, 4560, 600, 7000]
    for i in range(8):
        c[f"rand_bush_{i}"] = dict(
            forcing=(BUSH, starts_b[i], 500), cfg=dict(layer_soil_type=(15, 16, 17)),
            alpha=al[i], n=nn[i], ksat=ks[i])
    for i in range(8):
        c[f"rand_phil_{i}"] = dict(
            forcing=(PHIL, starts_p[i], 500), alpha=al[8 + i], n=nn[8 + i], ksat=ks[8 + i])
    # gradient goldens: reference autograd (Q14: NOT finite differences)
    G = ("AET", "infiltration", "runoff", "final_volume")
    c["grad_phil_0_300"] = dict(forcing=(PHIL, 0, 300), grad=G)
    c["grad_phil_4550_150"] = dict(forcing=(PHIL, 4550, 150), grad=G)
    c["grad_phil_4550_150_pdm2"] = dict(
        forcing=(PHIL, 4550, 150), grad=G, cfg=dict(ponded_depth_max=2.0))
    c["grad_bush_5540_160"] = dict(
        forcing=(BUSH, 5540, 160), grad=G, cfg=dict(layer_soil_type=(15, 16, 17)))
    c["grad_synth_2"] = dict(
        forcing=(f"{D}/forcing_data_synth_2.txt", 0, None), grad=G,
        cfg=dict(soil_params_file=f"{D}/vG_default_params_synth_2.dat", **s5), **SYNTH2)
    c["grad_rand_phil_1"] = dict(
        forcing=(PHIL, 4560, 120), grad=G, alpha=al[9], n=nn[9], ksat=ks[9])
    c["grad_phil_dt300_60"] = dict(forcing=(PHIL, 40, 60), grad=G, cfg=dict(subcycle_length=300.0))
    # closed-form Green-Ampt capillary drive (cfg.data.use_closed_form_G=True, green_ampt.py:85-98)
    CF = dict(use_closed_form_G=True)
    c["closed_phil_4500_400"] = dict(forcing=(PHIL, 4500, 400), cfg=dict(**CF))
    c["closed_bush_5500_400"] = dict(forcing=(BUSH, 5500, 400), cfg=dict(layer_soil_type=(15, 16, 17), **CF))
    c["closed_rand_phil_3"] = dict(forcing=(PHIL, 7000, 400), cfg=dict(**CF), alpha=al[11], n=nn[11], ksat=ks[11])
    c["closed_rand_bush_2"] = dict(forcing=(BUSH, 3300, 400), cfg=dict(layer_soil_type=(15, 16, 17), **CF),
                                   alpha=al[2], n=nn[2], ksat=ks[2])
    c["closed_phil_dt300_4560_60"] = dict(forcing=(PHIL, 4560, 60), cfg=dict(subcycle_length=300.0, **CF))
    c["grad_closed_phil_4550_150"] = dict(forcing=(PHIL, 4550, 150), grad=G, cfg=dict(**CF))
    c["grad_closed_rand_phil_1"] = dict(forcing=(PHIL, 4560, 120), grad=G, cfg=dict(**CF), alpha=al[9], n=nn[9], ksat=ks[9])
    # columns of the synthetic bench ensemble (lgar_b200.workloads, rank 0, B=16000, T=2560) that reach
    # Layer.wetting_front_cross_domain_boundary (Layer.py:1010-1053) with percolation != 0, and two that
    # die with an AttributeError on a missing neighbour (Q10)
    for col, T_ in ((1620, 620), (573, 800), (449, 1120)):
        c[f"a16_col{col}"] = dict(ens=(col, T_))
    for col, T_ in ((10049, 60), (8074, 160)):
        c[f"null_col{col}"] = dict(ens=(col, T_))
    # Q8: insert_water with ponded_depth_temp == ponded_depth_max EXACTLY falls through both branches (Layer.py:1509-1521).
    # ponded_depth_max is set to the bits of the ponded water that phil_4500_400_pdm2 holds after its first ponding
    # step (row 96): up to that step the state does not depend on ponded_depth_max, so ponded_depth_temp hits it.
    c["phil_4500_400_pdm_eq"] = dict(forcing=(PHIL, 4500, 400), cfg=dict(ponded_depth_max=0.40456950118861856))
    # frozen_factor != 1 (cfg.constants.frozen_factor): scales ksat at construction (dpLGAR.py:57), the K of a new
    # surficial front (Layer.py:1411), the free-drainage ksat of insert_water (:1467) and calc_bottom_sum_f_p (:1545)
    FF = dict(frozen_factor=0.7)
    c["frozen_phil_4500_400"] = dict(forcing=(PHIL, 4500, 400), cfg=dict(**FF))
    c["frozen_bush_5500_400_pdm2"] = dict(forcing=(BUSH, 5500, 400), cfg=dict(layer_soil_type=(15, 16, 17), ponded_depth_max=2.0, **FF))
    c["frozen_ens_col8074"] = dict(ens=(8074, 160), cfg=dict(**FF))      # free-drainage front in layer 2 (Q18) with the factor
    c["frozen_rand_phil_0"] = dict(forcing=(PHIL, 4500, 400), cfg=dict(**FF), alpha=al[8], n=nn[8], ksat=ks[8])
    c["grad_frozen_phil_4550_150"] = dict(forcing=(PHIL, 4550, 150), grad=G, cfg=dict(**FF))
    # ponded_depth_max as a gradient leaf (models/dpLGAR.py:48-49): 0.2 cm binds in this window (ponding
    # reaches 0.4 cm), so runoff and the later infiltration depend on it
    c["grad_pdmleaf_phil_4550_150"] = dict(forcing=(PHIL, 4550, 150), grad=G, cfg=dict(ponded_depth_max=0.2), pdm_leaf=True)
    # torch.min propagates NaN: column 185 of the C4 bench shard (a column whose top front has already been pushed to a
    # negative depth) reaches calc_dry_depth with theta == theta_e exactly: delta_theta = 0, tau = inf, Geff = 0,
    # tau * geff = NaN, and torch.min(cum_thickness, NaN) = NaN (Layer.py:1331-1333); the run dies one step later in
    # error_check.  Found by bench.py's parity block (the oracle's min_ used to drop the NaN).
    c["nan_dry_depth_col185"] = dict(ens_big=(185, 5160), fronts=False)
    # full-year known answers (config[0]); no per-step front dump to keep the files small
    c["phil_year"] = dict(forcing=(PHIL, 0, 8760), fronts=False)
    c["bush_year"] = dict(forcing=(BUSH, 0, 8760), cfg=dict(layer_soil_type=(15, 16, 17)),
                          fronts=False)
    return c


def run_case(name):
    from oracle.ref_harness import read_forcing_cm_per_h, run_reference

    spec = cases()[name]
    if "ens" in spec or "ens_big" in spec:
        import lgar_b200  # noqa: F401  (numpy-only workload generator; no GPU needed)
        from lgar_b200 import workloads
        big = "ens_big" in spec
        col, T_ = spec["ens_big" if big else "ens"]
        we = (workloads.synthetic_sites_ensemble(B=125_000, T=8760, sites=128, rank=0) if big  # the C4 bench shard
              else workloads.synthetic_sites_ensemble(B=16000, T=2560, sites=128, rank=0))
        site = int(we.site_index[col])
        f = we.forcing[site, :T_].copy()
        spec = dict(spec, alpha=we.alpha[:, col], n=we.n[:, col], ksat=we.ksat[:, col],
                    cfg=dict(spec.get("cfg") or {}, layer_soil_type=(12, 13, 14) if site % 2 == 0 else (15, 16, 17)))
        path, start, count = (f"workloads.synthetic_sites_ensemble(B={'125000,T=8760' if big else '16000,T=2560'},sites=128,rank=0)"
                              f"/site{site}/col{col}"), 0, T_
    else:
        path, start, count = spec["forcing"]
        f = read_forcing_cm_per_h(path)
        f = f[start:] if count is None else f[start:start + count]
    t0 = time.time()
    r = run_reference(
        f, cfg_kwargs=spec.get("cfg"), alpha=spec.get("alpha"), n=spec.get("n"),
        ksat=spec.get("ksat"), record_fronts=spec.get("fronts", True),
        grad_losses=spec.get("grad"), pdm_leaf=bool(spec.get("pdm_leaf", False)))
    cfg = dict(layer_thickness=(44.0, 131.0, 25.0), ponded_depth_max=0.0, subcycle_length=3600.0,
               forcing_resolution=3600.0, initial_psi=2000.0, wilting_point_psi=15495.0,
               nint=120, frozen_factor=1.0, giuh_ordinates=(0.06, 0.51, 0.28, 0.12, 0.03), use_closed_form_G=False)
    cfg.update({k: v for k, v in (spec.get("cfg") or {}).items() if k in cfg})
    r["layer_thickness"] = np.array(cfg["layer_thickness"], dtype=np.float64)
    r["theta_r"] = r["c"][:, 0].copy()
    r["theta_e"] = r["c"][:, 1].copy()
    r["ponded_depth_max"] = float(cfg["ponded_depth_max"])
    r["initial_psi"] = float(cfg["initial_psi"])
    r["wilting_point_psi"] = float(cfg["wilting_point_psi"])
    r["nint"] = int(cfg["nint"])
    r["frozen_factor"] = float(cfg["frozen_factor"])
    r["use_closed_form_G"] = bool(cfg["use_closed_form_G"])
    r["giuh_ordinates"] = np.array(cfg["giuh_ordinates"], dtype=np.float64)
    r["subcycle_length_h"] = cfg["subcycle_length"] * (1 / 3600.0)
    r["num_subcycles"] = int((cfg["forcing_resolution"] / 3600.0) / r["subcycle_length_h"])
    r["forcing_source"] = f"{os.path.basename(path)}[{start}:{start}+{count}]"
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **r)
    return name, time.time() - t0, r["crash_step"], r["crash_type"]


if __name__ == "__main__":
    names = sys.argv[1:] or list(cases().keys())
    # long cases first so the pool stays busy
    names.sort(key=lambda s: 0 if "year" in s else 1)
    with Pool(min(7, len(names))) as p:
        for name, dt, cs, ct in p.imap_unordered(run_case, names):
            print(f"{name}: {dt:.0f}s crash={cs} {ct}", flush=True)
