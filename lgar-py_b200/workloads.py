"""Synthetic ensembles of the shapes BASELINE.json names.

Pure numpy (no torch, no CUDA) so that the same generator feeds the CUDA path, the CPU oracle
and the tests.  The two base forcing records (Phillipsburg / Bushland, 8760 h, cm/h) are read
from the committed fixtures tests/golden/{phil,bush}_year.npz -- /root/reference does not
exist on the GPU box.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_GOLD = os.path.join(_ROOT, "tests", "golden")

# soils table rows 12-14 (P-1..3) and 15-17 (B-1..3) of data/vG_default_params.dat: theta_r, theta_e
SOILS = {
    "phil": (np.array([0.0648, 0.0831, 0.0668]), np.array([0.4513, 0.4773, 0.4617])),
    "bush": (np.array([0.0649, 0.0672, 0.0823]), np.array([0.4481, 0.4760, 0.4782])),
}
THICKNESS = np.array([44.0, 131.0, 25.0])


def base_forcing(site: str) -> np.ndarray:
    """[8760,2] (P, PET) in cm/h."""
    return np.load(os.path.join(_GOLD, f"{site}_year.npz"))["forcing"].astype(np.float64)


@dataclass
class Ensemble:
    alpha: np.ndarray       # [L,B]
    n: np.ndarray           # [L,B]
    ksat: np.ndarray        # [L,B]
    theta_r: np.ndarray     # [L,B]
    theta_e: np.ndarray     # [L,B]
    thickness: np.ndarray   # [L,B]
    forcing: np.ndarray     # [sites,T,2]
    site_index: np.ndarray  # [B] int32
    name: str = ""

    @property
    def num_columns(self):
        return self.alpha.shape[1]


def sample_parameters(rng: np.random.Generator, B: int, L: int = 3):
    """alpha ~ U[0.0015,0.015], n ~ U[1.1,3.0], Ks ~ logU[0.01,30] per layer: the parameter bounds
    of dpLGAR/models/config/*.yaml:18-32 (n's lower bound 1.0 is singular: m -> 0)."""
    alpha = rng.uniform(0.0015, 0.015, size=(L, B))
    n = rng.uniform(1.1, 3.0, size=(L, B))
    ks = np.exp(rng.uniform(np.log(0.01), np.log(30.0), size=(L, B)))
    return alpha, n, ks


def bushland_ensemble(B: int = 100_000, T: int = 8760, seed: int = 0) -> Ensemble:
    """C3: one shared Bushland record, B parameter members."""
    rng = np.random.default_rng(seed)
    a, n, k = sample_parameters(rng, B)
    thr, the = SOILS["bush"]
    rep = lambda v: np.repeat(v.reshape(-1, 1), B, axis=1)
    return Ensemble(a, n, k, rep(thr), rep(the), rep(THICKNESS), base_forcing("bush")[None, :T].copy(),
                    np.zeros(B, dtype=np.int32), name=f"bushland_ensemble_B{B}_T{T}")


def synthetic_sites_ensemble(B: int = 125_000, T: int = 8760, sites: int = 128, seed: int = 1,
                             forcing_seed: int = 1234, rank: int = 0, shared_sites: bool = False) -> Ensemble:
    """C4 shard for one GPU: `sites` synthetic site records (Phillipsburg / Bushland alternating,
    24 h-block log-normal storm scaling sigma = 0.3, circular shift by whole days), each shared by
    ~B/sites parameter members.  Rank r of an N-GPU job draws shard r of the 1M-column ensemble.

    shared_sites=False: every rank also draws its own `sites` records (N x sites sites in the job).
    shared_sites=True (bench.py): the job has `sites` sites in total and rank r holds ITS members of every
    site (own parameter draws, the site records of rank 0), so the shards are statistically equivalent --
    the weak-scaling shape.  With rank-specific records the pass times of two shards differed by 9 % at equal
    clocks and column-steps (profiles/bench_r2_n2_rank_sites.json: 19.6 s and 21.4 s), which reads as a
    scaling loss although no GPU waits for another.  Rank 0's shard is the same either way."""
    rng = np.random.default_rng([seed, rank])
    frng = np.random.default_rng([forcing_seed, 0 if shared_sites else rank])
    a, n, k = sample_parameters(rng, B)
    base = {s: base_forcing(s) for s in ("phil", "bush")}
    full_T = base["phil"].shape[0]
    forcing = np.empty((sites, T, 2))
    kinds = []
    for s in range(sites):
        kind = "phil" if s % 2 == 0 else "bush"
        kinds.append(kind)
        f = np.roll(base[kind], 24 * int(frng.integers(0, 365)), axis=0).copy()
        scale = np.exp(0.3 * frng.standard_normal(full_T // 24 + 1))
        f[:, 0] *= np.repeat(scale, 24)[:full_T]
        forcing[s] = f[:T]
    site_index = (np.arange(B) * sites // B).astype(np.int32)  # contiguous blocks of members per site
    thr = np.empty((3, B)); the = np.empty((3, B))
    for s in range(sites):
        m = site_index == s
        thr[:, m] = SOILS[kinds[s]][0].reshape(3, 1)
        the[:, m] = SOILS[kinds[s]][1].reshape(3, 1)
    thick = np.repeat(THICKNESS.reshape(3, 1), B, axis=1)
    return Ensemble(a, n, k, thr, the, thick, forcing, site_index,
                    name=f"synthetic_sites_B{B}_T{T}_sites{sites}_rank{rank}")


def algorithmic_flops(counters) -> float:
    """FP64 work from counted closure calls, convention of SURVEY.md 8(d): pow = 125 flop
    (41 DFMA + 34 DADD + 9 DMUL in libdevice's pow), div = sqrt = 15.
    counters: [geff, theta_from_h, h_from_se, k_from_se, se_from_h, root iters, colmass iters, substeps]."""
    c = [float(x) for x in counters]
    F_POW, F_DIV, F_SQRT = 125.0, 15.0, 15.0
    theta_h = 2 * F_POW + F_DIV + 4
    h_se = 2 * F_POW + 2 * F_DIV + 3
    k_se = 2 * F_POW + F_SQRT + 6
    se_h = 2 * F_POW + F_DIV + 2
    return (c[1] * theta_h + c[2] * h_se + c[3] * k_se + c[4] * se_h + c[0] * (120 * 4 + 40)
            + c[5] * 12 + c[6] * 40 + c[7] * 150)
