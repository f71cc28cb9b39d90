"""CPU: synthetic ensemble generator (shapes of BASELINE.json configs C3/C4) and the FLOP model."""
import numpy as np

import lgar_b200
from lgar_b200 import workloads


def test_c4_shard_shapes_and_determinism():
    a = workloads.synthetic_sites_ensemble(B=1000, T=240, sites=8, rank=0)
    b = workloads.synthetic_sites_ensemble(B=1000, T=240, sites=8, rank=0)
    c = workloads.synthetic_sites_ensemble(B=1000, T=240, sites=8, rank=1)
    assert a.alpha.shape == (3, 1000) and a.forcing.shape == (8, 240, 2) and a.site_index.shape == (1000,)
    np.testing.assert_array_equal(a.alpha, b.alpha)
    np.testing.assert_array_equal(a.forcing, b.forcing)
    assert not np.array_equal(a.alpha, c.alpha)          # every rank draws its own shard
    assert a.site_index.min() == 0 and a.site_index.max() == 7
    assert (np.diff(a.site_index) >= 0).all()           # members of one site are contiguous
    assert (a.alpha >= 0.0015).all() and (a.alpha <= 0.015).all()
    assert (a.n >= 1.1).all() and (a.n <= 3.0).all()
    assert (a.ksat >= 0.01).all() and (a.ksat <= 30.0).all()
    assert (a.forcing >= 0).all()


def test_shared_site_records_give_equivalent_shards():
    """bench.py's weak-scaling shape: every rank holds its own members of the SAME site records; rank 0's shard does not
    depend on the switch (the single-GPU bench lines of all rounds stay comparable)."""
    r0 = workloads.synthetic_sites_ensemble(B=640, T=120, sites=4, rank=0)
    s0 = workloads.synthetic_sites_ensemble(B=640, T=120, sites=4, rank=0, shared_sites=True)
    s3 = workloads.synthetic_sites_ensemble(B=640, T=120, sites=4, rank=3, shared_sites=True)
    r3 = workloads.synthetic_sites_ensemble(B=640, T=120, sites=4, rank=3)
    for k in ("alpha", "n", "ksat", "forcing", "theta_r", "theta_e", "site_index"):
        np.testing.assert_array_equal(getattr(r0, k), getattr(s0, k))
    np.testing.assert_array_equal(s3.forcing, s0.forcing)       # same sites ...
    np.testing.assert_array_equal(s3.alpha, r3.alpha)           # ... the rank's own parameter members
    assert not np.array_equal(s3.alpha, s0.alpha)
    assert not np.array_equal(r3.forcing, r0.forcing)           # (the A/B variant: own site records per rank)


def test_c3_bushland_ensemble():
    e = workloads.bushland_ensemble(B=64, T=100)
    assert e.forcing.shape == (1, 100, 2) and (e.site_index == 0).all()
    np.testing.assert_allclose(e.theta_e[:, 0], [0.4481, 0.4760, 0.4782])


def test_flop_model_matches_survey_convention():
    # one Geff call = 120 se_from_h + 121 k_from_se + 2 h_from_se + overhead ~ 65 kFLOP
    f = workloads.algorithmic_flops([1, 0, 2, 121, 120, 0, 0, 0])
    assert 6.0e4 < f < 7.2e4
