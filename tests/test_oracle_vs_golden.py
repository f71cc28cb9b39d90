"""CPU: the C++ oracle (oracle/lgar_oracle.cpp) against the golden vectors generated from the
UNMODIFIED Python reference (tests/golden/make_golden.py).  This is what pins the oracle.

Discrete state (front count, order, layer, to_bottom, exception step) must be identical.
Floating point: the reference's torch.sqrt goes through MKL VML in this torch build and is 1 ulp
low in ~0.6 % of calls (the oracle uses IEEE sqrt), so outputs agree to <= 1e-11 relative
rather than always bit-for-bit; most cases are bit-identical."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import lgar_oracle as O

ALL = golden_names()
STATUS_OF = {"AttributeError": (O.STATUS_NAMES.index("BOTTOM_REACHED"), O.STATUS_NAMES.index("NULL_NEIGHBOUR")),
             "ValueError": (1, 2, 3), "IndexError": (8,), "UnboundLocalError": (5,)}


@pytest.mark.parametrize("name", ALL)
def test_forward_matches_reference(name):
    g = load_golden(name)
    r = O.forward(O.cfg_from_golden(g), g["forcing"])
    T = g["forcing"].shape[0]
    cs = int(g["crash_step"])
    n_ok = T if cs < 0 else cs
    if cs < 0:
        assert r["status"] == 0
    else:
        assert r["crash_step"] == cs
        assert r["status"] in STATUS_OF[str(g["crash_type"])]
    assert abs(float(r["counters"][0]) >= 0)
    np.testing.assert_array_equal(r["nfronts"][:n_ok], g["nfronts"][:n_ok])
    for k in O.OUT_NAMES:
        np.testing.assert_allclose(r[k][:n_ok], g[k][:n_ok], rtol=1e-11, atol=1e-13, err_msg=k)
    if "fronts" in g.files:
        np.testing.assert_array_equal(r["front_layer"][:n_ok], g["front_layer"][:n_ok])
        np.testing.assert_array_equal(r["front_to_bottom"][:n_ok], g["front_to_bottom"][:n_ok])
        np.testing.assert_allclose(r["fronts"][:n_ok], g["fronts"][:n_ok], rtol=1e-11, atol=1e-13)


def test_known_answers_of_the_reference():
    """BASELINE.md section 2 totals (full 8760 h)."""
    want = {"phil_year": dict(start=45.11585035564168, precip=119.88800000000015, infiltration=93.88099503359282,
                              AET=91.87097408980136, runoff=26.007004966407244, end=47.12587129632408),
            "bush_year": dict(start=38.41923497862026, precip=27.330400000000015, infiltration=20.1198807739193,
                              AET=24.670086643148306, runoff=7.210519226080711, end=33.86902910761757)}
    for name, w in want.items():
        g = load_golden(name)
        r = O.forward(O.cfg_from_golden(g), g["forcing"], fronts=False)
        assert r["status"] == 0
        assert float(g["start_volume"]) == pytest.approx(w["start"], rel=1e-14)
        for key in ("precip", "infiltration", "AET", "runoff"):
            assert r[key].sum() == pytest.approx(w[key], rel=1e-10), key
        assert r["ending_volume"][-1] == pytest.approx(w["end"], rel=1e-10)
        assert r["percolation"].sum() == 0.0
        # global mass balance closes (MassBalance.report_mass)
        bal = w["start"] + r["precip"].sum() - r["runoff"].sum() - r["AET"].sum() - r["ponded_water"][-1] \
            - r["percolation"].sum() - r["ending_volume"][-1]
        assert abs(bal) < 1e-8


GRAD = golden_names(prefix="grad_")


@pytest.mark.parametrize("name", GRAD)
def test_tangents_match_reference_autograd(name):
    """Forward-mode tangents of the oracle == reference autograd (SURVEY Q13/Q14 semantics).
    Entries that are pure cancellation noise in the reference graph (|g| < 1e-11 x max|g|) are
    compared absolutely."""
    g = load_golden(name)
    r = O.forward_tangent(O.cfg_from_golden(g), g["forcing"])
    assert r["status"] == 0
    idx = {k: i for i, k in enumerate(O.OUT_NAMES)}
    for loss in ("AET", "infiltration", "runoff", "final_volume"):
        ref = g[f"grad_{loss}"]
        mine = (r["dout"][-1, idx["ending_volume"]] if loss == "final_volume"
                else r["dout"][:, idx[loss]].sum(axis=0)).reshape(3, -1)
        scale = max(np.abs(ref).max(), 1e-300)
        np.testing.assert_allclose(mine, ref, rtol=1e-9, atol=1e-11 * scale, err_msg=loss)


def test_batch_runner_equals_single_runs():
    g = load_golden("rand_phil_1")
    cfgs = [O.cfg_from_golden(load_golden(n)) for n in ("rand_phil_1", "rand_phil_5")]
    sums, st = O.forward_batch(cfgs, g["forcing"], nthreads=2)
    for i, n in enumerate(("rand_phil_1", "rand_phil_5")):
        r = O.forward(cfgs[i], g["forcing"], fronts=False)
        assert st[i] == r["status"]
        assert sums[i, 2] == pytest.approx(r["AET"].sum(), rel=1e-14)
        assert sums[i, 4] == r["ending_volume"][-1]


def test_goldens_enter_the_rare_branches():
    """The golden set must actually exercise the branches a port is most likely to get wrong: the saturated
    free-drainage depth search (A12, check_column_mass), the dry-over-wet fix (A17), the equality fall-through of
    insert_water (Q8), calc_bottom_sum_f_p with the free-drainage front in layer >= 2 (Q18) and the domain-boundary
    pop (A16).  Counted by the oracle (g_cnt[6], [8..11]) over all forward goldens."""
    tot = np.zeros(12, dtype=np.int64)
    for name in ALL:
        g = load_golden(name)
        if "year" in name:
            continue
        tot += O.forward(O.cfg_from_golden(g), g["forcing"], fronts=False)["counters"]
    assert tot[6] > 0, "check_column_mass never iterated"
    assert tot[8] > 0 and tot[9] > 0 and tot[10] > 0 and tot[11] > 0, tot[8:].tolist()
    # the frozen-factor goldens (frozen_factor = 0.7) hit Q18 with the factor applied (Layer.py:1467,1545)
    g = load_golden("frozen_ens_col8074")
    assert float(g["frozen_factor"]) == 0.7
    assert O.forward(O.cfg_from_golden(g), g["forcing"], fronts=False)["counters"][10] > 0


def test_batch_ex_crash_steps_and_tangent_sums():
    names = ("rand_phil_1", "thin_crash")
    g0 = load_golden("rand_phil_1")
    cfgs = [O.cfg_from_golden(g0)]
    r = O.forward_batch_ex(cfgs, g0["forcing"], nthreads=1, tangents=True)
    one = O.forward_tangent(cfgs[0], g0["forcing"])
    assert r["status"][0] == 0 and r["crash_step"][0] == -1
    np.testing.assert_allclose(r["dsums"][0, 0], one["dout"][:, 0].sum(axis=0), rtol=1e-12, atol=1e-300)
    np.testing.assert_array_equal(r["dsums"][0, 4], one["dout"][-1, 4])
    gc = load_golden("thin_crash")
    rc = O.forward_batch_ex([O.cfg_from_golden(gc)], gc["forcing"], nthreads=1)
    assert rc["status"][0] != 0 and rc["crash_step"][0] == int(gc["crash_step"])
    # per-column forcing records through `site`
    f = np.stack([g0["forcing"][:40], g0["forcing"][100:140]])
    two = O.forward_batch_ex([cfgs[0], cfgs[0]], f, site=[0, 1], nthreads=2)
    a = O.forward(cfgs[0], f[1], fronts=False)
    assert two["sums"][1, 2] == pytest.approx(a["AET"].sum(), rel=1e-14)
