"""-m gpu: the dpLGAR-compatible nn.Module (lgar_b200.dpLGAR) used the way the reference's agent uses its
model (agents/DifferentiableLGAR.py:109-171): one forward(x) per forcing row with the caller zeroing the
accumulators (MassBalance.change_mass), then the differentiable whole-record call for training."""
import numpy as np
import pytest
import torch

from conftest import load_golden, max_excess

pytestmark = pytest.mark.gpu


def _cfg(g):
    return dict(
        device="cuda",
        constants=dict(frozen_factor=float(g["frozen_factor"]), nint=int(g["nint"])),
        data=dict(layer_soil_type=[12, 13, 14], layer_thickness=[float(x) for x in g["layer_thickness"]],
                  initial_psi=float(g["initial_psi"]), ponded_depth_max=float(g["ponded_depth_max"]),
                  wilting_point_psi=float(g["wilting_point_psi"]), giuh_ordinates=[float(x) for x in g["giuh_ordinates"]]),
        models=dict(subcycle_length_h=float(g["subcycle_length_h"]), num_subcycles=int(g["num_subcycles"])))


def test_step_by_step_forward_matches_reference():
    from lgar_b200 import dpLGAR
    g = load_golden("phil_4500_400_pdm2")
    model = dpLGAR(_cfg(g), theta_r=g["theta_r"], theta_e=g["theta_e"])
    assert [float(p) for p in model.alpha] == pytest.approx(list(g["alpha"]))
    assert float(model.ending_volume) == pytest.approx(float(g["start_volume"]), rel=1e-12)
    T = 150
    x = torch.tensor(g["forcing"][:T])
    got = {k: [] for k in ("runoff", "percolation", "AET", "infiltration", "ending_volume", "ponded_water")}
    for t in range(T):
        runoff, perc = model(x[t])
        for k in got:
            got[k].append(float(getattr(model, k)))
        # MassBalance.change_mass (physics/MassBalance.py:45-53)
        for k in ("precip", "PET", "AET", "infiltration", "runoff", "percolation", "giuh_runoff", "discharge"):
            setattr(model, k, torch.tensor(0.0, dtype=torch.float64))
    for k in got:
        assert max_excess(np.array(got[k]), g[k][:T]) <= 1.0, k


def test_forward_record_trains():
    """One Adam step on the whole record moves the parameters along the reference-autograd gradient."""
    from lgar_b200 import dpLGAR
    g = load_golden("grad_phil_4550_150")
    model = dpLGAR(_cfg(g), theta_r=g["theta_r"], theta_e=g["theta_e"])
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    out = model.forward_record(torch.tensor(g["forcing"]), outputs=("runoff", "AET"))
    loss = out["AET"].sum()
    loss.backward()
    got = np.array([[float(p.grad) for p in plist] for plist in (model.alpha, model.n, model.ksat)])
    ref = g["grad_AET"]
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-11 * np.abs(ref).max())
    before = float(model.alpha[0])
    opt.step()
    assert float(model.alpha[0]) != before
    model.set_internal_states()
    assert float(model.ending_volume) != pytest.approx(float(g["start_volume"]), rel=1e-9)  # theta_init moved with alpha, n
