"""Golden vector for the calibration agent: ONE epoch of the UNMODIFIED reference agent
(dpLGAR/agents/DifferentiableLGAR.py: train_one_epoch -> validate: MSE + RangeBoundLoss, loss.backward(), Adam step)
on rows 4550..4749 of the Phillipsburg record (a window with surface runoff, so the loss has a gradient), written to
a temporary csv because the reference's Data class always reads from the top of cfg.data.forcing_file.  Build container only (needs /root/reference).
    python tests/golden/make_agent_golden.py
"""
import os
import sys
import logging

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402

START, NSTEPS, WARMUP = 4550, 200, 24
LB, UB = [0.0015, 1.0, 1e-6, 0.0], [0.015, 5.0, 30, 10.0]


def main():
    torch, DictConfig, _ = rh._import_reference()
    import time
    time.sleep = lambda s: None  # the reference sleeps 10 ms per forcing row
    from dpLGAR.agents.DifferentiableLGAR import DifferentiableLGAR
    logging.basicConfig(level=logging.WARNING)
    import pandas as pd, tempfile
    src = pd.read_csv(f"{rh.REF_ROOT}/data/forcing_data_resampled_uniform_Phillipsburg.csv")
    forcing_file = os.path.join(tempfile.mkdtemp(), "window.csv")
    src.iloc[START:START + NSTEPS].to_csv(forcing_file, index=False)
    cfg = rh.build_cfg(forcing_file=forcing_file, endtime_h=float(NSTEPS))
    cfg.models.hyperparameters = DictConfig(warmup=WARMUP, epochs=1, learning_rate=1e-3, minibatch=0.04166666667, lb=LB, ub=UB)
    agent = DifferentiableLGAR(cfg)
    before = [[float(p) for p in pl] for pl in (agent.model.alpha, agent.model.n, agent.model.ksat)]
    agent.train_one_epoch()
    m = agent.model
    grads = [[0.0 if p.grad is None else float(p.grad) for p in pl] for pl in (m.alpha, m.n, m.ksat)]
    after = [[float(p) for p in pl] for pl in (m.alpha, m.n, m.ksat)]
    y_hat, y_t = agent.y_hat.detach(), agent.y_t.detach()
    loss_mse = float(torch.mean((y_hat - y_t) ** 2))
    bound = float(agent.range_bound_loss([m.alpha, m.n, m.ksat, m.ponded_depth_max]))  # after the step
    from dpLGAR.data.metrics import calculate_nse
    out = dict(start=START, nsteps=NSTEPS, warmup=WARMUP, lb=np.array(LB), ub=np.array(UB), lr=1e-3,
               forcing=agent.data.x.numpy(), y=agent.data.y.numpy(), y_hat=y_hat.numpy(),
               params_before=np.array(before), grads=np.array(grads), params_after=np.array(after),
               loss_mse=loss_mse, bound_after=bound, nse=float(calculate_nse(y_hat.numpy(), y_t.numpy())),
               theta_r=m.c[:, 0].detach().numpy(), theta_e=m.c[:, 1].detach().numpy(),
               layer_thickness=np.array(cfg.data.layer_thickness), initial_psi=cfg.data.initial_psi,
               ponded_depth_max=cfg.data.ponded_depth_max, wilting_point_psi=cfg.data.wilting_point_psi,
               giuh_ordinates=np.array(cfg.data.giuh_ordinates), frozen_factor=float(cfg.constants.frozen_factor),
               nint=int(cfg.constants.nint))
    np.savez_compressed(os.path.join(HERE, "agent_phil_4550_200.npz"), **out)
    print("loss_mse", loss_mse, "nse", out["nse"], "\ngrads\n", out["grads"], "\nafter - before\n", out["params_after"] - out["params_before"])


if __name__ == "__main__":
    main()
