"""Forcing loader -> device tiles: the step in front of the hot path.

Reads the reference's forcing formats into the `[T, 2]` (P, PET) cm/h layout the kernels consume
(dpLGAR/data/Data.py:26-40 multiplies mm/h by cfg.conversions.mm_to_cm = 0.1):
  * `.csv` with a `Time,P(mm/h),PET(mm/h)` header (data/forcing_data_resampled_uniform_*.csv);
  * `.txt` with a `#Time,...` header (data/forcing_data_synth_*.txt) -- the reference's read_df rejects these;
  * whitespace-separated files without commas (data/forcing_data_syn_case3.txt).
`stack_sites` builds the pinned `[sites, T, 2]` host array for several records of equal length, ready for one
asynchronous H2D copy into `ColumnEnsemble.forcing`.
"""
from __future__ import annotations

import numpy as np
import torch

MM_TO_CM = 0.1


def read_forcing(path: str, nrows: int | None = None) -> np.ndarray:
    with open(path) as f:
        header = f.readline().strip().lstrip("#")
    sep = "," if "," in header else None
    names = [h.strip() for h in (header.split(sep) if sep else header.split())]
    lower = [n.lower() for n in names]
    try:
        ip = next(i for i, n in enumerate(lower) if n.startswith("p(") or n in ("p", "precip", "precipitation"))
        ie = next(i for i, n in enumerate(lower) if n.startswith("pet"))
    except StopIteration as e:
        raise ValueError(f"{path}: no P / PET columns in header {names}") from e
    if sep == ",":
        # comma files go through pandas exactly like the reference's read_df (data/utils.py:19-37), so the
        # parsed doubles are bit-identical to what Data.py feeds the model (pandas' default float parser is
        # not always correctly rounded)
        import pandas as pd
        df = pd.read_csv(path, nrows=nrows)
        x = np.stack([df.iloc[:, ip].values, df.iloc[:, ie].values], axis=1).astype(np.float64)
        return x * MM_TO_CM
    rows = []
    with open(path) as f:
        f.readline()
        for line in f:
            parts = line.split()
            if not parts:
                continue
            # a timestamp "YYYY-MM-DD HH:MM:SS" splits into two tokens in whitespace-separated files
            shift = len(parts) - len(names)
            rows.append((float(parts[ip + (shift if ip > 0 else 0)]), float(parts[ie + (shift if ie > 0 else 0)])))
            if nrows is not None and len(rows) >= nrows:
                break
    return np.asarray(rows, dtype=np.float64) * MM_TO_CM


def stack_sites(records, pin: bool = True) -> torch.Tensor:
    T = min(r.shape[0] for r in records)
    x = torch.from_numpy(np.stack([np.asarray(r[:T], dtype=np.float64) for r in records]))
    return x.pin_memory() if (pin and torch.cuda.is_available()) else x
