"""lgar_b200 -- B200-native time-stepping core of dpLGAR (hand-written sm_100a CUDA behind a C ABI).

Public surface:
  ColumnEnsemble, lgar_columns, forward_raw   batched columns (columns.py)
  dpLGAR                                      drop-in nn.Module mirroring the reference model API
  agent.DifferentiableLGAR                    calibration agent mirroring the reference training loop (multi-site, distributed)
  _capi                                       ctypes binding of include/lgar_b200.h
"""
from . import _capi
from ._capi import OUT_NAMES, STATUS_NAMES, LGARLibraryError
from .columns import ColumnEnsemble, ForwardResult, forward_raw, lgar_columns, output_mask
from .model import dpLGAR
from . import agent, forcing, parallel, workloads

__all__ = ["ColumnEnsemble", "ForwardResult", "forward_raw", "lgar_columns", "output_mask", "OUT_NAMES",
           "STATUS_NAMES", "LGARLibraryError", "dpLGAR", "agent", "forcing", "parallel", "workloads", "_capi"]
