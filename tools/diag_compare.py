"""Compare two gpu_diag.py result dumps (LGAR_DIAG_SAVE=1): status / crash step must be identical, sums bit-identical."""
import sys, numpy as np
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
same_st = np.array_equal(a["status"], b["status"]) and np.array_equal(a["crash"], b["crash"])
sa, sb = a["sums"], b["sums"]
bit = np.array_equal(sa.view(np.int64), sb.view(np.int64))
with np.errstate(invalid="ignore", divide="ignore"):
    rel = np.nanmax(np.abs(sa - sb) / np.maximum(np.abs(sa), 1e-300))
print(f"status/crash identical: {same_st}; sums bit-identical: {bit}; max rel diff {rel:.3g}; differing entries {(sa.view(np.int64) != sb.view(np.int64)).sum()} of {sa.size}")
