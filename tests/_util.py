"""Shared helpers for the test-suite."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_files(prefix):
    return sorted(f for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


def ref_errors(got, ref):
    """The reference's own two metrics (test/blocksparse_matmul_test.py:408-418)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    d = np.abs(got - ref)
    denom = np.abs(ref).mean()
    max_err = d.max() / denom if denom > 0 else d.max()
    nrm = np.sqrt((ref * ref).sum())
    l2_err = np.sqrt((d * d).sum()) / nrm if nrm > 0 else np.sqrt((d * d).sum())
    return float(max_err), float(l2_err)
