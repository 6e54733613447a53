import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIAG = os.path.join(ROOT, "gpurun_out", "parity_diag.jsonl")


def rel_l2(a, b):
    """||a - b||_2 / ||b||_2 (b = oracle); 0 when both are zero."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    nb = np.linalg.norm(b)
    if nb == 0:
        return float(np.linalg.norm(a))
    return float(np.linalg.norm(a - b) / nb)


def diag(**kw):
    """Appends one line of diagnostics to gpurun_out/parity_diag.jsonl (merged back from the GPU box)."""
    try:
        os.makedirs(os.path.dirname(DIAG), exist_ok=True)
        with open(DIAG, "a") as f:
            f.write(json.dumps(kw, default=float) + "\n")
    except OSError:
        pass
    print("DIAG", json.dumps(kw, default=float))
