"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled into
oracle/_ref/libfsim_ref_serial.so (serial -D_DEBUG build => bit-reproducible).  Run in the build
container (needs /root/reference to have been compiled by oracle/build_ref.sh):

    python tests/golden/make_golden.py

The reference has no golden vectors of its own; these fixtures are outputs of the
reference itself and are what pins oracle/fsim_oracle.c on machines without /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle.refsim import RefSim  # noqa: E402
import cases  # noqa: E402


def main():
    for name in cases.CASES:
        out = cases.run_case(RefSim, name, serial=True)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, {k: v.shape for k, v in out.items()}, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
