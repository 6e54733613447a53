#!/usr/bin/env python
"""One-step parity of the CUDA path against the compiled reference on the HEADLINE scene bench.py times: 3D FLIP dam break
256^3, 65,548,256 particles (and optionally cfg 3: APIC + the static box).  The reference needs ~11 GB of host memory and
minutes per step (MIC(0)-PCG: ~155 iterations at 256^3), so this is a tool, not a test: run it once per round on the GPU
box and commit its output (profiles/r2_parity_256.json).

    python tools/parity_256.py [--grid 256] [--tol 1e-6 1e-9] [--apic-box] [--out gpurun_out/parity_256.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--tol", type=float, nargs="+", default=[1e-6])
    ap.add_argument("--apic-box", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_256.json"))
    args = ap.parse_args()

    import scale_parity
    from fluid_simulator_b200 import abi, scenes
    from fluid_simulator_b200.sim import FluidSim

    n = args.grid
    out = []
    sc = scenes.dam_break_3d(n, abi.FLIP)
    for tol in args.tol:
        r = scale_parity.one_step(FluidSim, sc, tol=tol)
        r["scene"] = f"3D FLIP dam break {n}^3 (bench.py workload)"
        print(json.dumps(r, default=float), flush=True)
        out.append(r)
    if args.apic_box:
        sc.params.transfer_type = abi.APIC
        r = scale_parity.one_step(FluidSim, sc, obstacles=[scenes.cfg3_box(n)], apic=True)
        r["scene"] = f"3D APIC dam break {n}^3 + static box (SURVEY cfg 3)"
        print(json.dumps(r, default=float), flush=True)
        out.append(r)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1, default=float)


if __name__ == "__main__":
    main()
