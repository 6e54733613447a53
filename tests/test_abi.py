"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/fsim.h declares,
its struct layouts match the ctypes mirror, and it fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

from fluid_simulator_b200 import abi, sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_declares_entry_points():
    names = sim.exported_symbols()
    assert len(names) >= 30
    for must in ("fsim_create", "fsim_step", "fsim_set_params", "fsim_set_obstacles", "fsim_upload_particles",
                 "fsim_download_grid", "fsim_export_gfx", "fsim_get_step_durations", "fsim_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(sim.LIB_PATH), "libfsim_b200.so has not been built (run __graft_entry__.build())"
    L = C.CDLL(sim.LIB_PATH)
    for name in sim.exported_symbols():
        assert hasattr(L, name), f"{name} declared in include/fsim.h but not exported"
    assert L.fsim_abi_version() == abi.ABI_VERSION


def test_struct_sizes_match_header():
    """sizeof() as the C compiler sees include/fsim.h vs the ctypes mirror."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "fsim.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",sizeof(FsimGridDesc),' \
          'sizeof(FsimGridInfo),sizeof(FsimParams),sizeof(FsimObstacle),sizeof(FsimParticleGfx),sizeof(FsimTimings),sizeof(FsimSolveInfo));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    mirror = [C.sizeof(t) for t in (abi.GridDesc, abi.GridInfo, abi.Params, abi.Obstacle, abi.ParticleGfx, abi.Timings, abi.SolveInfo)]
    assert sizes == mirror


def test_header_cites_reference_for_each_entry_point():
    txt = open(os.path.join(ROOT, "include", "fsim.h")).read()
    assert len(re.findall(r"\w+\.(?:cpp|h|hpp):\d+", txt)) >= 25


def test_no_cpu_fallback():
    """Without a GPU the product path must raise, never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sim.FsimError) as e:
        sim.FluidSim((8.0, 8.0, 8.0))
    assert e.value.code == abi.ERR_CUDA


def test_product_does_not_load_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may touch oracle/."""
    pkg = os.path.join(ROOT, "fluid_simulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f"{f} imports oracle"
                assert "libfsim_oracle" not in txt and "libfsim_ref" not in txt, f"{f} loads an oracle library"
