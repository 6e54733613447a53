"""Host-side logic of the z-slab decomposition, without a GPU: the partition arithmetic (library vs numpy mirror), particle
ownership, stitching of local blocks, struct layouts, and the export exchange over a 2-rank gloo group (the transport the
production path uses: fluid_simulator_b200.slab.connect_torch)."""
import ctypes as C
import os
import socket
import subprocess
import tempfile

import numpy as np
import pytest

from fluid_simulator_b200 import abi, sim, slab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("gz,n", [(256, 8), (256, 2), (26, 3), (16, 4), (17, 5), (128, 1)])
def test_partition_matches_library(gz, n):
    L = sim.load_library()
    parts = slab.partition(gz, n)
    covered = np.zeros(gz, dtype=np.int32)
    for r, (lo, hi, zoff, gzl) in enumerate(parts):
        info = abi.SlabInfo()
        assert L.fsim_slab_partition(gz, r, n, C.byref(info)) == 0
        assert (info.own_lo, info.own_hi, info.z_offset, info.gz_local) == (lo, hi, zoff, gzl)
        covered[lo:hi] += 1
        assert zoff <= lo and zoff + gzl >= hi                       # owned planes are stored
        assert lo - zoff == (0 if r == 0 else 1)                      # exactly one ghost plane towards a neighbour
        assert zoff + gzl - hi == (0 if r == n - 1 else 1)
    assert np.all(covered == 1)                                       # every plane has exactly one owner
    assert L.fsim_slab_partition(gz, n, n, C.byref(abi.SlabInfo())) != 0


def test_owner_uses_the_device_binning_expression():
    # fp32-rounded coordinate first, then the fp64 product, then truncation (simulator.cpp:358-359 on the stored value)
    z = np.array([0.0, 0.999999, 1.0, 7.9999999, 8.0, 15.99, 200.0])
    own = slab.owner_of(z, 1.0, 16, 2)
    assert own.tolist() == [0, 0, 0, 1, 1, 1, 1]      # 7.9999999 rounds to 8.0f: it is binned in plane 8
    res = 1.508
    z = np.linspace(0.01, 10.5, 1000)
    iz = np.trunc(z.astype(np.float32).astype(np.float64) * res).astype(int)
    parts = slab.partition(16, 4)
    own = slab.owner_of(z, res, 16, 4)
    for k in range(z.size):
        lo, hi = parts[own[k]][0], parts[own[k]][1]
        assert lo <= min(iz[k], 15) < hi


def test_stitch_takes_each_plane_from_its_owner():
    gx, gy, gz, n = 3, 2, 11, 3
    full = np.arange(gx * gy * gz * 3, dtype=np.float64).reshape(gx, gy, gz, 3)
    parts = slab.partition(gz, n)
    blocks = []
    for lo, hi, zoff, gzl in parts:
        b = full[:, :, zoff:zoff + gzl].copy()
        if lo > zoff:
            b[:, :, 0] = -1       # ghost planes hold garbage: they must never be taken
        if zoff + gzl > hi:
            b[:, :, -1] = -1
        blocks.append(b.reshape(-1, 3))
    assert np.array_equal(slab.stitch(blocks, parts, (gx, gy, gz), (3,)), full)


def test_slab_struct_sizes_match_header():
    src = '#include <stdio.h>\n#include "fsim.h"\nint main(){printf("%zu %zu %zu %d\\n",sizeof(FsimSlabInfo),sizeof(FsimDistExport),sizeof(FsimDistWaitStats),FSIM_WAIT_CLASSES);return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert sizes == [C.sizeof(abi.SlabInfo), C.sizeof(abi.DistExport), C.sizeof(abi.DistWaitStats), len(abi.WAIT_CLASSES)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeSim:  # the connect plumbing only needs dist_export / dist_connect
        def dist_export(self):
            ex = abi.DistExport()
            ex.rank, ex.nranks, ex.device, ex.pid = rank, world, rank, os.getpid()
            ex.raw[3] = 0x1000 * (rank + 1)
            ex.ipc[5][7] = 40 + rank
            return ex

        def dist_connect(self, exports):
            self.seen = [(e.rank, e.nranks, int(e.raw[3]), int(e.ipc[5][7]), int(e.pid)) for e in exports]
    s = FakeSim()
    slab.connect_torch(s, dist)
    out.put((rank, s.seen))
    dist.destroy_process_group()


def test_export_exchange_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        seen = got[r]
        assert [(s[0], s[1], s[2], s[3]) for s in seen] == [(0, 2, 0x1000, 40), (1, 2, 0x2000, 41)]
        assert seen[0][4] != seen[1][4]      # two processes: the library will take the IPC path


def test_cpp_slab_host_builds_against_c_abi():
    """The C++ host of the z-slab ABI (tests/cpp/slab_host.cpp: std::thread per rank, plain fsim.h) compiles and links."""
    subprocess.check_call(["bash", os.path.join(ROOT, "tests", "cpp", "build_slab_host.sh")])
    assert os.path.exists(os.path.join(ROOT, "tests", "cpp", "build", "slab_host"))
