"""CPU test of the N>1 path: two ranks over gloo shard a batch by basis index, each rank runs its
shard through the kernels (SIMT-emulated build, CPU), results are gathered and compared with the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from blu_b200.shard import shard_range

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {here!r})
import numpy as np
import torch.distributed as dist
from blu_b200 import BLUBatch, gen, load_library
from blu_b200.shard import shard_range, max_over_ranks, gather_rows
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
NTOT, m = 5, 80
lo, hi = shard_range(NTOT, rank, world)
emu = load_library(os.path.join({here!r}, "emu", "libblu_emu.so"))
bb, be, bi, bx, rhs = gen.batch(hi - lo, m, 24, 4.0, 5000 + lo, 6000 + lo)
b = BLUBatch(hi - lo, m, int((be - bb).reshape(hi - lo, m).sum(1).max()), lib=emu)
st, status = b.factorize(bb, be, bi, bx)
assert st == 0 and (status == 0).all()
st, x, _ = b.solve_dense(rhs, "N")
allx = gather_rows(x, NTOT)
tmax = max_over_ranks(float(rank + 1))
assert tmax == float(world)
if rank == 0:
    np.save(os.environ["OUT"], allx)
dist.destroy_process_group()
'''


def test_shard_range_partitions():
    for n in (0, 1, 5, 4096, 4097):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gloo(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, here=HERE))
    out = tmp_path / "x.npy"
    env = dict(os.environ, OUT=str(out), MASTER_ADDR="127.0.0.1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)], env=env, timeout=600)
    x = np.load(out)
    from blu_b200 import gen
    from oracle_lib import Oracle
    for k in range(5):
        cp, ri, v = gen.basis(5000 + k, 80, 24, 4.0)
        o = Oracle(80, 400 * len(v))
        assert o.factorize(cp[:-1], cp[1:], ri, v) == 0
        _, xo = o.solve_dense(gen.rhs(6000 + k, 80), "N")
        assert np.abs(x[k] - xo).max() <= 1e-12 * np.abs(xo).max()
