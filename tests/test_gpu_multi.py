"""Two-rank run of the library-owned exchange (csrc/comm.cu: zk_comm_init_rank + zk_msm_sharded) on two GPUs of one node —
skipped on single-GPU boxes.  The N-way split is the reference's own benchmark shape, poly-commitment/benches/msm.rs:92-140."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ZK_ROOT"])
import proof_systems_b200 as zk
from proof_systems_b200.parallel import LibraryComm, shard_bounds
from oracle import oracle as orc
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ctx = zk.Context(rank)
comm = LibraryComm(ctx)
n = 1 << 13
pts = ctx.synthetic_points(zk.VESTA, n, seed=5)
sc = orc.random_scalars(orc.FP, n, seed=6)
lo, hi = shard_bounds(n, world, rank)
bases = ctx.upload_bases(zk.VESTA, pts[lo:hi], window_bits=12)
mine = np.ascontiguousarray(sc[lo:hi])
got = zk.jacobian_to_affine(zk.VESTA, comm.msm(bases, mine.ctypes.data, hi - lo))
want = orc.msm(orc.VESTA, pts, sc)
assert np.array_equal(got, want), rank
comm.close()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_sharded_msm_through_the_library_communicator(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, ZK_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2
