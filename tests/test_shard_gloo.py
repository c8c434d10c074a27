"""N>1 path on CPU: world_size-2 gloo run of the sharding + final-gather logic (the per-pair records are
produced by the CPU oracle here, as the checker — the product's compute needs a GPU)."""
import os
import subprocess
import sys

import numpy as np

import mvslam_b200 as mvs
from mvslam_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
import mvslam_b200 as mvs
from mvslam_b200 import shard
from oracle import cbind as orc
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
f = np.load(os.path.join(sys.argv[1], "tests", "golden", "tsukuba_orb2000.npz"))
descs = [f[f"desc{i}"][:400] for i in range(1, 6)]; kps = [f[f"kp{i}"][:400] for i in range(1, 6)]
pairs = np.array([(a, b) for a in range(5) for b in range(5) if a != b][:7], np.int32)
lo, hi = shard.shard_bounds(len(pairs), world, rank)
local = np.zeros(hi - lo, mvs.RESULT_DTYPE)
for i in range(lo, hi):        # stand-in for ctx.pair_batch(pairs[lo:hi], pair_id_base=lo)
    o = orc.image_pair(descs[pairs[i][0]], kps[pairs[i][0]], descs[pairs[i][1]], kps[pairs[i][1]], f["K"],
                       max_dist=30.0, H=16, seed=3, pair_id=i)
    for k in ("status", "n_matches", "n_inliers", "best_hypothesis", "n_points", "candidate", "residual"):
        local[i - lo][k] = o[k]
    for k in ("F", "E", "R1to2", "t1to2", "R2in1", "t2in1"):
        local[i - lo][k] = o[k]
full = shard.gather_records(local, len(pairs), dist)
if rank == 0:
    np.save(sys.argv[2], full)
dist.barrier()
dist.destroy_process_group()
'''


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 130816):
        for w in (1, 2, 3, 8):
            b = [shard.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def test_gather_world2_gloo_equals_single_process(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    outs = {}
    for world in (1, 2):
        out = tmp_path / f"res{world}.npy"
        env = dict(os.environ, OMP_NUM_THREADS="1")
        subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29620 + world), str(script), ROOT, str(out)],
                       check=True, env=env, timeout=300, capture_output=True)
        outs[world] = np.load(out)
    assert outs[1].dtype == mvs.RESULT_DTYPE and len(outs[1]) == 7
    assert outs[1].tobytes() == outs[2].tobytes()       # sharding-invariant, order preserved
    assert (outs[1]["status"] == 0).sum() >= 5
