"""CPU: the N>1 plumbing of bench.py (shard assignment, barrier, max-over-ranks timing, whole-job aggregation) under
torch.distributed with the gloo backend and world_size 2. The data path has no collective (independent proofs)."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_shard_ranges_cover_everything():
    sys.path.insert(0, str(ROOT))
    import bench
    for total in (1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = bench.shard_range(total, r, world)
                got.extend(range(lo, hi))
            assert got == list(range(total))
    # MSM point-range split (config 5): host adds the partial points
    assert bench.shard_range(1 << 20, 3, 8) == (3 << 17, 4 << 17)


def test_strong_scaling_shards_and_gather_of_partial_points(tmp_path):
    """The fixed-total reading of config 4 (request i -> rank i mod N) covers every request exactly once, and the gather that
    bench.py uses for the partial points of the split MSM (all_gather of 8 x u64 per rank) returns them in rank order."""
    sys.path.insert(0, str(ROOT))
    for world in (1, 2, 4, 8, 3):
        idx = sorted(i for r in range(world) for i in list(range(1024))[r::world])
        assert idx == list(range(1024))
    script = tmp_path / "g.py"
    script.write_text("""
import json, numpy as np, torch, torch.distributed as dist
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
part = (np.arange(8, dtype=np.uint64) + np.uint64(0xFFFFFFFF00000000) + np.uint64(100 * r))
t = torch.from_numpy(part.view(np.int64).copy())
out = [torch.empty_like(t) for _ in range(w)]
dist.all_gather(out, t)
parts = [g.numpy().view(np.uint64) for g in out]
if r == 0:
    print(json.dumps([[int(x) for x in p] for p in parts]))
dist.destroy_process_group()
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29534")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29534", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    parts = json.loads([l for l in out.stdout.splitlines() if l.startswith("[")][-1])
    assert parts == [[0xFFFFFFFF00000000 + k + 100 * r for k in range(8)] for r in range(2)]


def test_gloo_world2_aggregation(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(f"""
import sys, json
sys.path.insert(0, {str(ROOT)!r})
import torch.distributed as dist
import bench
dist.init_process_group("gloo")
r = dist.get_rank()
ms, units = bench.aggregate(local_ms=10.0 * (r + 1), local_units=100 + r, device=None)
if r == 0:
    print(json.dumps({{"ms": ms, "units": units}}))
dist.destroy_process_group()
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ms"] == 20.0 and res["units"] == 201   # max over ranks, sum of units
