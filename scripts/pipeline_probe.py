"""Two-lane pipelined schedule (g16_set_schedule pipeline = 1: sub-batch k runs its whole chain on lane k % 2) against the default
single main stream, on the current kernels.  BATCH, RUNS as in profile_batch.py."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import gnark_symmetric_crypto_b200 as G
n = int(os.environ.get("BATCH", "1024")); runs = int(os.environ.get("RUNS", "4"))
ctx = G.Groth16Context(open("tests/golden/pk.chacha20", "rb").read(), open("tests/golden/r1cs.chacha20", "rb").read())
k, no, c, i, r = ctx._pack(*bench.make_requests(n, b"g16-b200-batch"))[1:]
ctx.stage(k, no, c, i, r)
for pipe, sb in ((0, 512), (1, 512), (1, 384), (1, 342), (1, 256), (0, 512)):
    ctx.set_schedule(pipe, sb)
    best = min(ctx.run() for _ in range(runs))
    print(json.dumps({"pipeline": pipe, "sub_batch": sb, "ms": round(best, 2), "proofs_per_s": round(n / best * 1e3, 1)}), flush=True)
