"""Two lanes for batches below one sub-batch: n requests cut into sub-batches of n / 2 (or n / 4) against the single stream."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import gnark_symmetric_crypto_b200 as G
ctx = G.Groth16Context(open("tests/golden/pk.chacha20", "rb").read(), open("tests/golden/r1cs.chacha20", "rb").read())
# let the context learn the wire classes first (256 witnesses) and build its tables
k, no, c, i, r = ctx._pack(*bench.make_requests(512, b"g16-b200-batch"))[1:]
ctx.set_schedule(0, 512); ctx.stage(k, no, c, i, r); ctx.run(); ctx.run()
for n in (64, 128, 256, 512, 768):
    k, no, c, i, r = ctx._pack(*bench.make_requests(n, b"g16-b200-batch"))[1:]
    ctx.stage(k, no, c, i, r)
    for pipe, sb in ((0, 512), (1, n // 2), (1, n // 4), (1, (n + 2) // 3)):
        ctx.set_schedule(pipe, sb)
        for _ in range(2): ctx.run()
        best = min(ctx.run() for _ in range(5))
        print(json.dumps({"n": n, "pipeline": pipe, "sub_batch": sb, "ms": round(best, 2), "proofs_per_s": round(n / best * 1e3, 1)}), flush=True)
