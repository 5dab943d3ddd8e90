"""Profiling target: init + 2 device-resident runs of a 1024-proof ChaCha batch (no oracle, no CPU work)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import gnark_symmetric_crypto_b200 as G
from conftest import batch_inputs
n = int(os.environ.get("BATCH", "1024"))
runs = int(os.environ.get("RUNS", "2"))
ctx = G.Groth16Context(open("tests/golden/pk.chacha20", "rb").read(), open("tests/golden/r1cs.chacha20", "rb").read())
k, no, c, i, r = ctx._pack(*batch_inputs(n))[1:]
ctx.stage(k, no, c, i, r)
for _ in range(runs):
    ms = ctx.run()
print("ms", ms, ctx.stage_ms(), ctx.counters())
