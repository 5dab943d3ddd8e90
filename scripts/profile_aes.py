"""Profiling target: AES-128 context, 2 device-resident runs of a 256-proof batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gnark_symmetric_crypto_b200 as G
from conftest import aes_keys
batch = int(os.environ.get("BATCH", "256"))
pk, vk, r1 = aes_keys(128)
ctx = G.Groth16Context(pk, r1, device=0)
rng = np.random.default_rng(1)
k = rng.integers(0, 256, batch * 16, dtype=np.uint8); no = rng.integers(0, 256, batch * 12, dtype=np.uint8)
c = rng.integers(0, 1 << 31, batch, dtype=np.uint32); i = rng.integers(0, 256, batch * 64, dtype=np.uint8)
ctx.stage_aes(k, 16, no, c, i, None)
for _ in range(int(os.environ.get("RUNS", "2"))):
    ms = ctx.run()
print("ms", ms, ctx.stage_ms())
