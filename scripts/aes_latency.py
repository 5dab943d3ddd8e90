"""Single-request latency of the AES-128/256 circuits (keys from the oracle's Setup restatement)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gnark_symmetric_crypto_b200 as G
from conftest import aes_keys
for bits in (128, 256):
    pk, vk, r1 = aes_keys(bits)
    ctx = G.Groth16Context(pk, r1, device=0)
    rng = np.random.default_rng(bits)
    args = ([rng.bytes(bits // 8)], [rng.bytes(12)], [5], [rng.bytes(64)], None)
    for _ in range(3):
        ctx.prove_aes_batch(*args)
    lat = []
    for _ in range(10):
        t = time.perf_counter(); ctx.prove_aes_batch(*args); lat.append((time.perf_counter() - t) * 1e3)
    print(f"aes{bits}: one proof {np.median(lat):.2f} ms", ctx.stage_ms(), flush=True)
    ctx.close()
