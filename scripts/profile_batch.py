"""Profiling / tuning target: init + RUNS device-resident runs of a BATCH-proof ChaCha batch (no oracle, no CPU work).
G16_LIB=<path> loads an experimental build of the CUDA library (scripts/ba_variants.sh) instead of lib/libg16b200.so."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from gnark_symmetric_crypto_b200 import _lib
if os.environ.get("G16_LIB"):
    from pathlib import Path
    _lib.LIB_PATH = Path(os.environ["G16_LIB"]).resolve()
import gnark_symmetric_crypto_b200 as G
n = int(os.environ.get("BATCH", "1024"))
runs = int(os.environ.get("RUNS", "2"))
ctx = G.Groth16Context(open("tests/golden/pk.chacha20", "rb").read(), open("tests/golden/r1cs.chacha20", "rb").read())
k, no, c, i, r = ctx._pack(*bench.make_requests(n, b"g16-b200-batch"))[1:]
ctx.stage(k, no, c, i, r)
best = None
for _ in range(runs):
    ms = ctx.run()
    st = ctx.stage_ms()
    if best is None or st["total"] < best["total"]:
        best = st
proofs = np.zeros(n * ctx.proof_bytes, dtype=np.uint8); cts = np.zeros(n * 64, dtype=np.uint8)
ctx.fetch(proofs, cts)
import hashlib
print(json.dumps({"tag": os.environ.get("TAG", ""), "batch": n, "ms": round(best["total"], 2), "proofs_per_s": round(n / best["total"] * 1e3, 1),
                  "stages": {a: round(b, 2) for a, b in best.items()}, "sha": hashlib.sha256(proofs.tobytes()).hexdigest()[:16]}))
