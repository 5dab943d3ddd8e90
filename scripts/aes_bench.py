"""AES-128 / AES-256 (V2 lookup-table circuits, BSB22 commitment) batched proving throughput on one GPU.
Keys come from the oracle's Setup restatement (tests/conftest.py::aes_keys) because the reference ships no pk.aes*.
    python scripts/aes_bench.py [batch] -> gpurun_out/aes_bench.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gnark_symmetric_crypto_b200 as G
from conftest import aes_keys

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
out = {}
for bits in (128, 256):
    pk, vk, r1 = aes_keys(bits)
    ctx = G.Groth16Context(pk, r1, device=0)
    rng = np.random.default_rng(bits)
    klen = bits // 8
    k = rng.integers(0, 256, batch * klen, dtype=np.uint8); no = rng.integers(0, 256, batch * 12, dtype=np.uint8)
    c = rng.integers(0, 1 << 31, batch, dtype=np.uint32); i = rng.integers(0, 256, batch * 64, dtype=np.uint8)
    ctx.stage_aes(k, klen, no, c, i, None)
    ctx.run(); ctx.run()
    ms = [ctx.run() for _ in range(3)]
    st = ctx.stage_ms(); cn = ctx.counters()
    proofs = np.zeros(batch * ctx.proof_bytes, dtype=np.uint8); cts = np.zeros(batch * 64, dtype=np.uint8)
    ctx.fetch(proofs, cts)
    assert len({proofs[j * 196:(j + 1) * 196].tobytes() for j in range(batch)}) == batch
    out[f"aes{bits}"] = {"batch": batch, "ms_per_batch": float(np.mean(ms)), "proofs_per_s": batch / (float(np.mean(ms)) / 1e3),
                         "stages_ms": st, "counters": cn, "domain": ctx.n, "wires": ctx.nb_wires, "constraints": ctx.nb_constraints}
    print(bits, out[f"aes{bits}"], flush=True)
    ctx.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "aes_bench.json"), "w"), indent=1)
