"""Single-request latency (BASELINE config 1): one ChaCha20-V3 proof through the reference-compatible outer ABI
(InitAlgorithm once, then Prove(JSON) -> JSON, libprove.go:20-47) and through the inner seam with host buffers
(g16_prove_chacha_batch, n = 1), plus small-batch points. The CPU oracle prover is timed beside it (single proof, all
host threads). Writes gpurun_out/latency.json.
    python scripts/latency.py [iters]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gnark_symmetric_crypto_b200 as G
from conftest import batch_inputs

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
pk = open(os.path.join(ROOT, "tests/golden/pk.chacha20"), "rb").read()
r1 = open(os.path.join(ROOT, "tests/golden/r1cs.chacha20"), "rb").read()
key, nonce, ctr = bytes([2]) * 32, bytes([3]) * 12, 3
pt = bytes.fromhex("a3f7e592aeda1507a7f51b35812dfc50a263d5a6d2df625e563b02e49c08bf30"
                   "d0e7483f5b13ff079532224ee8fbc31ab1899b18e453d36d9793a8355eb0dee9")
out = {}

t = time.perf_counter()
assert G.InitAlgorithm(G.CHACHA20, pk, r1)
out["init_algorithm_s"] = time.perf_counter() - t
req = G.InputParams("chacha20", key, nonce, ctr, pt).to_json()
for _ in range(5):
    G.Prove(req)
lat = []
for _ in range(iters):
    t = time.perf_counter(); G.Prove(req); lat.append((time.perf_counter() - t) * 1e3)
out["prove_json_ms"] = {"median": float(np.median(lat)), "min": float(np.min(lat)), "p90": float(np.percentile(lat, 90))}
print("Prove(JSON) single request:", out["prove_json_ms"], flush=True)

ctx = G.Groth16Context(pk, r1, device=0)
for nb in [b for b in (1, 2, 4, 8, 16, 32, 64, 128, 256) if b <= int(os.environ.get("LAT_MAXB", "256"))]:
    keys, nonces, ctrs, ins, rs = batch_inputs(nb)
    ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)
    ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)
    lat = []
    for _ in range(max(5, iters // 5)):
        t = time.perf_counter(); ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs); lat.append((time.perf_counter() - t) * 1e3)
    st = ctx.stage_ms()
    out[f"batch{nb}"] = {"ms_median": float(np.median(lat)), "ms_min": float(np.min(lat)), "proofs_per_s": nb / (float(np.median(lat)) / 1e3),
                         "stages_ms": st, "launches": ctx.counters()["launches"]}
    print(nb, out[f"batch{nb}"], flush=True)

if "--cpu" in sys.argv:
    from oracle import oracle as O
    orc = O.ChaChaOracleProver(pk, r1)
    r, s = int("11" * 20, 16), int("22" * 20, 16)
    orc.prove(key, nonce, ctr, pt, r, s, nthreads=os.cpu_count())
    lat = []
    for _ in range(3):
        t = time.perf_counter(); orc.prove(key, nonce, ctr, pt, r, s, nthreads=os.cpu_count()); lat.append((time.perf_counter() - t) * 1e3)
    out["cpu_oracle_single_proof_ms"] = {"median": float(np.median(lat)), "threads": os.cpu_count()}
    print("cpu oracle:", out["cpu_oracle_single_proof_ms"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "latency.json"), "w"), indent=1)
