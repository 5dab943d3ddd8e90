"""Throughput of the reference-shaped API (one request per Prove(JSON) call) under concurrent callers: the library coalesces
concurrent calls into GPU batches (libprove_abi.cpp). T caller threads each issue R calls back to back.
    python scripts/serve_bench.py [threads] [calls_per_thread] -> gpurun_out/serve_bench.json"""
import json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gnark_symmetric_crypto_b200 as G

pk = open(os.path.join(ROOT, "tests/golden/pk.chacha20"), "rb").read()
r1 = open(os.path.join(ROOT, "tests/golden/r1cs.chacha20"), "rb").read()
assert G.InitAlgorithm(G.CHACHA20, pk, r1)
rng = np.random.default_rng(7)
out = {}
for T in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1,16,64,256,1024".split(","))]:
    R = int(sys.argv[2]) if len(sys.argv) > 2 else max(2, 2048 // T)
    reqs = [[G.InputParams("chacha20", rng.bytes(32), rng.bytes(12), int(rng.integers(0, 1 << 32)), rng.bytes(64)).to_json()
             for _ in range(R)] for _ in range(T)]
    lat = [[] for _ in range(T)]

    def worker(t):
        for q in reqs[t]:
            t0 = time.perf_counter(); G.Prove(q); lat[t].append(time.perf_counter() - t0)

    for rep in range(2):   # the first round warms up (device buffers grow to the batch sizes that occur); the second is reported
        for l in lat:
            l.clear()
        th = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
        t0 = time.perf_counter()
        [x.start() for x in th]; [x.join() for x in th]
        dt = time.perf_counter() - t0
    allat = np.array([x for l in lat for x in l[-R:]]) * 1e3
    out[f"threads{T}"] = {"calls": T * R, "wall_s": dt, "proofs_per_s": T * R / dt, "latency_ms_median": float(np.median(allat)),
                          "latency_ms_p95": float(np.percentile(allat, 95))}
    print(T, out[f"threads{T}"], flush=True)
# the same for Verify(JSON): 1024 proofs from one batch, T caller threads
import struct
from conftest import batch_inputs
ctx = G.Groth16Context(pk, r1, device=0)
n = 1024
keys, nonces, ctrs, ins, rs = batch_inputs(n)
proofs, cts = ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)
assert G.InitVerifier(G.CHACHA20, open(os.path.join(ROOT, "tests/golden/vk.chacha20"), "rb").read())
vreqs = [G.InputVerifyParams("chacha20", proofs[i], cts[i] + nonces[i] + struct.pack("<I", ctrs[i]) + ins[i]).to_json() for i in range(n)]
for T in (1, 64, 1024):
    R = n // T
    okc = [0] * T

    def vworker(t):
        for q in vreqs[t * R:(t + 1) * R][:(8 if T == 1 else R)]:
            okc[t] += int(G.Verify(q))

    for rep in range(2):
        okc = [0] * T
        th = [threading.Thread(target=vworker, args=(t,)) for t in range(T)]
        t0 = time.perf_counter()
        [x.start() for x in th]; [x.join() for x in th]
        dt = time.perf_counter() - t0
    calls = 8 if T == 1 else n
    assert sum(okc) == calls
    out[f"verify_threads{T}"] = {"calls": calls, "wall_s": dt, "verifications_per_s": calls / dt}
    print("verify", T, out[f"verify_threads{T}"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "serve_bench.json"), "w"), indent=1)
