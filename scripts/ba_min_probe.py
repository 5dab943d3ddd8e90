"""Small ChaCha batches under different thresholds of the batch-affine path (G16_MSM_BA_MIN = entries of one MSM from which the
pairwise affine levels run in front of the XYZZ accumulation): median wall time of g16_prove_chacha_batch and the accumulate stage.
    G16_MSM_BA_MIN=16777216 python scripts/ba_min_probe.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import gnark_symmetric_crypto_b200 as G
ctx = G.Groth16Context(open(os.path.join(ROOT, "tests/golden/pk.chacha20"), "rb").read(), open(os.path.join(ROOT, "tests/golden/r1cs.chacha20"), "rb").read())
keys, nonces, ctrs, ins, rs = bench.make_requests(256, b"g16-b200-batch")
res = {}
for nb in [int(v) for v in os.environ.get("BATCHES", "2,4,8,16,32,64,128,256").split(",")]:
    a = (keys[:nb], nonces[:nb], ctrs[:nb], ins[:nb], rs[:nb])
    for _ in range(3):
        ctx.prove_chacha_batch(*a)
    lat = []
    for _ in range(7):
        t = time.perf_counter(); ctx.prove_chacha_batch(*a); lat.append((time.perf_counter() - t) * 1e3)
    st = ctx.stage_ms()
    res[nb] = (round(float(np.median(lat)), 2), round(st["msm_accumulate"], 2), round(st["msm_reduce"], 2))
print(json.dumps({"G16_MSM_BA_MIN": os.environ.get("G16_MSM_BA_MIN", "default"), "batch: (ms, accumulate, reduce)": res}), flush=True)
