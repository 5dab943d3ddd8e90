"""Batched groth16.Verify on one GPU: the 1024 proofs of BASELINE config 4 are proved on the GPU and then
verified on the GPU under the reference's vk.chacha20; the CPU oracle verifier is timed on a few of them.
    python scripts/verify_bench.py [batch] -> gpurun_out/verify_bench.json"""
import json, os, struct, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gnark_symmetric_crypto_b200 as G
from conftest import batch_inputs
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pk = open(os.path.join(ROOT, "tests/golden/pk.chacha20"), "rb").read()
r1 = open(os.path.join(ROOT, "tests/golden/r1cs.chacha20"), "rb").read()
vk = open(os.path.join(ROOT, "tests/golden/vk.chacha20"), "rb").read()
ctx = G.Groth16Context(pk, r1, device=0)
keys, nonces, ctrs, ins, rs = batch_inputs(n)
proofs, cts = ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)
pubs = [O.chacha_public_from_signals(cts[i] + nonces[i] + struct.pack("<I", ctrs[i]) + ins[i]) for i in range(n)]
mont = np.stack([O.to_mont(1, O.ints_to_limbs(p)) for p in pubs])
ver = G.Groth16Verifier(vk)
out = {}
for nb in sorted({1, 16, 128, n}):
    ok = ver.verify_batch(proofs[:nb], mont[:nb])
    t = time.perf_counter(); ok = ver.verify_batch(proofs[:nb], mont[:nb]); dt = time.perf_counter() - t
    assert ok.all()
    out[f"batch{nb}"] = {"wall_ms": dt * 1e3, "device_ms": ver.last_ms, "proofs_per_s": nb / dt}
    print(nb, out[f"batch{nb}"], flush=True)
# every 16th proof corrupted: verdicts must single them out
bad = [bytes([p[0]]) + bytes([p[1] ^ 1]) + p[2:] if i % 16 == 3 else p for i, p in enumerate(proofs)]
ok = ver.verify_batch(bad, mont)
assert [bool(x) for x in ok] == [i % 16 != 3 for i in range(n)]
ovk = O.VerifyingKeyOracle(vk)
t = time.perf_counter()
for i in range(4):
    assert ovk.verify(proofs[i], pubs[i])
out["cpu_oracle_ms_per_proof"] = (time.perf_counter() - t) * 1e3 / 4
print(out["cpu_oracle_ms_per_proof"], "ms per proof on the CPU oracle verifier (single thread)")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "verify_bench.json"), "w"), indent=1)
