"""Wall time of g16_prove_chacha_batch for a sequence of varying batch sizes (what the serving worker issues)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gnark_symmetric_crypto_b200 as G
from conftest import batch_inputs
ctx = G.Groth16Context(open("tests/golden/pk.chacha20", "rb").read(), open("tests/golden/r1cs.chacha20", "rb").read())
keys, nonces, ctrs, ins, rs = batch_inputs(1024)
out = []
for nb in [1, 37, 300, 1024, 700, 1024, 512, 130, 1024, 127, 128, 1, 900, 1024, 1024]:
    t = time.perf_counter(); ctx.prove_chacha_batch(keys[:nb], nonces[:nb], ctrs[:nb], ins[:nb], rs[:nb]); dt = (time.perf_counter() - t) * 1e3
    out.append((nb, round(dt, 1), round(ctx.stage_ms()["total"], 1)))
print(os.environ.get("G16_EVAL_Z", "auto"), out)
