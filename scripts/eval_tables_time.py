"""Times the one-off construction of the evaluation-basis tables of the Z query (first large batch of a context)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gnark_symmetric_crypto_b200 as G
from conftest import batch_inputs
t0 = time.perf_counter()
ctx = G.Groth16Context(open("tests/golden/pk.chacha20", "rb").read(), open("tests/golden/r1cs.chacha20", "rb").read())
t1 = time.perf_counter()
k, no, c, i, r = ctx._pack(*batch_inputs(256))[1:]
ctx.stage(k, no, c, i, r)
os.environ["X"] = "1"
ctx.set_schedule(False, 512)
t2 = time.perf_counter(); ctx.run(); t3 = time.perf_counter(); ctx.run(); t4 = time.perf_counter()
print({"ctx_init_s": t1 - t0, "first_run_256_s": t3 - t2, "second_run_256_s": t4 - t3, "counters": ctx.counters()})
