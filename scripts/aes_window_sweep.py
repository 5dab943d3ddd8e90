"""AES-128 batch throughput for one setting of the window environment variables (G16_C_A / _B / _K / _B2 / _PED / _Z), which
the context reads when it is created.   G16_C_B2=12 python scripts/aes_window_sweep.py [batch]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gnark_symmetric_crypto_b200 as G
from conftest import aes_keys

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pk, vk, r1 = aes_keys(128)
ctx = G.Groth16Context(pk, r1, device=0)
rng = np.random.default_rng(128)
k = rng.integers(0, 256, batch * 16, dtype=np.uint8); no = rng.integers(0, 256, batch * 12, dtype=np.uint8)
c = rng.integers(0, 1 << 31, batch, dtype=np.uint32); i = rng.integers(0, 256, batch * 64, dtype=np.uint8)
ctx.stage_aes(k, 16, no, c, i, None)
ctx.run(); ctx.run()
ms = [ctx.run() for _ in range(3)]
st = ctx.stage_ms()
env = {k: v for k, v in os.environ.items() if k.startswith("G16_C_")}
print(json.dumps({"env": env, "batch": batch, "ms_per_batch": round(float(np.mean(ms)), 2), "proofs_per_s": round(batch / float(np.mean(ms)) * 1e3, 1),
                  "stages_ms": {a: round(b, 1) for a, b in st.items()}}), flush=True)
