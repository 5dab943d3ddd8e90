"""Timing of the pairing product check (4 pairs per check, the Groth16 shape) for several batch sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import gnark_symmetric_crypto_b200 as G
from oracle import oracle as O
rng = np.random.default_rng(1)
R = O.R_MOD
a = [int(x) for x in rng.integers(1, 1 << 62, 3)]; b = [int(x) for x in rng.integers(1, 1 << 62, 3)]
g1 = O.g1_fixed_base(O.ints_to_limbs(a + [(-sum(x * y for x, y in zip(a, b))) % R]))
g2 = O.g2_fixed_base(O.ints_to_limbs(b + [1]))
for n in (1, 64, 256, 1024):
    P = np.tile(g1, (n, 1)); Q = np.tile(g2, (n, 1))
    G.pairing_check(P, Q, 4)
    t = time.perf_counter(); ok = G.pairing_check(P, Q, 4); dt = time.perf_counter() - t
    assert ok.all()
    print(f"checks {n}: {dt*1e3:.1f} ms  ({n/dt:.0f} checks/s)", flush=True)
