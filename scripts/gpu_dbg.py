import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
import gnark_symmetric_crypto_b200 as G
rng = np.random.default_rng(17)
n = 8
P = O.g1_fixed_base(O.rand_field(rng, 1, n))
for kval in (1, 2, 3, 5, 2**31, 2**32, 2**33+1, 2**64+7, 2**200+3, O.R_MOD-1):
    sc = O.ints_to_limbs([kval]*n)
    got = G.group_op(1, "mul", P, sc)
    ref = np.array([O.g1_mul(P[i], kval) for i in range(n)])
    print(kval.bit_length(), np.array_equal(got, ref), [bool(np.array_equal(got[i], ref[i])) for i in range(n)])
got = G.group_op(1, "dbl", P); ref = np.array([O.g1_add(P[i], P[i]) for i in range(n)]); print('dbl', np.array_equal(got, ref))
P2 = O.g2_fixed_base(O.rand_field(rng, 1, 4))
for kval in (1, 2, 3, 2**64+7, O.R_MOD-1):
    sc = O.ints_to_limbs([kval]*4)
    got = G.group_op(2, "mul", P2, sc); ref = np.array([O.g2_mul(P2[i], kval) for i in range(4)])
    print('g2', kval.bit_length(), np.array_equal(got, ref))
# identify what the wrong lanes computed
for kval in (2, 3):
    sc = O.ints_to_limbs([kval]*n)
    got = G.group_op(1, "mul", P, sc)
    for i in range(n):
        match = [m for m in range(0, 9) if np.array_equal(got[i], O.g1_mul(P[i], m) if m else np.zeros(8, dtype=np.uint64))]
        print('k', kval, 'lane', i, 'equals multiples', match, 'on curve', bool(O.lib().orc_g1_on_curve(np.ascontiguousarray(got[i]).ctypes.data_as(O.u64p), 1)))
