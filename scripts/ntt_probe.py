"""NTT throughput probe: batch of 2^15 transforms (the compute_h shape) and one 2^24 transform."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnark_symmetric_crypto_b200 as G
for n, batch in ((1 << 15, 512), (1 << 24, 1)):
    ms, chk = G.ntt_bench(n, batch, 10)
    lg = n.bit_length() - 1
    print(f"n=2^{lg} batch={batch}: {ms:.3f} ms  {64.0 * n * batch / ms / 1e6:.1f} GB/s alg  {(n // 2) * lg * batch / ms / 1e6:.2f} Gmul/s  chk={chk}")
