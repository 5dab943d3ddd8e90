"""compute-sanitizer target: one small pass over every kernel family of the library — a ChaCha batch
(both Z-query bases), an AES-128 batch (commitment path), a standalone MSM with and without tables, NTT round trips, the
batch verifier — without the oracle or any CPU work. Run as
    compute-sanitizer --tool memcheck  python scripts/sanitize_run.py
    compute-sanitizer --tool racecheck python scripts/sanitize_run.py
WHAT=chacha|aes|msm|ntt|verify (comma separated) selects parts; sizes are small because the tools run kernels 10-100x slower."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import gnark_symmetric_crypto_b200 as G

what = os.environ.get("WHAT", "chacha,aes,msm,ntt,verify").split(",")
os.environ.setdefault("G16_LAZY_TABLES", "0")
rng = np.random.default_rng(1)
if "chacha" in what or "verify" in what:
    pk = open("tests/golden/pk.chacha20", "rb").read(); r1 = open("tests/golden/r1cs.chacha20", "rb").read()
    ctx = G.Groth16Context(pk, r1)
    n = int(os.environ.get("BATCH", "3"))
    reqs = bench.make_requests(n, b"g16-b200-batch")
    proofs, cts = ctx.prove_chacha_batch(*reqs)
    print("chacha coefficient-basis batch ok", len(proofs))
    os.environ["G16_EVAL_Z"] = "1"
    ctx2 = G.Groth16Context(pk, r1)
    p2, _ = ctx2.prove_chacha_batch(*reqs)
    assert p2 == proofs
    print("chacha evaluation-basis batch ok")
    ctx2.close()
    del os.environ["G16_EVAL_Z"]
    if "verify" in what:
        ver = G.Groth16Verifier(open("tests/golden/vk.chacha20", "rb").read())
        k, no, c, i, r = ctx._pack(*reqs)[1:]
        pub = bench.chacha_public_inputs_be(np.frombuffer(b"".join(cts), dtype=np.uint8), no, c, i)
        ok = ver.verify_batch(proofs, [[int(pub[j, k, 31]) for k in range(1152)] for j in range(n)])   # the public witness is bits
        assert ok.all()
        print("verify ok", ok.tolist())
        ver.close()
    ctx.close()
if "aes" in what and os.path.exists("tests/golden/_gen/pk.aes128"):
    pk = open("tests/golden/_gen/pk.aes128", "rb").read(); r1 = open("tests/golden/r1cs.aes128", "rb").read()
    ctx = G.Groth16Context(pk, r1)
    n = 2
    proofs, cts = ctx.prove_aes_batch([rng.bytes(16) for _ in range(n)], [rng.bytes(12) for _ in range(n)], [1, 7], [rng.bytes(64) for _ in range(n)])
    print("aes batch ok", len(proofs[0]))
    ctx.close()
if "msm" in what:
    g1 = np.zeros((1, 8), dtype=np.uint64)
    g1[0, :4] = G.field_op(0, "to_mont", np.array([[1, 0, 0, 0]], dtype=np.uint64))[0]
    g1[0, 4:] = G.field_op(0, "to_mont", np.array([[2, 0, 0, 0]], dtype=np.uint64))[0]
    n = 3000
    a = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.int64).astype(np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1)
    pts = G.group_op(1, "mul", np.repeat(g1, n, axis=0), a)
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.int64).astype(np.uint64); sc[:, 3] &= np.uint64((1 << 60) - 1)
    sc[:40] = 0; sc[40:200, 1:] = 0; sc[40:200, 0] = 1
    r1_, _ = G.msm(1, pts, sc)
    plan = G.MsmPlan(1, pts, precompute=True); plan.set_scalars(sc); r2_, _ = plan.run(); plan.close()
    assert np.array_equal(r1_, r2_)
    print("msm ok")
if "ntt" in what:
    for n in (256, 1 << 12, 1 << 15):
        ms, bad = G.ntt_bench(n, 3, 1)
        assert bad == 0
    print("ntt ok")
print("sanitize_run done")
