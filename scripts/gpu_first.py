"""First GPU check-out: stage-level parity vs the oracle, KAT proof, batch timing, IMAD peak. Writes gpurun_out/first.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
import gnark_symmetric_crypto_b200 as G

out = {}
rng = np.random.default_rng(1)
t0 = time.time()
out["imad"] = G.imad_peak(); print("imad", out["imad"], flush=True)
# fields
for fld in (0, 1):
    n = 4096
    a = O.to_mont(fld, O.rand_field(rng, fld, n)); b = O.to_mont(fld, O.rand_field(rng, fld, n))
    a[0] = 0; b[1] = 0
    for op in ("add", "sub", "mul", "inv", "sqr", "neg"):
        got = G.field_op(fld, op, a, b if op in ("add", "sub", "mul") else None)
        ref = O.f_op(fld, op, a, b if op in ("add", "sub", "mul") else None)
        assert np.array_equal(got, ref), (fld, op)
print("field ok", flush=True)
P1 = O.g1_fixed_base(O.rand_field(rng, 1, 64)); Q1 = O.g1_fixed_base(O.rand_field(rng, 1, 64)); Q1[0] = P1[0]; Q1[2] = 0
for op in ("add", "add_xyzz"):
    got = G.group_op(1, op, P1, Q1); ref = np.array([O.g1_add(P1[i], Q1[i]) for i in range(64)]); assert np.array_equal(got, ref), op
P2 = O.g2_fixed_base(O.rand_field(rng, 1, 16)); Q2 = O.g2_fixed_base(O.rand_field(rng, 1, 16)); Q2[0] = P2[0]
for op in ("add", "add_xyzz"):
    got = G.group_op(2, op, P2, Q2); ref = np.array([O.g2_add(P2[i], Q2[i]) for i in range(16)]); assert np.array_equal(got, ref), op
assert np.array_equal(G.decompress(1, O.g1_compress(P1)), P1); assert np.array_equal(G.decompress(2, O.g2_compress(P2)), P2)
print("group ok", flush=True)
for n, c in ((1, 0), (300, 7), (5000, 0), (1 << 16, 0)):
    pts = O.g1_fixed_base(O.rand_field(rng, 1, n)); sc = O.rand_field(rng, 1, n)
    got, ms = G.msm(1, pts, sc, False, c); assert np.array_equal(got, O.g1_msm(pts, sc)), ("msm", n)
    print("msm g1 ok", n, ms, flush=True)
    out[f"msm_g1_{n}"] = [float(x) for x in ms]
pts = O.g2_fixed_base(O.rand_field(rng, 1, 2000)); sc = O.rand_field(rng, 1, 2000)
got, ms = G.msm(2, pts, sc); assert np.array_equal(got, O.g2_msm(pts, sc)); print("msm g2 ok", ms, flush=True)
for n in (2, 256, 1 << 12, 1 << 15, 1 << 17):
    x = O.to_mont(1, O.rand_field(rng, 1, n))
    y, ms = G.ntt(x); assert np.array_equal(y, O.ntt(x)), n
    z, ms2 = G.ntt(y, inverse=True); assert np.array_equal(z, x), n
    print("ntt ok", n, ms, ms2, flush=True)
for n, batch in ((1 << 15, 448), (1 << 20, 8), (1 << 24, 1)):
    ms, bad = G.ntt_bench(n, batch, 5); out[f"ntt_{n}x{batch}"] = {"ms_per_transform": ms, "mismatch": bad, "GBs": 64 * n * batch / ms / 1e6}
    print("ntt bench", n, batch, out[f"ntt_{n}x{batch}"], flush=True)
pkb = open("tests/golden/pk.chacha20", "rb").read(); r1b = open("tests/golden/r1cs.chacha20", "rb").read()
t = time.time(); ctx = G.Groth16Context(pkb, r1b); out["init_s"] = time.time() - t; print("init", out["init_s"], flush=True)
key = bytes([2]) * 32; nonce = bytes([3]) * 12
pt = bytes.fromhex("a3f7e592aeda1507a7f51b35812dfc50a263d5a6d2df625e563b02e49c08bf30d0e7483f5b13ff079532224ee8fbc31ab1899b18e453d36d9793a8355eb0dee9")
rr = int("11" * 20, 16); ss = int("22" * 20, 16); rs = rr.to_bytes(32, "big") + ss.to_bytes(32, "big")
kat = "d73f52bc6800d1c4a07c95d0b21876d2ed029d442b2df690a2fe2a711b77f6e195200aa0384e7f1ea31d47954f1fa672350b66ecf2608c5db062aa7f9ff7153b0f0896aa890cca5296834e8bf266931d6df1f412d99cfcf9d5786b7f3e7cb441ddfbcad67781ec5c223db4246c4c4e3860638f210422b7c296e145b1df4153f8000000004000000000000000000000000000000000000000000000000000000000000000"
proofs, cts = ctx.prove_chacha_batch([key], [nonce], [3], [pt], [rs])
out["kat_match"] = proofs[0].hex() == kat; print("KAT", out["kat_match"], ctx.stage_ms(), flush=True)
if not out["kat_match"]:
    orc = O.ChaChaOracleProver(pkb, r1b)
    pr, ct, det = orc.prove(key, nonce, 3, pt, rr, ss, detail=True)
    inputs, _ = O.chacha_assignment(key, nonce, 3, pt)
    wit = O.to_mont(1, O.ints_to_limbs(inputs[1:]))
    W, A, B, Cc = ctx.solve(wit)
    print("W", np.array_equal(W[0], det["W"]), "A", np.array_equal(A[0], det["A"]), "B", np.array_equal(B[0], det["B"]), "C", np.array_equal(Cc[0], det["C"]))
    h = ctx.compute_h(det["A"], det["B"], det["C"]); print("h", np.array_equal(h, det["h"][O.bitrev_perm(ctx.n)]))
    p2, d2 = ctx.prove_witness(wit, rs, detail=True)
    for k in ("msmA", "msmB1", "msmK", "msmZ", "msmB2"): print(k, np.array_equal(d2[k], det[k]))
for nb in (8, 64, 256):
    ks = [rng.bytes(32) for _ in range(nb)]; ns = [rng.bytes(12) for _ in range(nb)]; cs = [int(x) for x in rng.integers(0, 1 << 32, nb)]
    ins = [rng.bytes(64) for _ in range(nb)]; rss = [rs] * nb
    t = time.time(); proofs, cts = ctx.prove_chacha_batch(ks, ns, cs, ins, rss); dt = time.time() - t
    st = ctx.stage_ms(); out[f"batch_{nb}"] = {"wall_s": dt, "stages": st, "proofs_per_s_dev": nb / (st["total"] / 1e3)}
    print("batch", nb, dt, st, flush=True)
    if nb == 8:
        orc = O.ChaChaOracleProver(pkb, r1b)
        okc = 0
        for i in range(nb):
            pr, ct = orc.prove(ks[i], ns[i], cs[i], ins[i], rr, ss)
            okc += (pr == proofs[i]) and (ct == cts[i])
        out["batch8_match"] = okc; print("batch8 match", okc, "/ 8", flush=True)
out["elapsed"] = time.time() - t0
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/first.json", "w"), indent=1)
print("DONE", out["elapsed"])
