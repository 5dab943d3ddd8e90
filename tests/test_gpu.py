"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes), against the oracle on the same seeded
inputs; the committed golden vectors; and size-independent properties at BASELINE.json's full sizes.
Bit-exact everywhere — all arithmetic is integer (254-bit prime fields)."""
import struct

import numpy as np
import pytest

from conftest import batch_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gnark_symmetric_crypto_b200 as G
    return G


@pytest.mark.parametrize("fld", [0, 1])
def test_field_ops(G, oracle, fld):
    rng = np.random.default_rng(200 + fld)
    n = 1 << 14
    mod = oracle.P_MOD if fld == 0 else oracle.R_MOD
    a = oracle.to_mont(fld, oracle.rand_field(rng, fld, n)); b = oracle.to_mont(fld, oracle.rand_field(rng, fld, n))
    a[0] = 0; b[1] = 0
    a[2] = b[2] = oracle.to_mont(fld, oracle.ints_to_limbs([mod - 1]))[0]
    a[3] = oracle.to_mont(fld, oracle.ints_to_limbs([1]))[0]
    # limbs of all-ones / carries through every limb
    a[4] = oracle.ints_to_limbs([mod - 1])[0]; b[4] = oracle.ints_to_limbs([mod - 2])[0]
    for op in ("add", "sub", "mul", "sqr", "neg"):
        bb = b if op in ("add", "sub", "mul") else None
        assert np.array_equal(G.field_op(fld, op, a, bb), oracle.f_op(fld, op, a, bb)), op
    assert np.array_equal(G.field_op(fld, "inv", a[:512]), oracle.f_op(fld, "inv", a[:512]))
    can = oracle.rand_field(rng, fld, 1024)
    m = G.field_op(fld, "to_mont", can)
    assert np.array_equal(m, oracle.to_mont(fld, can))
    assert np.array_equal(G.field_op(fld, "from_mont", m), can)


def test_group_ops(G, oracle):
    rng = np.random.default_rng(17)
    n = 128
    P = oracle.g1_fixed_base(oracle.rand_field(rng, 1, n)); Q = oracle.g1_fixed_base(oracle.rand_field(rng, 1, n))
    Q[0] = P[0]; Q[1] = P[1]; Q[1][4:8] = oracle.f_op(0, "neg", P[1][4:8].reshape(1, 4))[0]; Q[2] = 0; P[3] = 0
    ref = np.array([oracle.g1_add(P[i], Q[i]) for i in range(n)])
    for op in ("add", "add_xyzz", "add_team"):
        assert np.array_equal(G.group_op(1, op, P, Q), ref), op
    # four-warp forms with both operands projective (csrc/team.cuh): 4a + 2b with b = 2a, b = -2a, infinities
    Q5 = Q.copy()
    Q5[0] = oracle.g1_add(P[0], P[0])
    Q5[1] = oracle.g1_add(P[1], P[1]); Q5[1][4:8] = oracle.f_op(0, "neg", Q5[1][4:8].reshape(1, 4))[0]
    ref5 = np.array([oracle.g1_add(oracle.g1_mul(P[i], 4), oracle.g1_mul(Q5[i], 2)) for i in range(n)])
    assert np.array_equal(G.group_op(1, "dbl_add_team", P, Q5), ref5) and not ref5[1].any()
    sc = oracle.rand_field(rng, 1, n)
    ref = np.array([oracle.g1_mul(P[i], oracle.limbs_to_ints(sc[i:i + 1])[0]) for i in range(n)])
    assert np.array_equal(G.group_op(1, "mul", P, sc), ref)
    P2 = oracle.g2_fixed_base(oracle.rand_field(rng, 1, 32)); Q2 = oracle.g2_fixed_base(oracle.rand_field(rng, 1, 32)); Q2[0] = P2[0]
    ref2 = np.array([oracle.g2_add(P2[i], Q2[i]) for i in range(32)])
    for op in ("add", "add_xyzz"):
        assert np.array_equal(G.group_op(2, op, P2, Q2), ref2), op
    ref2 = np.array([oracle.g2_mul(P2[i], oracle.limbs_to_ints(sc[i:i + 1])[0]) for i in range(32)])
    assert np.array_equal(G.group_op(2, "mul", P2, sc[:32]), ref2)


def test_decompress_whole_proving_key(G, oracle, oracle_prover, pk_bytes):
    from oracle import formats
    lay = formats.parse_pk_layout(pk_bytes)
    pk = oracle_prover.pk
    for name in ("A", "B", "Z", "K"):
        raw = pk_bytes[lay.offs[name]:lay.offs[name] + 32 * lay.counts[name]]
        assert np.array_equal(G.decompress(1, raw), pk.array(name)), name
    raw = pk_bytes[lay.offs["B2"]:lay.offs["B2"] + 64 * lay.counts["B2"]]
    assert np.array_equal(G.decompress(2, raw), pk.array("B2"))
    with pytest.raises(G.ProverError):
        G.decompress(1, b"\x00" * 32)   # uncompressed flag in a compressed slot


@pytest.mark.parametrize("n,c", [(1, 0), (2, 0), (300, 7), (5000, 0), (5000, 16), (1 << 16, 0)])
def test_msm_g1(G, oracle, n, c):
    rng = np.random.default_rng(n + c)
    pts = oracle.g1_fixed_base(oracle.rand_field(rng, 1, n)); sc = oracle.rand_field(rng, 1, n)
    if n >= 300:
        sc[0] = 0; sc[1] = oracle.ints_to_limbs([1])[0]; sc[2] = oracle.ints_to_limbs([oracle.R_MOD - 1])[0]
        pts[5] = pts[4]; sc[5] = sc[4]; pts[7] = 0
        sc[10:130] = oracle.ints_to_limbs([1] * 120); sc[130:200] = oracle.ints_to_limbs([oracle.R_MOD - 1] * 70)
    ref = oracle.g1_msm(pts, sc)
    got, _ = G.msm(1, pts, sc, False, c)
    assert np.array_equal(got, ref)
    got, _ = G.msm(1, pts, oracle.to_mont(1, sc), True, c)
    assert np.array_equal(got, ref)


def test_msm_g2_and_edges(G, oracle):
    rng = np.random.default_rng(23)
    pts = oracle.g2_fixed_base(oracle.rand_field(rng, 1, 3000)); sc = oracle.rand_field(rng, 1, 3000)
    got, _ = G.msm(2, pts, sc)
    assert np.array_equal(got, oracle.g2_msm(pts, sc))
    p1 = oracle.g1_fixed_base(oracle.rand_field(rng, 1, 64))
    got, _ = G.msm(1, p1, np.zeros((64, 4), dtype=np.uint64))
    assert not got.any()
    same = np.repeat(p1[:1], 64, axis=0).copy(); s = oracle.rand_field(rng, 1, 64)
    got, _ = G.msm(1, same, s, False, 4)
    assert np.array_equal(got, oracle.g1_msm(same, s))


def test_msm_2_20_against_field_only_oracle(G, oracle):
    """BASELINE config 5 at 2^20: P_i = a_i*G (a small set of distinct points tiled), result must equal (sum a_i s_i)*G."""
    rng = np.random.default_rng(5)
    n, distinct = 1 << 20, 1 << 12
    a = oracle.rand_field(rng, 1, distinct)
    base = oracle.g1_fixed_base(a)
    idx = rng.integers(0, distinct, n)
    pts = base[idx]
    sc = oracle.rand_field(rng, 1, n)
    ai = np.array(oracle.limbs_to_ints(a), dtype=object)[idx]
    tot = int(sum(int(x) * int(y) for x, y in zip(ai, oracle.limbs_to_ints(sc))) % oracle.R_MOD)
    got, ms = G.msm(1, pts, sc)
    assert np.array_equal(got, oracle.g1_mul(oracle.g1_gen(), tot))


@pytest.mark.parametrize("n", [2, 256, 1 << 11, 1 << 12, 1 << 15, 1 << 17, 1 << 20])
def test_ntt(G, oracle, n):
    rng = np.random.default_rng(n)
    x = oracle.to_mont(1, oracle.rand_field(rng, 1, n))
    y, _ = G.ntt(x)
    assert np.array_equal(y, oracle.ntt(x))
    z, _ = G.ntt(y, inverse=True)
    assert np.array_equal(z, x)
    yc, _ = G.ntt(x, coset=True)
    zc, _ = G.ntt(yc, inverse=True, coset=True)
    assert np.array_equal(zc, x)


def test_ntt_2_24_round_trip_and_point_checks(G, oracle):
    n = 1 << 24
    ms, mismatches = G.ntt_bench(n, 1, 1)
    assert mismatches == 0
    # 8 random output positions of a forward transform against direct evaluation sum x_j w^(ij) on a sparse input
    rng = np.random.default_rng(24)
    x = np.zeros((n, 4), dtype=np.uint64)
    pos = rng.integers(0, n, 64)
    vals = oracle.rand_field(rng, 1, 64)
    x[pos] = oracle.to_mont(1, vals)
    y, _ = G.ntt(x)
    w = oracle.root_of_unity(n)
    yi = oracle.from_mont(1, y)
    vi = oracle.limbs_to_ints(vals)
    last = {int(p): v for p, v in zip(pos, vi)}   # duplicate positions: last write wins, as in numpy
    for i in rng.integers(0, n, 8):
        want = sum(v * pow(w, (int(i) * p) % n, oracle.R_MOD) for p, v in last.items()) % oracle.R_MOD
        assert oracle.limbs_to_ints(yi[int(i):int(i) + 1])[0] == want


def test_solver_and_h_match_oracle(G, gpu_ctx, oracle, oracle_prover, kat):
    rng = np.random.default_rng(31)
    reqs = [(kat["key"], kat["nonce"], kat["counter"], kat["input"])] + \
           [(rng.bytes(32), rng.bytes(12), int(rng.integers(0, 1 << 32)), rng.bytes(64)) for _ in range(3)]
    wit, refs = [], []
    for k, n, c, i in reqs:
        inputs, _ = oracle.chacha_assignment(k, n, c, i)
        wit.append(oracle.to_mont(1, oracle.ints_to_limbs(inputs[1:])))
        refs.append(oracle_prover.cs.solve(inputs))
    # batch 4 takes the term-parallel kernel (a warp per instruction and witness, single-request latency path), batch 36
    # the witness-parallel one (a lane per witness): both must reproduce the oracle's wires and A/B/C evaluations
    W, A, B, Cc = gpu_ctx.solve(np.stack(wit), batch=4)
    for j in range(4):
        for got, ref in zip((W[j], A[j], B[j], Cc[j]), refs[j]):
            assert np.array_equal(got, ref)
    W, A, B, Cc = gpu_ctx.solve(np.stack(wit * 9), batch=36)
    for j in range(36):
        for got, ref in zip((W[j], A[j], B[j], Cc[j]), refs[j % 4]):
            assert np.array_equal(got, ref)
    n = gpu_ctx.n
    h = gpu_ctx.compute_h(refs[0][1], refs[0][2], refs[0][3])
    assert np.array_equal(h, oracle.compute_h(refs[0][1], refs[0][2], refs[0][3], n)[oracle.bitrev_perm(n)])
    bad = np.stack(wit)
    bad[1, 700] = oracle.to_mont(1, oracle.ints_to_limbs([1 - oracle.limbs_to_ints(oracle.from_mont(1, bad[1, 700:701]))[0]]))[0]
    with pytest.raises(G.ProverError) as e:
        gpu_ctx.solve(bad, batch=4)
    assert e.value.rc == 4   # G16_ERR_UNSAT
    with pytest.raises(G.ProverError) as e:
        gpu_ctx.solve(np.concatenate([bad] * 9), batch=36)
    assert e.value.rc == 4


def test_kat_proof_msm_points_and_h(G, gpu_ctx, oracle, oracle_prover, oracle_vk, kat):
    """BASELINE config 1: the reference's benchmark input (core_test.go:285), fixed r,s: serialized proof, the five MSM
    points and every H coefficient equal the oracle's; the proof equals the committed KAT and vk.chacha20 accepts it."""
    rs = kat["r"].to_bytes(32, "big") + kat["s"].to_bytes(32, "big")
    inputs, ct = oracle.chacha_assignment(kat["key"], kat["nonce"], kat["counter"], kat["input"])
    wit = oracle.to_mont(1, oracle.ints_to_limbs(inputs[1:]))
    proof, det = gpu_ctx.prove_witness(wit, rs, detail=True)
    assert proof == kat["proof"]
    pr, _, ref = oracle_prover.prove(kat["key"], kat["nonce"], kat["counter"], kat["input"], kat["r"], kat["s"], detail=True)
    assert pr == proof
    for k in ("msmA", "msmB1", "msmK", "msmZ", "msmB2"):
        assert np.array_equal(det[k], ref[k]), k
    assert np.array_equal(det["h"], ref["h"][oracle.bitrev_perm(gpu_ctx.n)])
    assert oracle_vk.verify(proof, inputs[1:gpu_ctx.nb_public])
    # library-shaped entry gives the same bytes and the ciphertext
    proofs, cts = gpu_ctx.prove_chacha_batch([kat["key"]], [kat["nonce"]], [kat["counter"]], [kat["input"]], [rs])
    assert proofs[0] == proof and cts[0] == kat["ct"] == ct
    # fresh randomness (rs = None): different bytes, still accepted
    p2 = gpu_ctx.prove_witness(wit, None)
    assert p2 != proof and oracle_vk.verify(p2, inputs[1:gpu_ctx.nb_public])


def test_libprove_abi_round_trip(G, oracle, oracle_vk, pk_bytes, r1cs_bytes):
    """The reference's own integration test shape (core_test.go:130-172 TestFullChaCha20, :120-128 TestPanic) through
    the libprove-compatible exports: InitAlgorithm -> Prove(JSON) -> verifier accepts."""
    assert G.InitAlgorithm(G.CHACHA20, pk_bytes, r1cs_bytes) is True
    assert G.InitAlgorithm(G.CHACHA20, pk_bytes, r1cs_bytes) is True   # initDone short-cut
    assert G.InitAlgorithm(9, pk_bytes, r1cs_bytes) is False
    rng = np.random.default_rng(77)
    key, nonce, pt, counter = rng.bytes(32), rng.bytes(12), rng.bytes(64), 1
    out = G.OutputParams.from_json(G.Prove(G.InputParams("chacha20", key, nonce, counter, pt).to_json()))
    assert len(out.proof_json) == 164
    signals = out.public_signals + nonce + struct.pack("<I", counter) + pt   # core_test.go:157-163
    assert out.public_signals == oracle.chacha20_xor(key, nonce, counter, pt)
    assert oracle_vk.verify(out.proof_json, oracle.chacha_public_from_signals(signals))
    with pytest.raises(RuntimeError, match="could not find prover"):
        G.Prove(b'{"cipher":"aes-256-ctr1","key":[0],"nonce":[0],"counter":1,"input":[0]}')
    with pytest.raises(RuntimeError, match="counter"):   # TestPanic's payload has an array there: json.Unmarshal fails
        G.Prove(b'{"cipher":"aes-256-ctr1","key":[0],"nonce":[0],"counter":[0,1],"input":[0]}')
    with pytest.raises(RuntimeError, match="key length must be 32"):
        G.Prove(G.InputParams("chacha20", key[:31], nonce, counter, pt).to_json())
    with pytest.raises(RuntimeError, match="plaintext length must be 64"):
        G.Prove(G.InputParams("chacha20", key, nonce, counter, pt + b"x").to_json())
    with pytest.raises(RuntimeError, match="not initialized"):
        G.Prove(G.InputParams("aes-128-ctr", key[:16], nonce, counter, pt).to_json())


def test_batch_1024_config4(G, gpu_ctx, oracle, oracle_prover, oracle_vk):
    """BASELINE config 4: 1 024 seeded requests in one batch. Every ciphertext is checked; a 1/16 sample of the proofs is
    compared byte-for-byte with the oracle and pairing-verified against the reference's vk.chacha20; all proofs must be
    distinct and well-formed."""
    n = 1024
    keys, nonces, ctrs, ins, rs = batch_inputs(n)
    proofs, cts = gpu_ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)
    assert len(set(proofs)) == n
    for i in range(n):
        assert cts[i] == oracle.chacha20_xor(keys[i], nonces[i], ctrs[i], ins[i])
        assert proofs[i][128:] == b"\x00\x00\x00\x00\x40" + b"\x00" * 31 and proofs[i][0] & 0x80
    rng = np.random.default_rng(4)
    for i in sorted(rng.choice(n, n // 16, replace=False).tolist()):
        r = int.from_bytes(rs[i][:32], "big"); s = int.from_bytes(rs[i][32:], "big")
        pr, ct = oracle_prover.prove(keys[i], nonces[i], ctrs[i], ins[i], r, s)
        assert pr == proofs[i], i
        signals = cts[i] + nonces[i] + struct.pack("<I", ctrs[i]) + ins[i]
        assert oracle_vk.verify(proofs[i], oracle.chacha_public_from_signals(signals)), i
    # sub-batching must not matter: proving request 5 alone gives the same bytes
    p5, _ = gpu_ctx.prove_chacha_batch(keys[5:6], nonces[5:6], ctrs[5:6], ins[5:6], rs[5:6])
    assert p5[0] == proofs[5]


def test_evaluation_basis_z_query_is_bit_identical(G, gpu_ctx, oracle, pk_bytes, r1cs_bytes, kat, monkeypatch):
    """Large batches take the Z query over the evaluation-basis tables (four transforms, H never materialised).
    Forced on for a single request it must reproduce the KAT proof; forced off for a batch above the switch-over it must
    give the bytes the default path gives."""
    rs = kat["r"].to_bytes(32, "big") + kat["s"].to_bytes(32, "big")
    monkeypatch.setenv("G16_EVAL_Z", "1")
    on = G.Groth16Context(pk_bytes, r1cs_bytes, device=0)
    monkeypatch.setenv("G16_EVAL_Z", "0")
    off = G.Groth16Context(pk_bytes, r1cs_bytes, device=0)
    monkeypatch.delenv("G16_EVAL_Z")
    try:
        proofs, cts = on.prove_chacha_batch([kat["key"]], [kat["nonce"]], [kat["counter"]], [kat["input"]], [rs])
        assert proofs[0] == kat["proof"] and cts[0] == kat["ct"]
        n = 300
        keys, nonces, ctrs, ins, rss = batch_inputs(n, b"g16-b200-evalz")
        p_def, _ = gpu_ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rss)     # n >= 128: evaluation basis
        p_off, _ = off.prove_chacha_batch(keys, nonces, ctrs, ins, rss)         # coefficient basis (compute_h + pk.G1.Z)
        p_on, _ = on.prove_chacha_batch(keys[:40], nonces[:40], ctrs[:40], ins[:40], rss[:40])
        assert p_def == p_off
        assert p_on == p_off[:40]
        # a caller that reads H back gets the coefficients even from a context forced onto the evaluation basis, and the
        # Z-query point is the same group element on both paths
        inputs, _ = oracle.chacha_assignment(kat["key"], kat["nonce"], kat["counter"], kat["input"])
        wit = oracle.to_mont(1, oracle.ints_to_limbs(inputs[1:]))
        pr_on, det_on = on.prove_witness(wit, rs, detail=True)
        pr_ref, det_ref = gpu_ctx.prove_witness(wit, rs, detail=True)
        assert pr_on == pr_ref == kat["proof"]
        assert np.array_equal(det_on["h"], det_ref["h"]) and np.array_equal(det_on["msmZ"], det_ref["msmZ"])
        assert on.prove_witness(wit, rs) == kat["proof"]          # no detail: evaluation basis, same bytes
    finally:
        on.close()
        off.close()


def test_evaluation_basis_tables_two_builders(G, pk_bytes, r1cs_bytes, kat, monkeypatch):
    """The evaluation-basis tables are built as a DFT over group elements; the cross-check builder derives every table point
    as an MSM over pk.G1.Z of a column of compute_h's own linear map. Both must give the KAT proof."""
    rs = kat["r"].to_bytes(32, "big") + kat["s"].to_bytes(32, "big")
    monkeypatch.setenv("G16_EVAL_Z", "1")
    monkeypatch.setenv("G16_EVAL_BUILD_MSM", "1")
    ctx = G.Groth16Context(pk_bytes, r1cs_bytes, device=0)
    try:
        proofs, _ = ctx.prove_chacha_batch([kat["key"]], [kat["nonce"]], [kat["counter"]], [kat["input"]], [rs])
        assert proofs[0] == kat["proof"]
        assert ctx.counters()["eval_basis_z"]
    finally:
        ctx.close()


def test_pairing_check_matches_oracle(G, oracle):
    """SURVEY §8f rank 4 (the arithmetic under groth16.Verify): products of 1..4 pairings that are 1 by bilinearity are
    accepted, perturbed ones rejected, infinity pairs contribute 1 — the same verdicts as the oracle's pairing."""
    rng = np.random.default_rng(99)
    R = oracle.R_MOD
    Ps, Qs, want = [], [], []
    for k in (1, 3):
        a = [int(x) for x in rng.integers(1, 1 << 62, k)]; b = [int(x) for x in rng.integers(1, 1 << 62, k)]
        g1 = oracle.g1_fixed_base(oracle.ints_to_limbs(a + [(-sum(x * y for x, y in zip(a, b))) % R] + [0] * (3 - k)))
        g2 = oracle.g2_fixed_base(oracle.ints_to_limbs(b + [1] + [7] * (3 - k)))   # padding pairs: P = infinity
        for flip in (False, True):
            p = g1.copy()
            if flip:
                p[0] = oracle.g1_fixed_base(oracle.ints_to_limbs([a[0] + 1]))[0]
            Ps.append(p); Qs.append(g2); want.append(not flip)
            assert oracle.pairing_check(p[:k + 1], g2[:k + 1]) == (not flip)
    got = G.pairing_check(np.concatenate(Ps), np.concatenate(Qs), pairs_per_check=4)
    assert got.tolist() == want


def test_msm_plan_with_precomputed_tables(G, oracle):
    """Fixed-base mode of the standalone MSM (the prover context's mode, exposed for BASELINE config 5): same points as the
    one-shot pipeline and as the oracle, for G1 and G2, automatic and explicit windows, scalars replaced between runs."""
    rng = np.random.default_rng(61)
    n = 3000
    pts = oracle.g1_fixed_base(oracle.rand_field(rng, 1, n)); pts[7] = 0
    for window in (0, 9):
        plan = G.MsmPlan(1, pts, window=window, precompute=True)
        for _ in range(2):
            sc = oracle.rand_field(rng, 1, n); sc[3] = 0
            plan.set_scalars(sc)
            out, _ms = plan.run()
            assert np.array_equal(out, oracle.g1_msm(pts, sc))
        plan.close()
    p2 = oracle.g2_fixed_base(oracle.rand_field(rng, 1, 200)); s2 = oracle.rand_field(rng, 1, 200)
    plan = G.MsmPlan(2, p2, window=7, precompute=True)
    plan.set_scalars(s2)
    out, _ms = plan.run()
    assert np.array_equal(out, oracle.g2_msm(p2, s2))
    plan.close()


def test_concurrent_prove_calls_are_batched(G, oracle, oracle_vk, pk_bytes, r1cs_bytes):
    """The reference's API takes one request per Prove call and its callers issue calls concurrently; the library coalesces
    concurrent calls into GPU batches. 96 threads x 2 calls: every caller gets the proof of ITS request (ciphertext and
    pairing check against vk.chacha20), and a failing call (bad key length) does not disturb the others."""
    import threading
    assert G.InitAlgorithm(G.CHACHA20, pk_bytes, r1cs_bytes) is True
    nthreads, per = 96, 2
    rng = np.random.default_rng(1234)
    reqs = [[(rng.bytes(32), rng.bytes(12), int(rng.integers(0, 1 << 32)), rng.bytes(64)) for _ in range(per)] for _ in range(nthreads)]
    results = [[None] * per for _ in range(nthreads)]
    errors = []

    def worker(t):
        try:
            for j, (key, nonce, counter, pt) in enumerate(reqs[t]):
                if t == 5 and j == 0:
                    with pytest.raises(RuntimeError, match="key length must be 32"):
                        G.Prove(G.InputParams("chacha20", key[:31], nonce, counter, pt).to_json())
                results[t][j] = G.OutputParams.from_json(G.Prove(G.InputParams("chacha20", key, nonce, counter, pt).to_json()))
        except Exception as e:   # noqa: BLE001
            errors.append((t, repr(e)))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
    [x.start() for x in th]; [x.join() for x in th]
    assert not errors, errors[:3]
    proofs = set()
    for t in range(nthreads):
        for j, (key, nonce, counter, pt) in enumerate(reqs[t]):
            out = results[t][j]
            assert out.public_signals == oracle.chacha20_xor(key, nonce, counter, pt)
            proofs.add(out.proof_json)
            if (t + j) % 16 == 0:
                signals = out.public_signals + nonce + struct.pack("<I", counter) + pt
                assert oracle_vk.verify(out.proof_json, oracle.chacha_public_from_signals(signals))
    assert len(proofs) == nthreads * per
    # concurrent Verify calls are coalesced the same way: every caller gets the verdict of ITS proof
    assert G.InitVerifier(G.CHACHA20, open("tests/golden/vk.chacha20", "rb").read()) is True
    verdicts = [[None] * per for _ in range(nthreads)]

    def vworker(t):
        for j, (key, nonce, counter, pt) in enumerate(reqs[t]):
            out = results[t][j]
            signals = bytearray(out.public_signals + nonce + struct.pack("<I", counter) + pt)
            if (t + j) % 7 == 0:
                signals[t % 64] ^= 0x10   # a wrong ciphertext byte: must be rejected
            verdicts[t][j] = G.Verify(G.InputVerifyParams("chacha20", out.proof_json, bytes(signals)).to_json())

    th = [threading.Thread(target=vworker, args=(t,)) for t in range(nthreads)]
    [x.start() for x in th]; [x.join() for x in th]
    for t in range(nthreads):
        for j in range(per):
            assert verdicts[t][j] is ((t + j) % 7 != 0), (t, j)
