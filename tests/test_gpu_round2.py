"""GPU tests added in round 2: BASELINE config 5 at its full sizes (MSM 2^22 / 2^24 against the field-only oracle, the
8-range point split combined on the host), the multi-device handle (sharding request i -> device i mod G, exercised on one
GPU by listing it twice), per-request verdicts of a batch, the ProveBatch export, and the canonical-encoding checks of the
verifier's point decoder. Everything goes through the C-ABI; bit-exact."""
import json
import os
import struct
import threading

import numpy as np
import pytest

from conftest import batch_inputs, aes_keys

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gnark_symmetric_crypto_b200 as G
    return G


# ---------------------------------------------------------------------------------------------- config 5, full sizes
def _tiled_msm_inputs(oracle, n, seed, distinct=1 << 12):
    """P_i = a_i G1, uniform scalars; expected = (sum a_i s_i mod r) G1, an O(N) field-only oracle.
    The a_i are tiled from `distinct` values so that the points cost 4096 CPU scalar products, and the big sum is grouped
    per distinct point: sum_d a_d (sum_{i -> d} s_i), the inner sums taken on 32-bit limb columns with numpy."""
    rng = np.random.default_rng(seed)
    a = oracle.rand_field(rng, 1, distinct)
    base = oracle.g1_fixed_base(a)
    idx = rng.integers(0, distinct, n)
    sc = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.int64).astype(np.uint64)
    sc[:, 3] &= np.uint64((1 << 60) - 1)                                  # < r
    limbs = sc.view(np.uint32).reshape(n, 8).astype(np.uint64)            # little-endian 32-bit limbs
    col = np.zeros((distinct, 8), dtype=np.uint64)
    np.add.at(col, idx, limbs)                                            # <= 2^24 terms of 32 bits: fits 64 bits
    a_int = oracle.limbs_to_ints(a)
    tot = 0
    for d in range(distinct):
        s_d = sum(int(col[d, k]) << (32 * k) for k in range(8))
        tot += a_int[d] * s_d
    return base, idx, sc, tot % oracle.R_MOD


@pytest.mark.parametrize("lg", [22, 24])
def test_msm_full_size_against_field_only_oracle(G, oracle, lg):
    """BASELINE config 5 at 2^22 and 2^24 points, one-shot (no tables) and, at 2^22, fixed-base."""
    n = 1 << lg
    base, idx, sc, tot = _tiled_msm_inputs(oracle, n, 50 + lg)
    want = oracle.g1_mul(oracle.g1_gen(), tot)
    pts = base[idx]
    plan = G.MsmPlan(1, pts)
    plan.set_scalars(sc)
    got, ms = plan.run()
    plan.close()
    assert np.array_equal(got, want)
    if lg == 22:
        plan = G.MsmPlan(1, pts, precompute=True)
        plan.set_scalars(sc)
        got, ms = plan.run()
        plan.close()
        assert np.array_equal(got, want)


def test_msm_point_range_split_combined_on_host(G, oracle):
    """a single MSM split by point range into 8 partial MSMs (what 8 GPUs would each compute), the 8 partial
    points added on the host side of the ABI (g16_group_op). One GPU runs the 8 ranges one after the other."""
    from bench import shard_range
    n, parts = 1 << 22, 8
    base, idx, sc, tot = _tiled_msm_inputs(oracle, n, 77)
    want = oracle.g1_mul(oracle.g1_gen(), tot)
    acc = None
    for g in range(parts):
        lo, hi = shard_range(n, g, parts)
        plan = G.MsmPlan(1, base[idx[lo:hi]])
        plan.set_scalars(sc[lo:hi])
        part, _ = plan.run()
        plan.close()
        acc = part if acc is None else G.group_op(1, "add", acc.reshape(1, 8), part.reshape(1, 8))[0]
    assert np.array_equal(acc, want)


# ---------------------------------------------------------------------------------------------- multi-device handle
def test_multi_device_handle_shards_and_gathers_in_order(G, gpu_ctx, pk_bytes, r1cs_bytes):
    """g16_init_multi with the same GPU listed twice: request i -> slot i mod 2, both slots run concurrently from their own
    host threads, proofs and ciphertexts come back in input order and equal the single-device results byte for byte."""
    n = 37                                                   # odd: the two shards differ in size
    keys, nonces, ctrs, ins, rs = batch_inputs(n, b"g16-b200-multi")
    ref_p, ref_c = gpu_ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)
    multi = G.Groth16Context(pk_bytes, r1cs_bytes, devices=[0, 0])
    try:
        assert multi.devices == [0, 0]
        p, c = multi.prove_chacha_batch(keys, nonces, ctrs, ins, rs)
        assert p == ref_p and c == ref_c
        assert multi.counters()["proofs"] == n               # work counters are summed over the devices
        p1, c1 = multi.prove_chacha_batch(keys[3:4], nonces[3:4], ctrs[3:4], ins[3:4], rs[3:4])   # one request: device slot 0 alone
        assert p1[0] == ref_p[3] and c1[0] == ref_c[3]
        p2, _ = multi.prove_chacha_batch(keys[:2], nonces[:2], ctrs[:2], ins[:2], rs[:2])
        assert p2 == ref_p[:2]
        assert not multi.batch_status(2).any()
    finally:
        multi.close()


def test_libprove_serves_concurrent_calls_from_every_device(G, oracle, pk_bytes, r1cs_bytes, monkeypatch):
    """G16_DEVICES=0,0: InitAlgorithm loads the key twice and starts one batching worker per device slot over the shared
    queue; every caller still gets the proof of its own request."""
    from gnark_symmetric_crypto_b200 import _lib
    L = _lib.load()
    L.g16_libprove_reset.restype = None
    L.g16_libprove_reset()
    monkeypatch.setenv("G16_DEVICES", "0,0")
    monkeypatch.setenv("G16_PREWARM", "64")
    try:
        assert G.InitAlgorithm(G.CHACHA20, pk_bytes, r1cs_bytes) is True
        st = np.zeros(3, dtype=np.uint64)
        assert L.g16_libprove_stats(0, st.ctypes.data_as(_lib.u64p)) == 0 and int(st[2]) == 2
        rng = np.random.default_rng(99)
        reqs = [(rng.bytes(32), rng.bytes(12), int(rng.integers(0, 1 << 32)), rng.bytes(64)) for _ in range(64)]
        outs = [None] * len(reqs)

        def worker(t):
            k, no, c, pt = reqs[t]
            outs[t] = G.OutputParams.from_json(G.Prove(G.InputParams("chacha20", k, no, c, pt).to_json()))
        th = [threading.Thread(target=worker, args=(t,)) for t in range(len(reqs))]
        [x.start() for x in th]; [x.join() for x in th]
        for (k, no, c, pt), o in zip(reqs, outs):
            assert o is not None and o.public_signals == oracle.chacha20_xor(k, no, c, pt) and len(o.proof_json) == 164
        assert L.g16_libprove_stats(0, st.ctypes.data_as(_lib.u64p)) == 0 and int(st[1]) == len(reqs)
    finally:
        L.g16_libprove_reset()


# ---------------------------------------------------------------------------------------------- per-request verdicts
def test_unsatisfiable_request_fails_alone(G, oracle, aes128_oracle):
    """One AES request whose counter range wraps (counter > 0xFFFFFFFB violates the circuit's AssertIsLessOrEqual,
    circuits/aesV2/aes128.go:50-53) among good ones: the batch runs ONCE, the call reports G16_ERR_UNSAT, the per-request
    status names the offender, every other proof is the oracle's."""
    pk, vk, r1 = aes_keys(128)
    ctx = G.Groth16Context(pk, r1)
    try:
        rng = np.random.default_rng(21)
        n, bad = 12, 7
        keys = [rng.bytes(16) for _ in range(n)]; nonces = [rng.bytes(12) for _ in range(n)]
        ctrs = [int(x) for x in rng.integers(0, 1 << 31, n)]; ins = [rng.bytes(64) for _ in range(n)]
        ctrs[bad] = 0xFFFFFFFD
        rsm = [b"".join(int(x).to_bytes(32, "big") for x in rng.integers(1, 1 << 62, 3)) for _ in range(n)]
        k = np.frombuffer(b"".join(keys), dtype=np.uint8).copy(); no = np.frombuffer(b"".join(nonces), dtype=np.uint8).copy()
        i = np.frombuffer(b"".join(ins), dtype=np.uint8).copy(); c = np.asarray(ctrs, dtype=np.uint32)
        r = np.frombuffer(b"".join(rsm), dtype=np.uint8).copy()
        proofs = np.zeros(n * 196, dtype=np.uint8); cts = np.zeros(n * 64, dtype=np.uint8)
        from gnark_symmetric_crypto_b200._lib import u8p, u32p
        p8 = lambda a: a.ctypes.data_as(u8p)
        rc = ctx._L.g16_prove_aes_batch(ctx._h, n, p8(k), 16, p8(no), c.ctypes.data_as(u32p), p8(i), p8(r), p8(proofs), p8(cts))
        assert rc == 4 and b"1 of 12" in ctx._L.g16_last_error()         # G16_ERR_UNSAT
        st = ctx.batch_status(n)
        assert [j for j in range(n) if st[j]] == [bad]
        assert not proofs[bad * 196:(bad + 1) * 196].any()
        for j in (0, bad - 1, bad + 1, n - 1):
            rr, ss, mm = (int.from_bytes(rsm[j][32 * t:32 * t + 32], "big") for t in range(3))
            want, ct = aes128_oracle.prove(keys[j], nonces[j], ctrs[j], ins[j], rr, ss, mm)
            assert proofs[j * 196:(j + 1) * 196].tobytes() == want and cts[j * 64:(j + 1) * 64].tobytes() == ct, j
    finally:
        ctx.close()


def test_prove_rejects_unprovable_aes_requests_on_the_host(G):
    """Requests the circuit cannot accept never enter a batch: the counter check and the key-size/cipher check answer
    immediately with an error-value payload (an object, as json.Marshal(err) gives in libprove.go:36-41)."""
    from gnark_symmetric_crypto_b200 import _lib
    L = _lib.load()
    L.g16_libprove_reset.restype = None
    L.g16_libprove_reset()
    pk, vk, r1 = aes_keys(128)
    try:
        assert G.InitAlgorithm(G.AES_128, pk, r1) is True
        good = G.InputParams("aes-128-ctr", bytes(16), bytes(12), 0xFFFFFFFB, bytes(64))
        out = G.OutputParams.from_json(G.Prove(good.to_json()))
        assert len(out.proof_json) == 196
        st = np.zeros(3, dtype=np.uint64)
        assert L.g16_libprove_stats(1, st.ctypes.data_as(_lib.u64p)) == 0
        proved_before = int(st[1])
        with pytest.raises(RuntimeError, match="error value"):
            G.Prove(G.InputParams("aes-128-ctr", bytes(16), bytes(12), 0xFFFFFFFC, bytes(64)).to_json())
        with pytest.raises(RuntimeError, match="error value"):
            G.Prove(G.InputParams("aes-128-ctr", bytes(32), bytes(12), 1, bytes(64)).to_json())
        with pytest.raises(RuntimeError, match="key length must be 16 or 32"):
            G.Prove(G.InputParams("aes-128-ctr", bytes(24), bytes(12), 1, bytes(64)).to_json())
        assert L.g16_libprove_stats(1, st.ctypes.data_as(_lib.u64p)) == 0 and int(st[1]) == proved_before   # nothing was queued
    finally:
        L.g16_libprove_reset()


# ---------------------------------------------------------------------------------------------- ProveBatch
def test_prove_batch_export(G, oracle, oracle_vk, pk_bytes, r1cs_bytes):
    """ProveBatch beside Prove: one JSON array in, one JSON array out, order kept, bad requests answered
    in place with what Prove would have returned, good ones verifier-accepted under the reference's vk.chacha20."""
    assert G.InitAlgorithm(G.CHACHA20, pk_bytes, r1cs_bytes) is True
    rng = np.random.default_rng(5150)
    reqs = [G.InputParams("chacha20", rng.bytes(32), rng.bytes(12), int(rng.integers(0, 1 << 32)), rng.bytes(64)) for _ in range(9)]
    payload = [json.loads(r.to_json()) for r in reqs]
    payload.insert(4, {"cipher": "chacha20", "key": [1, 2], "nonce": [], "counter": 0, "input": []})
    payload.insert(7, {"cipher": "rot13"})
    res = G.ProveBatch(payload)
    assert len(res) == 11
    assert isinstance(res[4], RuntimeError) and "key length must be 32" in str(res[4])
    assert isinstance(res[7], RuntimeError) and "could not find prover" in str(res[7])
    good = [x for j, x in enumerate(res) if j not in (4, 7)]
    for r, o in zip(reqs, good):
        out = G.OutputParams.from_json(o)
        assert out.public_signals == oracle.chacha20_xor(r.key, r.nonce, r.counter, r.input)
        signals = out.public_signals + r.nonce + struct.pack("<I", r.counter) + r.input
        assert oracle_vk.verify(out.proof_json, oracle.chacha_public_from_signals(signals))
    assert G.ProveBatch([]) == []


# ---------------------------------------------------------------------------------------------- canonical encodings
def test_verifier_rejects_noncanonical_coordinates(G, oracle, kat):
    """gnark-crypto's G1/G2 decoders use fp.Element.SetBytesCanonical: a coordinate encoded as x + p is not a valid proof
    encoding, although it reduces to the same point. Ar.x / Krs.x carry two flag bits (x + p must stay below 2^254 to be
    expressible: tried on whichever of the two fits), both halves of Bs.x are tried; infinity with a non-zero payload too."""
    vk = open("tests/golden/vk.chacha20", "rb").read()
    ver = G.Groth16Verifier(vk)
    inputs, _ = oracle.chacha_assignment(kat["key"], kat["nonce"], kat["counter"], kat["input"])
    pub = inputs[1:1153]
    P = oracle.P_MOD
    proof = kat["proof"]
    forged = []
    for off, masked in ((0, True), (96, True), (32, True), (64, False)):   # Ar.x, Krs.x, Bs.x.A1 (flag byte), Bs.x.A0
        raw = proof[off:off + 32]
        flags = raw[0] & 0xC0 if masked else 0
        x = int.from_bytes(bytes([raw[0] & 0x3F]) + raw[1:], "big") if masked else int.from_bytes(raw, "big")
        if x + P >= (1 << (254 if masked else 256)):
            continue
        enc = bytearray((x + P).to_bytes(32, "big"))
        enc[0] |= flags
        f = bytearray(proof); f[off:off + 32] = enc
        forged.append(bytes(f))
    assert len(forged) >= 1                                   # the unmasked Bs half always fits
    inf = bytearray(proof); inf[96:128] = bytes([0x40]) + bytes(30) + b"\x01"   # "infinity" Krs with a stray payload byte
    forged.append(bytes(inf))
    verdicts = ver.verify_batch([proof] + forged, [pub] * (1 + len(forged)))
    assert verdicts.tolist() == [True] + [False] * len(forged)
    ver.close()


# ---------------------------------------------------------------------------------------------- batch-affine accumulation, row sort
@pytest.mark.parametrize("env", [
    {"G16_MSM_BA_MIN": "1", "G16_MSM_BA_K": "1"},
    {"G16_MSM_BA_MIN": "1", "G16_MSM_BA_K": "2"},
    {"G16_MSM_BA_MIN": "1", "G16_MSM_BA_K": "3", "G16_BA_INV2_MIN": "1"},
    {"G16_MSM_BA_MIN": "1", "G16_MSM_BA_K": "3", "G16_MSM_BA_LEFT": "0"},
    {"G16_MSM_ROWSORT": "2"},
    {"G16_MSM_ROWSORT": "2", "G16_MSM_BA_MIN": "1", "G16_MSM_BA_K": "3"},
    {"G16_MSM_ROWSORT": "0", "G16_MSM_BA": "0"},
], ids=["batch_affine_k1", "batch_affine_k2", "batch_affine_k3_two_level_inversion", "batch_affine_k3_plain_padding", "rowsort", "rowsort_batch_affine_k3", "round1_paths"])
def test_msm_paths_forced_on_small_and_edge_cases(env):
    """The batch-affine pairwise levels (csrc/msm_ba.cuh; gnark's own bucket-addition algorithm, multiexp_affine.go:35-176)
    normally start at 2^21 entries and the per-row shared-memory sort at 32 rows. Here they are forced onto the existing small
    MSM parity tests — sizes 1 to 2^16 against the oracle, zeros, +-1 runs, duplicate points (the doubling branch), P + (-P),
    points at infinity, fixed-base tables, the KAT proof and a 37-request batch — in a child process (the switches are read
    once per process). By default a bucket's remainder of <= 5 (K = 3) or <= 2 (K = 2) entries skips the levels (direct
    leftovers, 64-bit placement cursor); one variant restores the plain padding. The last variant switches both off: the round-1
    paths stay correct."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu.py", "tests/test_gpu_round2.py", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                          "-k", "test_msm_g1 or test_msm_g2_and_edges or test_msm_plan_with_precomputed_tables or test_kat_proof "
                                "or test_multi_device_handle"],
                         capture_output=True, text=True, env=dict(os.environ, **env), timeout=1200)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout


# ---------------------------------------------------------------------------------------------- combination tables (bit wires)
@pytest.mark.parametrize("group,n,rows", [(1, 8, 3), (1, 300, 40), (1, 5000, 64), (2, 50, 4), (2, 700, 33)])
def test_bit_wire_combination_table_kernels(oracle, group, n, rows):
    """csrc/k_bitq.cu on the GPU: per-witness subset sums through the 255-entry tables of every group of 8 points against the
    oracle's MSM on the same 0 / 1 scalars (infinity, equal points, an all-zero and an all-one witness, the exception flag)."""
    from gnark_symmetric_crypto_b200 import _lib
    from test_emu import _bitq_case
    _bitq_case(_lib.load(), oracle, group, n, rows, 77 * group + n)


def test_bit_wire_tables_leave_the_proofs_unchanged(G, oracle, pk_bytes, r1cs_bytes, monkeypatch):
    """The wire-driven queries (A, B1, K, B2) through the combination tables: the context learns from its first witnesses which
    wires are bits (here 40, normally 256), builds the tables, and from then on proves batches with one table point per group
    of 8 wires. Same group elements, so the same proof bytes as the general path of the first batch; then the fallback: a
    forced exception makes the context prove the batch again on the general path and learn the classification anew."""
    monkeypatch.setenv("G16_BITQ_MIN_ROWS", "40")
    monkeypatch.setenv("G16_BITQ_MIN_BATCH", "8")
    keys, nonces, ctrs, ins, rs = batch_inputs(40, seed=b"g16-b200-bitq")
    ctx = G.Groth16Context(pk_bytes, r1cs_bytes)
    e = np.zeros(16, dtype=np.uint64)
    u64p = e.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_uint64))
    p1, c1 = ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)             # general path; the classification is learned
    assert ctx._L.g16_last_counters_ex(ctx._h, u64p) == 0 and (int(e[11]) & 0xFF, (int(e[11]) >> 8) & 1) == (0, 0)
    p2, c2 = ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)             # tables built, used
    assert ctx._L.g16_last_counters_ex(ctx._h, u64p) == 0 and (int(e[11]) & 0xFF, (int(e[11]) >> 8) & 1) == (1, 1)
    assert (int(e[11]) >> 16) & 0xFFFFFF > 2000                               # groups of the A query: binary (8) and ternary (5) ones
    assert int(e[12]) == 0                                                    # every ChaCha wire is 0 / 1 / -1: nothing left on the general path
    assert p2 == p1 and c2 == c1
    keys3, nonces3, ctrs3, ins3, rs3 = batch_inputs(24, seed=b"g16-b200-bitq-other")
    p3, _ = ctx.prove_chacha_batch(keys3, nonces3, ctrs3, ins3, rs3)         # other witnesses through the same tables
    ver = G.Groth16Verifier(open("tests/golden/vk.chacha20", "rb").read())
    pubs = [oracle.chacha_assignment(keys3[i], nonces3[i], ctrs3[i], ins3[i])[0][1:1153] for i in range(24)]
    assert ver.verify_batch(p3, pubs).all()
    ver.close(); ctx.close()
    monkeypatch.setenv("G16_BITQ_TEST_EXC", "1")
    ctx = G.Groth16Context(pk_bytes, r1cs_bytes)
    q1, _ = ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)
    q2, _ = ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)             # exception -> proved again without the tables, learning re-opened
    assert ctx._L.g16_last_counters_ex(ctx._h, u64p) == 0 and (int(e[11]) & 0xFF, (int(e[11]) >> 8) & 1) == (0, 0)
    q3, _ = ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)             # tables rebuilt from the narrowed classification, used again
    assert ctx._L.g16_last_counters_ex(ctx._h, u64p) == 0 and (int(e[11]) & 0xFF, (int(e[11]) >> 8) & 1) == (1, 1)
    assert q1 == p1 and q2 == p1 and q3 == p1
    ctx.close()


def test_two_lane_schedule_matches_the_single_stream(G, pk_bytes, r1cs_bytes, monkeypatch):
    """Circuits without a commitment default to the two-lane schedule for batches above one sub-batch (sub-batch k runs its chain
    solve -> transforms -> Z query on lane k % 2). Forced here with sub-batches of 16: 40 requests = three sub-batches over two
    lanes, compared byte for byte with the single main stream (g16_set_schedule), with the combination tables learned in
    between (the classification is also taken in the two-lane path)."""
    monkeypatch.setenv("G16_SUBBATCH", "16")
    monkeypatch.setenv("G16_BITQ_MIN_ROWS", "40")
    monkeypatch.setenv("G16_BITQ_MIN_BATCH", "8")
    keys, nonces, ctrs, ins, rs = batch_inputs(40, seed=b"g16-b200-lanes")
    ctx = G.Groth16Context(pk_bytes, r1cs_bytes)
    p1, c1 = ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)
    cn = ctx.counters()
    assert cn["pipelined"] and cn["sub_batch"] == 16 and cn["bitq_state"] == 0
    p2, _ = ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)             # two lanes + combination tables
    cn = ctx.counters()
    assert cn["pipelined"] and cn["bitq_live"]
    ctx.set_schedule(False, 16)
    p3, c3 = ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)
    assert not ctx.counters()["pipelined"]
    assert p1 == p2 == p3 and c1 == c3
    ctx.close()
    # the second lane's scratch does not fit (simulated): the batch is proved on the single stream, which stays the schedule
    monkeypatch.setenv("G16_TEST_LANE_OOM", "1")
    ctx = G.Groth16Context(pk_bytes, r1cs_bytes)
    p4, _ = ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)
    assert p4 == p1 and not ctx.counters()["pipelined"]
    p5, _ = ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)
    assert p5 == p1 and not ctx.counters()["pipelined"]
    ctx.close()


# ---------------------------------------------------------------------------------------------- product-side Setup (8f rank 1)
def test_setup_reproduces_oracle_keys_byte_for_byte(G, oracle):
    """g16_setup (QAP evaluation on the host, every key element a fixed-base product on the GPU) against the oracle's Setup
    restatement with the same trapdoor: the AES-128 proving and verifying key are byte-identical (13.8 MB, 143 k + 50 k
    points, the BSB22 commitment key included) — keygen.go:359-396 generateAES128."""
    from oracle import setup as S
    pk_ref, vk_ref, r1 = aes_keys(128)                       # oracle Setup, trapdoor from the seed "g16-b200-aes128"
    pk, vk = G.Setup(r1, S.toxic_from_seed(b"g16-b200-aes128"))
    assert vk == vk_ref
    assert len(pk) == len(pk_ref) and pk == pk_ref


def test_setup_keys_prove_and_verify(G, oracle, oracle_vk, r1cs_bytes, kat):
    """A fresh ChaCha key pair from g16_setup (random trapdoor from the OS CSPRNG inside the library): same structure as the
    reference's shipped pk.chacha20 / vk.chacha20, a GPU proof under the new pk verifies under the new vk (GPU verifier and
    the oracle's pairing check), and does NOT verify under the reference's shipped vk (different trapdoor)."""
    pk, vk = G.Setup(r1cs_bytes)
    assert len(pk) == len(open("tests/golden/pk.chacha20", "rb").read()) and len(vk) == len(open("tests/golden/vk.chacha20", "rb").read())
    pk2, vk2 = G.Setup(r1cs_bytes)
    assert pk2 != pk and vk2 != vk                           # fresh toxic waste every time
    ctx = G.Groth16Context(pk, r1cs_bytes)
    proofs, cts = ctx.prove_chacha_batch([kat["key"]], [kat["nonce"]], [kat["counter"]], [kat["input"]])
    assert cts[0] == kat["ct"] and proofs[0] != kat["proof"]
    inputs, _ = oracle.chacha_assignment(kat["key"], kat["nonce"], kat["counter"], kat["input"])
    pub = inputs[1:1153]
    ver = G.Groth16Verifier(vk)
    assert ver.verify(proofs[0], pub)
    assert oracle.VerifyingKeyOracle(vk).verify(proofs[0], pub)
    assert not oracle_vk.verify(proofs[0], pub)
    ctx.close(); ver.close()
    with pytest.raises(G.ProverError):                       # a trapdoor value that is not reduced modulo r
        G.Setup(r1cs_bytes, [oracle.R_MOD, 2, 3, 4, 5, 6])
    with pytest.raises(G.ProverError):
        G.Setup(r1cs_bytes[:-9])
