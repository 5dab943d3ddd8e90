"""CPU: kernel logic and host orchestration of the product, exercised through the TEST-ONLY host-emulation build
(tests/emu: the same .cu sources compiled as C++ with -DG16_EMU) and compared with the oracle. The PTX carry chains
themselves only run on the GPU (tests/test_gpu.py); their C twins run here."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN
from gnark_symmetric_crypto_b200 import _lib


def p64(a):
    return a.ctypes.data_as(_lib.u64p)


def p8(a):
    return a.ctypes.data_as(_lib.u8p)


def ok(L, rc):
    assert rc == 0, L.g16_last_error().decode()


@pytest.mark.parametrize("fld", [0, 1])
def test_field_ops(emu, oracle, fld):
    rng = np.random.default_rng(100 + fld)
    n = 300
    mod = oracle.P_MOD if fld == 0 else oracle.R_MOD
    a = oracle.to_mont(fld, oracle.rand_field(rng, fld, n)); b = oracle.to_mont(fld, oracle.rand_field(rng, fld, n))
    a[0] = 0; b[1] = 0
    a[2] = b[2] = oracle.to_mont(fld, oracle.ints_to_limbs([mod - 1]))[0]
    a[3] = oracle.to_mont(fld, oracle.ints_to_limbs([1]))[0]
    for op, code in (("add", 0), ("sub", 1), ("mul", 2), ("inv", 3), ("sqr", 4), ("neg", 5)):
        out = np.empty_like(a)
        ok(emu, emu.g16_field_op(fld, code, p64(a), p64(b) if code < 3 else None, p64(out), n))
        assert np.array_equal(out, oracle.f_op(fld, op, a, b if code < 3 else None)), op
    can = oracle.rand_field(rng, fld, n)
    out = np.empty_like(can)
    ok(emu, emu.g16_field_op(fld, 6, p64(can), None, p64(out), n))
    assert np.array_equal(out, oracle.to_mont(fld, can))
    back = np.empty_like(can)
    ok(emu, emu.g16_field_op(fld, 7, p64(out), None, p64(back), n))
    assert np.array_equal(back, can)


def test_group_ops_and_special_cases(emu, oracle):
    rng = np.random.default_rng(7)
    P = oracle.g1_fixed_base(oracle.rand_field(rng, 1, 12)); Q = oracle.g1_fixed_base(oracle.rand_field(rng, 1, 12))
    Q[0] = P[0]                                                             # P + P
    Q[1] = P[1]; Q[1][4:8] = oracle.f_op(0, "neg", P[1][4:8].reshape(1, 4))[0]   # P + (-P)
    Q[2] = 0; P[3] = 0                                                      # infinity on either side
    out = np.empty_like(P)
    ref = np.array([oracle.g1_add(P[i], Q[i]) for i in range(12)])
    for op in (0, 3):
        ok(emu, emu.g16_group_op(1, op, p64(P), p64(Q), p64(out), 12))
        assert np.array_equal(out, ref), op
    assert not ref[1].any()
    ok(emu, emu.g16_group_op(1, 4, p64(P), p64(Q), p64(out), 12))          # the four-warp addition (team.cuh)
    assert np.array_equal(out, ref)
    # 4a + 2b with both operands projective: b = 2a (equal points, different representations), b = -2a (infinity), infinities
    Q5 = Q.copy(); P5 = P.copy()
    Q5[0] = oracle.g1_add(P[0], P[0])
    Q5[1] = oracle.g1_add(P[1], P[1]); Q5[1][4:8] = oracle.f_op(0, "neg", Q5[1][4:8].reshape(1, 4))[0]
    ref5 = np.array([oracle.g1_add(oracle.g1_mul(P5[i], 4), oracle.g1_mul(Q5[i], 2)) for i in range(12)])
    ok(emu, emu.g16_group_op(1, 5, p64(P5), p64(Q5), p64(out), 12))
    assert np.array_equal(out, ref5) and not ref5[1].any()
    sc = oracle.rand_field(rng, 1, 12)
    ok(emu, emu.g16_group_op(1, 1, p64(P), p64(sc), p64(out), 12))
    assert np.array_equal(out, np.array([oracle.g1_mul(P[i], oracle.limbs_to_ints(sc[i:i + 1])[0]) for i in range(12)]))
    P2 = oracle.g2_fixed_base(oracle.rand_field(rng, 1, 6)); Q2 = oracle.g2_fixed_base(oracle.rand_field(rng, 1, 6)); Q2[0] = P2[0]
    out2 = np.empty_like(P2)
    ref2 = np.array([oracle.g2_add(P2[i], Q2[i]) for i in range(6)])
    for op in (0, 3):
        ok(emu, emu.g16_group_op(2, op, p64(P2), p64(Q2), p64(out2), 6))
        assert np.array_equal(out2, ref2), op
    c1 = np.frombuffer(oracle.g1_compress(P), dtype=np.uint8).copy(); d1 = np.empty_like(P)
    ok(emu, emu.g16_decompress(1, p8(c1), p64(d1), 12)); assert np.array_equal(d1, P)
    c2 = np.frombuffer(oracle.g2_compress(P2), dtype=np.uint8).copy(); d2 = np.empty_like(P2)
    ok(emu, emu.g16_decompress(2, p8(c2), p64(d2), 6)); assert np.array_equal(d2, P2)
    bad = c1.copy(); bad[0] &= 0x3F   # "uncompressed" flag in a 32-byte slot
    assert emu.g16_decompress(1, p8(bad), p64(d1), 12) == 2   # G16_ERR_PARSE


@pytest.mark.parametrize("n,c", [(1, 0), (7, 3), (300, 0), (300, 7), (700, 10)])
def test_msm_g1(emu, oracle, n, c):
    rng = np.random.default_rng(n * 31 + c)
    pts = oracle.g1_fixed_base(oracle.rand_field(rng, 1, n)); sc = oracle.rand_field(rng, 1, n)
    if n >= 300:   # zeros, +-1 runs (one huge bucket), duplicate points, infinity, empty buckets
        sc[0] = 0; sc[1] = oracle.ints_to_limbs([1])[0]; sc[2] = oracle.ints_to_limbs([oracle.R_MOD - 1])[0]
        pts[5] = pts[4]; sc[5] = sc[4]; pts[7] = 0
        sc[10:130] = oracle.ints_to_limbs([1] * 120); sc[130:200] = oracle.ints_to_limbs([oracle.R_MOD - 1] * 70)
    out = np.empty(8, dtype=np.uint64)
    ok(emu, emu.g16_msm(1, p64(pts), p64(sc), 0, n, c, p64(out), None))
    ref = oracle.g1_msm(pts, sc)
    assert np.array_equal(out, ref)
    scm = oracle.to_mont(1, sc)
    ok(emu, emu.g16_msm(1, p64(pts), p64(scm), 1, n, c, p64(out), None))
    assert np.array_equal(out, ref)


def test_msm_fixed_base_plan(emu, oracle):
    """Fixed-base mode (tables 2^(cw) P_i, one bucket set): the shape the prover's pk queries use. Run by
    tests/test_emu_variants.py once more with the per-row shared-memory sort and the batch-affine levels forced on."""
    rng = np.random.default_rng(31)
    n = 400
    pts = oracle.g1_fixed_base(oracle.rand_field(rng, 1, n)); pts[7] = 0; pts[9] = pts[8]
    for window in (6, 9):
        h = C.c_void_p()
        ok(emu, emu.g16_msm_plan_create(1, p64(pts), n, window, 0, C.byref(h)))
        ok(emu, emu.g16_msm_plan_precompute(h, window))
        for rep in range(2):
            sc = oracle.rand_field(rng, 1, n); sc[3] = 0; sc[8] = sc[9]
            sc[20:90] = oracle.ints_to_limbs([1] * 70); sc[90:120] = oracle.ints_to_limbs([oracle.R_MOD - 1] * 30)
            ok(emu, emu.g16_msm_plan_set_scalars(h, p64(sc), 0))
            out = np.empty(8, dtype=np.uint64)
            ok(emu, emu.g16_msm_plan_run(h, p64(out), None))
            assert np.array_equal(out, oracle.g1_msm(pts, sc)), (window, rep)
        emu.g16_msm_plan_free(h)


def test_msm_edge_cases(emu, oracle):
    rng = np.random.default_rng(2)
    pts = oracle.g1_fixed_base(oracle.rand_field(rng, 1, 50))
    out = np.empty(8, dtype=np.uint64)
    zeros = np.zeros((50, 4), dtype=np.uint64)
    ok(emu, emu.g16_msm(1, p64(pts), p64(zeros), 0, 50, 0, p64(out), None))
    assert not out.any()                                   # all-zero scalars -> infinity
    sc = oracle.rand_field(rng, 1, 50)
    same = np.repeat(pts[:1], 50, axis=0).copy()           # 50 copies of one point: every bucket add is a doubling
    ok(emu, emu.g16_msm(1, p64(same), p64(sc), 0, 50, 4, p64(out), None))
    assert np.array_equal(out, oracle.g1_msm(same, sc))
    assert emu.g16_msm(1, p64(pts), p64(sc), 0, 0, 0, p64(out), None) == 1   # n = 0 -> G16_ERR_ARG
    p2 = oracle.g2_fixed_base(oracle.rand_field(rng, 1, 60)); s2 = oracle.rand_field(rng, 1, 60)
    out2 = np.empty(16, dtype=np.uint64)
    ok(emu, emu.g16_msm(2, p64(p2), p64(s2), 0, 60, 5, p64(out2), None))
    assert np.array_equal(out2, oracle.g2_msm(p2, s2))


def _bitq_case(L, oracle, group, n, rows, seed):
    """Combination-table sums (csrc/k_bitq.cu) against the oracle's MSM on the same scalars: the first ~60 % of the wires hold
    0 / 1 (binary groups of 8 points), the others 0 / 1 / -1 (ternary groups of 5)."""
    rng = np.random.default_rng(seed)
    pts = (oracle.g1_fixed_base if group == 1 else oracle.g2_fixed_base)(oracle.rand_field(rng, 1, n))
    nb = n if n < 3 else (3 * n) // 5
    if n > 20:
        pts[3] = pts[2]      # equal points inside one group: a subset sum that is a doubling
        pts[9] = 0           # the point at infinity as a table operand
        pts[n - 1] = pts[n - 2]   # ... and inside a ternary group: P - P = infinity, P + P a doubling
    vals = rng.integers(0, 2, size=(n, rows))
    vals[nb:] = rng.integers(-1, 2, size=(n - nb, rows))
    vals[:, 0] = 0           # a witness without any non-zero wire: infinity
    if rows > 1:
        vals[:, 1] = 1       # every wire 1
    if rows > 2:
        vals[nb:, 2] = -1    # every trit -1
    one = oracle.to_mont(1, oracle.ints_to_limbs([1]))[0]
    mone = oracle.to_mont(1, oracle.ints_to_limbs([oracle.R_MOD - 1]))[0]
    wires = np.zeros((n, rows, 4), dtype=np.uint64)
    wires[vals == 1] = one
    wires[vals == -1] = mone
    w = 8 if group == 1 else 16
    out = np.zeros((rows, w), dtype=np.uint64)
    exc = C.c_uint32(7)
    ok(L, L.g16_bitq_sum(group, p64(pts), n, nb, p64(wires), rows, p64(out), C.byref(exc)))
    assert exc.value == 0
    msm = oracle.g1_msm if group == 1 else oracle.g2_msm
    minus1 = oracle.ints_to_limbs([oracle.R_MOD - 1])[0]
    for r in range(rows):
        sc = np.zeros((n, 4), dtype=np.uint64); sc[:, 0] = (vals[:, r] == 1)
        sc[vals[:, r] == -1] = minus1
        assert np.array_equal(out[r], msm(pts, sc)), (group, r)
    assert not out[0].any()
    # a wire that holds something its group does not allow is reported, never silently used
    bad = wires.copy(); bad[n // 2, rows - 1] = oracle.to_mont(1, oracle.ints_to_limbs([2]))[0]
    ok(L, L.g16_bitq_sum(group, p64(pts), n, nb, p64(bad), rows, p64(out), C.byref(exc)))
    assert exc.value == 1
    if nb > 0:
        bad = wires.copy(); bad[0, rows - 1] = mone      # -1 on a wire classified as a bit
        ok(L, L.g16_bitq_sum(group, p64(pts), n, nb, p64(bad), rows, p64(out), C.byref(exc)))
        assert exc.value == 1


@pytest.mark.parametrize("group,n,rows", [(1, 1, 1), (1, 8, 3), (1, 77, 5), (1, 300, 40), (2, 50, 4)])
def test_bit_wire_combination_tables(emu, oracle, group, n, rows):
    _bitq_case(emu, oracle, group, n, rows, 1000 * group + n)


@pytest.mark.parametrize("n", [2, 8, 256, 4096])
def test_ntt(emu, oracle, n):
    rng = np.random.default_rng(n)
    x = oracle.to_mont(1, oracle.rand_field(rng, 1, n))
    y = x.copy(); ok(emu, emu.g16_ntt(p64(y), n, 0, 0, None)); assert np.array_equal(y, oracle.ntt(x))
    z = y.copy(); ok(emu, emu.g16_ntt(p64(z), n, 1, 0, None)); assert np.array_equal(z, x)
    # coset transforms invert each other
    y = x.copy(); ok(emu, emu.g16_ntt(p64(y), n, 0, 1, None))
    z = y.copy(); ok(emu, emu.g16_ntt(p64(z), n, 1, 1, None)); assert np.array_equal(z, x)
    assert emu.g16_ntt(p64(x), 3, 0, 0, None) == 1   # not a power of two


def test_ntt_multi_pass_plan(emu, oracle):
    # 2^13 needs a strided pass on top of the 2^11 tile pass
    rng = np.random.default_rng(13)
    n = 1 << 13
    x = oracle.to_mont(1, oracle.rand_field(rng, 1, n))
    y = x.copy(); ok(emu, emu.g16_ntt(p64(y), n, 0, 0, None)); assert np.array_equal(y, oracle.ntt(x))
    z = y.copy(); ok(emu, emu.g16_ntt(p64(z), n, 1, 0, None)); assert np.array_equal(z, x)


@pytest.fixture(scope="module")
def emu_ctx(emu, pk_bytes, r1cs_bytes):
    os.environ["G16_LAZY_TABLES"] = "1"   # the fixed-base tables take minutes to build under emulation
    h = C.c_void_p()
    try:
        ok(emu, emu.g16_init(pk_bytes, len(pk_bytes), r1cs_bytes, len(r1cs_bytes), 0, C.byref(h)))
    finally:
        os.environ.pop("G16_LAZY_TABLES", None)
    yield h
    emu.g16_free(h)


def test_host_parsers_and_solver_match_oracle(emu, emu_ctx, oracle, oracle_prover, kat):
    info = np.zeros(16, dtype=np.uint64)
    ok(emu, emu.g16_info(emu_ctx, p64(info)))
    r = oracle_prover.cs.r
    assert [int(x) for x in info[:13]] == [32768, 22001, 12529, 32767, 22128, 12529, r.n_wires, r.n_public, r.n_secret,
                                          r.n_constraints, r.n_instr, len(r.levels), 0]
    rng = np.random.default_rng(9)
    reqs = [(kat["key"], kat["nonce"], kat["counter"], kat["input"]), (rng.bytes(32), rng.bytes(12), 0xFFFFFFFF, rng.bytes(64))]
    wit = []
    refs = []
    for k, n, c, i in reqs:
        inputs, _ = oracle.chacha_assignment(k, n, c, i)
        wit.append(oracle.to_mont(1, oracle.ints_to_limbs(inputs[1:])))
        refs.append(oracle_prover.cs.solve(inputs))
    w = np.stack(wit)
    nw, nc = r.n_wires, r.n_constraints
    W = np.zeros((2, nw, 4), dtype=np.uint64); A = np.zeros((2, nc, 4), dtype=np.uint64); B = np.zeros_like(A); Cc = np.zeros_like(A)
    ok(emu, emu.g16_solve(emu_ctx, p64(w), w.shape[1], 2, p64(W), p64(A), p64(B), p64(Cc)))
    for j in range(2):
        for got, ref in zip((W[j], A[j], B[j], Cc[j]), refs[j]):
            assert np.array_equal(got, ref)
    # a witness with one flipped ciphertext bit is rejected with G16_ERR_UNSAT
    bad = w.copy()
    bad[0, 640] = oracle.to_mont(1, oracle.ints_to_limbs([1 - oracle.limbs_to_ints(oracle.from_mont(1, w[0, 640:641]))[0]]))[0]
    assert emu.g16_solve(emu_ctx, p64(bad), bad.shape[1], 2, None, None, None, None) == 4
    assert emu.g16_solve(emu_ctx, p64(w), 5, 1, None, None, None, None) == 1   # wrong witness length


def test_compute_h_matches_oracle(emu, emu_ctx, oracle, oracle_prover, kat):
    inputs, _ = oracle.chacha_assignment(kat["key"], kat["nonce"], kat["counter"], kat["input"])
    W, A, B, Cc = oracle_prover.cs.solve(inputs)
    n = oracle_prover.pk.n
    h = np.zeros((n, 4), dtype=np.uint64)
    ok(emu, emu.g16_compute_h(emu_ctx, p64(A), p64(B), p64(Cc), p64(h)))
    ref = oracle.compute_h(A, B, Cc, n)
    assert np.array_equal(h, ref[oracle.bitrev_perm(n)])   # gnark's array order = bit-reversed coefficients
    assert not h[oracle.bitrev_perm(n)[n - 1]].any()       # h_{n-1} = 0


def test_malformed_keys_are_rejected(emu, pk_bytes, r1cs_bytes):
    h = C.c_void_p()
    assert emu.g16_init(pk_bytes[:1000], 1000, r1cs_bytes, len(r1cs_bytes), 0, C.byref(h)) == 2
    assert b"truncated" in emu.g16_last_error()
    assert emu.g16_init(pk_bytes, len(pk_bytes), r1cs_bytes[:-7], len(r1cs_bytes) - 7, 0, C.byref(h)) == 2
    assert emu.g16_init(None, 0, r1cs_bytes, len(r1cs_bytes), 0, C.byref(h)) == 1


def _call_json(fn, free, payload: bytes):
    import json
    buf = (C.c_uint8 * max(len(payload), 1)).from_buffer_copy(payload or b"\0")
    r = fn(_lib.GoSlice(C.cast(buf, C.c_void_p), len(payload), len(payload)))
    out = C.string_at(r.r0, r.r1)
    free(r.r0)
    return json.loads(out)


def test_libprove_json_layer(emu):
    """Error behaviour of the outer ABI (libprove.go:30-47, prove_impl.go:116-143) without any initialised cipher. A string
    panic (log.Panicf, fmt.Sprintf) comes back as a JSON string, panic(err) as the marshalled error VALUE: an object."""
    prove = lambda b: _call_json(emu.Prove, emu.Free, b)
    assert "could not find prover" in prove(b'{"cipher":"aes-256-ctr1","key":[1],"nonce":[1],"counter":1,"input":[1]}')
    assert "not initialized" in prove(b'{"cipher":"chacha20","key":"AAEC","nonce":[],"counter":1,"input":[]}')
    assert prove(b"{not json") == {"Offset": 2}                        # *json.SyntaxError: only Offset is exported
    assert prove(b"") == {"Offset": 0}
    err = prove(b'{"cipher":"chacha20","counter":[0,1]}')            # core_test.go:122 passes an array: *json.UnmarshalTypeError
    assert err["Field"] == "counter" and err["Struct"] == "InputParams" and err["Value"] == "array"
    # encoding/json hands unsigned fields to strconv.ParseUint: exponents, fractions, signs and hex are type errors
    for lit in (b"1e3", b"1.0", b"-1", b"4294967296"):
        assert prove(b'{"cipher":"chacha20","counter":' + lit + b'}')["Field"] == "counter", lit
    assert "Offset" in prove(b'{"cipher":"chacha20","counter":0x10}')   # not JSON at all
    assert prove(b'{"cipher":"chacha20","key":[256]}')["Field"] == "key"
    assert "Offset" in prove(b'{"cipher":"chacha20"} trailing')
    assert "Offset" in prove(b'{"cipher":"chacha20","counter":12')      # input ends inside a value: no read past the slice
    assert "Offset" in prove(b'{"cipher":"chacha20","counter":tru')
    # base64 as base64.StdEncoding decodes it: padding is mandatory and only at the end
    assert isinstance(prove(b'{"cipher":"chacha20","key":"AAE"}'), int)          # CorruptInputError marshals as its int64 value
    assert isinstance(prove(b'{"cipher":"chacha20","key":"AA=A"}'), int)
    assert "not initialized" in prove(b'{"cipher":"chacha20","key":"AAE=","nonce":null}')
    # field names match case-insensitively, \u escapes decode to UTF-8, unknown fields are skipped
    assert "not initialized" in prove(b'{"CIPHER":"chacha20","Key":[1],"extra":{"a":[1,2,{"b":null}]},"x":"\\u00e9"}')
    assert prove('{"cipher":"ch\\u0061cha20\\u00e9"}'.encode()) == "could not find prover forchacha20\u00e9"


def test_libprove_prove_batch_json_layer(emu):
    """ProveBatch: per-request results in order; without initialised ciphers every request reports what
    Prove would report. A payload that is not an array fails as one value, like one json.Unmarshal."""
    pb = lambda b: _call_json(emu.ProveBatch, emu.Free, b)
    assert pb(b"[]") == []
    out = pb(b'[{"cipher":"chacha20","key":[],"nonce":[],"counter":1,"input":[]},{"cipher":"nope"},{"cipher":"aes-128-ctr"}]')
    assert len(out) == 3 and "not initialized" in out[0] and "could not find prover" in out[1] and "not initialized" in out[2]
    assert pb(b'{"cipher":"chacha20"}') == {}
    assert pb(b'[{"cipher":"chacha20"},') == {"Offset": 23}
    assert "Field" in pb(b'[{"cipher":"chacha20","counter":"7"}]')


def test_aes_witness_and_bsb22_hash(emu, oracle):
    """provers.go:184-210 (AES-CTR keystream, byte-valued witness) and the BSB22 commitment hash (RFC 9380
    expand_message_xmd / SHA-256 -> Fr) as the device kernels compute them, against the oracle."""
    from oracle import setup as S
    rng = np.random.default_rng(11)
    for klen in (16, 32):
        n = 5
        keys = [rng.bytes(klen) for _ in range(n)]; nonces = [rng.bytes(12) for _ in range(n)]
        ctrs = [0, 0xFFFFFFFC, 0xFFFFFFFE, 7, 1 << 31]; ins = [rng.bytes(64) for _ in range(n)]
        nonces[2] = bytes([0xFF]) * 12   # cipher.NewCTR carries the counter overflow into the nonce bytes
        k = np.frombuffer(b"".join(keys), dtype=np.uint8).copy(); no = np.frombuffer(b"".join(nonces), dtype=np.uint8).copy()
        i = np.frombuffer(b"".join(ins), dtype=np.uint8).copy(); c = np.asarray(ctrs, dtype=np.uint32)
        cts = np.zeros(64 * n, dtype=np.uint8); wit = np.zeros((n, 142 + klen, 4), dtype=np.uint64)
        ok(emu, emu.g16_aes_witness(p8(k), klen, p8(no), c.ctypes.data_as(_lib.u32p), p8(i), n, p8(cts), p64(wit)))
        for j in range(n):
            a, ct = S.aes_assignment(keys[j], nonces[j], ctrs[j], ins[j])
            assert cts[64 * j:64 * j + 64].tobytes() == ct, (klen, j)
            assert np.array_equal(wit[j], oracle.to_mont(1, oracle.ints_to_limbs(a)))
    assert emu.g16_aes_witness(p8(k), 24, p8(no), c.ctypes.data_as(_lib.u32p), p8(i), n, p8(cts), None) == 1   # AES-192: bad argument
    pts = oracle.g1_fixed_base(oracle.rand_field(rng, 1, 9)); pts[4] = 0
    out = np.zeros((9, 4), dtype=np.uint64)
    ok(emu, emu.g16_bsb22_challenge(p64(pts), 9, p64(out)))
    ref = oracle.to_mont(1, oracle.ints_to_limbs([S.hash_to_fr(S.g1_uncompressed(p), b"bsb22-commitment") for p in pts]))
    assert np.array_equal(out, ref)


def _bilinear_pairs(oracle, rng, k):
    """k pairs (a_i G1, b_i G2) plus one closing pair (-(sum a_i b_i) G1, G2): their pairing product is 1."""
    R = oracle.R_MOD
    a = [int(x) for x in rng.integers(1, 1 << 62, k)]; b = [int(x) for x in rng.integers(1, 1 << 62, k)]
    g1 = oracle.g1_fixed_base(oracle.ints_to_limbs(a + [(-sum(x * y for x, y in zip(a, b))) % R]))
    g2 = oracle.g2_fixed_base(oracle.ints_to_limbs(b + [1]))
    return g1, g2


def test_pairing_check(emu, oracle):
    """Kernel logic of the pairing product check (Miller loop lines, block-cooperative Fp12 product, final exponentiation)
    against the oracle's pairing: a bilinear identity is accepted, a perturbed one rejected."""
    rng = np.random.default_rng(12)
    g1, g2 = _bilinear_pairs(oracle, rng, 1)
    assert oracle.pairing_check(g1, g2)
    bad1 = g1.copy(); bad1[0] = oracle.g1_fixed_base(oracle.ints_to_limbs([5]))[0]
    assert not oracle.pairing_check(bad1, g2)
    P = np.concatenate([g1, bad1]); Q = np.concatenate([g2, g2])
    out = np.full(2, 7, dtype=np.uint8)
    ok(emu, emu.g16_pairing_check(p64(P), p64(Q), 2, 2, p8(out)))
    assert out.tolist() == [1, 0]
    assert emu.g16_pairing_check(p64(P), p64(Q), 0, 2, p8(out)) == 1   # G16_ERR_ARG


def test_verifier_orchestration(emu, oracle, kat):
    """Host orchestration of the batched groth16.Verify (vk parsing, proof unpacking, public-input MSM, kSum, pairing inputs)
    under emulation: the committed known-answer proof is accepted under the reference's vk.chacha20. (Rejections are covered
    by test_pairing_check here and by tests/test_gpu_verify.py on the GPU.)"""
    vk = (GOLDEN / "vk.chacha20").read_bytes()
    h = C.c_void_p()
    ok(emu, emu.g16_verify_init(vk, len(vk), 0, C.byref(h)))
    info = np.zeros(4, dtype=np.uint64)
    ok(emu, emu.g16_verify_info(h, p64(info)))
    assert info.tolist() == [1152, 0, 164, 1153]
    inputs, _ = oracle.chacha_assignment(kat["key"], kat["nonce"], kat["counter"], kat["input"])
    pub = np.frombuffer(b"".join(int(x).to_bytes(32, "big") for x in inputs[1:1153]), dtype=np.uint8).copy()
    pr = np.frombuffer(kat["proof"], dtype=np.uint8).copy()
    out = np.zeros(1, dtype=np.uint8)
    ok(emu, emu.g16_verify_batch(h, 1, p8(pr), pub.ctypes.data_as(C.c_void_p), 1, p8(out), None))
    assert out[0] == 1
    assert emu.g16_verify_batch(h, 0, p8(pr), pub.ctypes.data_as(C.c_void_p), 1, p8(out), None) == 1   # G16_ERR_ARG
    # gnark-crypto decodes coordinates with SetBytesCanonical: x + p names the same point but is not a valid encoding.
    # The A0 half of Bs.x (bytes 64..96) carries no flag bits, so x + p < 2^255 always fits.
    x = int.from_bytes(kat["proof"][64:96], "big")
    forged = bytearray(kat["proof"])
    forged[64:96] = (x + oracle.P_MOD).to_bytes(32, "big")
    both = np.frombuffer(bytes(forged) + kat["proof"], dtype=np.uint8).copy()
    out2 = np.zeros(2, dtype=np.uint8)
    ok(emu, emu.g16_verify_batch(h, 2, p8(both), np.concatenate([pub, pub]).ctypes.data_as(C.c_void_p), 1, p8(out2), None))
    assert out2.tolist() == [0, 1]
    emu.g16_verify_free(h)
    bad = C.c_void_p()
    assert emu.g16_verify_init(vk[:-3], len(vk) - 3, 0, C.byref(bad)) == 2   # G16_ERR_PARSE


def test_verifier_two_rows_take_a_two_level_bucket_tree(emu, aes128_oracle):
    """Two AES proofs in one g16_verify_batch: the public-input MSM runs with rows = 2 over 4096 buckets whose contents are byte
    values (not just bucket 1 as for the ChaCha bit inputs), so the bucket-reduction tree has two levels and every child of the
    upper node counts. (Found on B200: the block-parallel tree kept its arity of 256 when it computed the size of a following
    thread-serial level; only batches of >= 2 rows over > 256 buckets were affected.)"""
    from conftest import aes_keys, AES_KAT, AES_RSM
    _, vk, _ = aes_keys(128)
    k = AES_KAT[128]
    proof, ct, _ = aes128_oracle.prove(k["key"], k["nonce"], k["counter"], k["input"], *AES_RSM, detail=True)
    vals = list(k["nonce"]) + [k["counter"]] + list(k["input"]) + list(ct)
    pub = np.frombuffer(b"".join(int(x).to_bytes(32, "big") for x in vals) * 2, dtype=np.uint8).copy()
    pr = np.frombuffer(bytes(proof) * 2, dtype=np.uint8).copy()
    h = C.c_void_p()
    ok(emu, emu.g16_verify_init(vk, len(vk), 0, C.byref(h)))
    out = np.zeros(2, dtype=np.uint8)
    ok(emu, emu.g16_verify_batch(h, 2, p8(pr), pub.ctypes.data_as(C.c_void_p), 1, p8(out), None))
    assert out.tolist() == [1, 1]
    emu.g16_verify_free(h)


def test_libverify_json_layer(emu):
    """libverify.go:14-17 / verify_impl.go:64-82 without a key: every failure is `false`, nothing throws across the ABI."""
    from gnark_symmetric_crypto_b200._lib import GoSlice

    def call(b):
        buf = (C.c_uint8 * max(len(b), 1)).from_buffer_copy(b or b"\0")
        return emu.Verify(GoSlice(C.cast(buf, C.c_void_p), len(b), len(b)))
    assert call(b'{"cipher":"chacha20","proof":[1,2],"publicSignals":[3]}') == 0     # verifier not initialised
    assert call(b'{"cipher":"nope","proof":[],"publicSignals":[]}') == 0
    assert call(b'{"cipher":"chacha20","proof":"####"}') == 0                           # illegal base64
    assert call(b'[1,2') == 0 and call(b'') == 0


def _twist_point_outside_g2(L):
    """A point of the twist E'(Fp2) that is (almost surely) not in the r-torsion: compressed x = small values until the
    decoder finds a square root. The twist's group order is r times a 254-bit cofactor."""
    for t in range(2, 200):
        raw = bytearray(64)
        raw[63] = t          # X.A0 = t (second half of the encoding), X.A1 = 1
        raw[31] = 1
        raw[0] |= 0x80
        out = np.zeros(16, dtype=np.uint64)
        if L.g16_decompress(2, p8(np.frombuffer(bytes(raw), dtype=np.uint8).copy()), p64(out), 1) == 0 and out.any():
            return out
    raise AssertionError("no twist point found")


def test_g2_subgroup_check(emu, oracle):
    """gnark's G2 decoder rejects points outside the r-torsion subgroup (g2.go IsInSubGroup); so does the verifier."""
    rng = np.random.default_rng(8)
    good = oracle.g2_fixed_base(oracle.rand_field(rng, 1, 3))
    bad = _twist_point_outside_g2(emu)
    assert oracle.g2_add(oracle.g2_mul(bad, oracle.R_MOD - 1), bad).any()   # [r]P != 0: really outside G2 (scalars are taken mod r)
    pts = np.concatenate([good, bad.reshape(1, 16), np.zeros((1, 16), dtype=np.uint64)])
    ok_out = np.full(5, 7, dtype=np.uint8)
    ok(emu, emu.g16_g2_subgroup_check(p64(pts), 5, p8(ok_out)))
    assert ok_out.tolist() == [1, 1, 1, 0, 1]               # infinity is in the subgroup
