"""CPU: the oracle against every golden vector the reference holds for this path — the KAT proof, the
shipped vk.chacha20 (pairing acceptance), the RFC 7539 vectors quoted in circuits/chachaV3/chacha_test.go, and internal
cross-checks (naive vs Pippenger MSM, NTT round trips, compression round trips)."""
import struct

import numpy as np
import pytest


def test_chacha_quarter_round_rfc7539(oracle):
    # circuits/chachaV3/chacha_test.go:22-30 (RFC 7539 §2.1.1)
    a, b, c, d = 0x11111111, 0x01020304, 0x9B8D6F43, 0x01234567
    rot = lambda x, n: ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF
    a = (a + b) & 0xFFFFFFFF; d = rot(d ^ a, 16); c = (c + d) & 0xFFFFFFFF; b = rot(b ^ c, 12)
    a = (a + b) & 0xFFFFFFFF; d = rot(d ^ a, 8); c = (c + d) & 0xFFFFFFFF; b = rot(b ^ c, 7)
    assert (a, b, c, d) == (0xEA2A92F4, 0xCB1CF8CE, 0x4581472E, 0x5881C4BB)


def test_chacha_block_rfc7539(oracle):
    # circuits/chachaV3/chacha_test.go:95-105 (RFC 7539 §2.3.2)
    key = bytes(range(32))
    nonce = bytes.fromhex("000000090000004a00000000")
    ks = oracle.chacha20_block(key, 1, nonce)
    want = [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
            0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]
    assert list(struct.unpack("<16I", ks)) == want


def test_chacha_matches_cryptography_package(oracle):
    # the reference's cipher oracle is x/crypto/chacha20 (provers.go:95-101); here: the `cryptography` package
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms
    rng = np.random.default_rng(5)
    for _ in range(5):
        key, nonce, pt = rng.bytes(32), rng.bytes(12), rng.bytes(64)
        ctr = int(rng.integers(0, 1 << 32))
        enc = Cipher(algorithms.ChaCha20(key, struct.pack("<I", ctr) + nonce), mode=None).encryptor()
        assert enc.update(pt) == oracle.chacha20_xor(key, nonce, ctr, pt)


def test_input_length_checks(oracle):
    # provers.go:81-89 log.Panicf
    with pytest.raises(ValueError, match="key length must be 32"):
        oracle.chacha_assignment(bytes(31), bytes(12), 0, bytes(64))
    with pytest.raises(ValueError, match="nonce length must be 12"):
        oracle.chacha_assignment(bytes(32), bytes(11), 0, bytes(64))
    with pytest.raises(ValueError, match="plaintext length must be 64"):
        oracle.chacha_assignment(bytes(32), bytes(12), 0, bytes(63))


def test_r1cs_header_counts(oracle_prover):
    r = oracle_prover.cs.r
    # SURVEY.md §8.0
    assert (r.n_public, r.n_secret, r.n_internal, r.n_constraints) == (1153, 256, 21872, 23617)
    assert r.n_instr == 23954 and len(r.levels) == 163 and len(r.calldata) == 452112 and len(r.coeffs) == 40
    kinds = np.array([r.bp_kind[b] for b in r.bp_id])
    assert (kinds == 0).sum() == 23617 and (kinds == 1).sum() == 337
    lv = np.concatenate(r.levels)
    assert sorted(lv.tolist()) == list(range(r.n_instr))


def test_pk_layout(oracle, oracle_prover, pk_bytes):
    from oracle import formats
    lay = formats.parse_pk_layout(pk_bytes)
    pk = oracle_prover.pk
    assert (lay.n, lay.counts["A"], lay.counts["B"], lay.counts["Z"], lay.counts["K"], lay.counts["B2"]) == \
        (32768, 22001, 12529, 32767, 22128, 12529) == (pk.n, pk.nA, pk.nB, pk.nZ, pk.nK, pk.nB2)
    assert (lay.nb_wires, lay.nb_inf_a, lay.nb_inf_b) == (23281, 1280, 10752)
    assert lay.hdr_fr[1] == oracle.root_of_unity(32768) and lay.hdr_fr[3] == 5
    assert oracle.lib().orc_g1_on_curve(pk.array("A").ctypes.data_as(oracle.u64p), pk.nA)
    assert oracle.lib().orc_g2_on_curve(pk.array("B2").ctypes.data_as(oracle.u64p), pk.nB2)
    # compression round trip on real key points
    raw = pk_bytes[lay.offs["K"]:lay.offs["K"] + 32 * 64]
    assert oracle.g1_compress(oracle.g1_decompress(raw)) == raw
    raw2 = pk_bytes[lay.offs["B2"]:lay.offs["B2"] + 64 * 16]
    assert oracle.g2_compress(oracle.g2_decompress(raw2)) == raw2


def test_pk_vk_consistent(oracle, oracle_prover, oracle_vk):
    # pk.chacha20 and vk.chacha20 carry the same alpha, beta, delta and e(beta1,G2) = e(G1,beta2)
    g1 = oracle_prover.pk.array("g1"); g2 = oracle_prover.pk.array("g2")
    v1 = oracle_vk.array("g1"); v2 = oracle_vk.array("g2")
    assert np.array_equal(g1, v1) and np.array_equal(g2[0], v2[0]) and np.array_equal(g2[1], v2[2])
    neg_g1 = oracle.g1_gen().copy()
    neg_g1[4:8] = oracle.f_op(oracle.FP, "neg", neg_g1[4:8].reshape(1, 4))[0]
    assert oracle.pairing_check(np.stack([g1[1], neg_g1]), np.stack([oracle.g2_gen(), g2[0]]))
    assert not oracle.pairing_check(np.stack([g1[0], neg_g1]), np.stack([oracle.g2_gen(), g2[0]]))


def test_kat_proof_and_verifier(oracle, oracle_prover, oracle_vk, kat):
    proof, ct = oracle_prover.prove(kat["key"], kat["nonce"], kat["counter"], kat["input"], kat["r"], kat["s"])
    assert ct == kat["ct"]
    assert proof == kat["proof"]
    signals = ct + kat["nonce"] + struct.pack("<I", kat["counter"]) + kat["input"]   # core_test.go:157-163
    pub = oracle.chacha_public_from_signals(signals)
    assert oracle_vk.verify(proof, pub)
    # a different public input, a tampered proof element and the README's stale proof must all be rejected
    bad = list(pub); bad[7] ^= 1
    assert not oracle_vk.verify(proof, bad)
    p2, _ = oracle_prover.prove(kat["key"], kat["nonce"], kat["counter"], kat["input"], kat["r"] + 1, kat["s"])
    assert p2 != proof and oracle_vk.verify(p2, pub)
    swapped = p2[:32] + proof[32:]
    assert not oracle_vk.verify(swapped, pub)


def test_solver_rejects_wrong_ciphertext(oracle, oracle_prover, kat):
    inputs, _ = oracle.chacha_assignment(kat["key"], kat["nonce"], kat["counter"], kat["input"])
    W, A, B, C = oracle_prover.cs.solve(inputs)
    ab = oracle.f_op(oracle.FR, "mul", A, B)
    assert np.array_equal(ab, C)   # every constraint satisfied
    vals = set(oracle.limbs_to_ints(oracle.from_mont(oracle.FR, W)))
    assert vals <= {0, 1, oracle.R_MOD - 1}   # SURVEY finding 7
    bad = list(inputs); bad[641] ^= 1   # flip one ciphertext bit
    with pytest.raises(ValueError):
        oracle_prover.cs.solve(bad)


def test_msm_naive_vs_pippenger(oracle):
    rng = np.random.default_rng(11)
    for n in (1, 2, 33, 400):
        pts = oracle.g1_fixed_base(oracle.rand_field(rng, oracle.FR, n)); sc = oracle.rand_field(rng, oracle.FR, n)
        if n > 2:
            sc[0] = 0; sc[1] = oracle.ints_to_limbs([oracle.R_MOD - 1])[0]; pts[2] = 0
        assert np.array_equal(oracle.g1_msm(pts, sc), oracle.g1_msm(pts, sc, naive=True))
    pts = oracle.g2_fixed_base(oracle.rand_field(rng, oracle.FR, 40)); sc = oracle.rand_field(rng, oracle.FR, 40)
    assert np.array_equal(oracle.g2_msm(pts, sc), oracle.g2_msm(pts, sc, naive=True))
    # linearity in the exponent: MSM(k_i G, s_i) = (sum k_i s_i) G
    ks = oracle.rand_field(rng, oracle.FR, 100); sc = oracle.rand_field(rng, oracle.FR, 100)
    tot = sum(a * b for a, b in zip(oracle.limbs_to_ints(ks), oracle.limbs_to_ints(sc))) % oracle.R_MOD
    assert np.array_equal(oracle.g1_msm(oracle.g1_fixed_base(ks), sc), oracle.g1_mul(oracle.g1_gen(), tot))


def test_ntt_and_compute_h(oracle):
    rng = np.random.default_rng(3)
    n = 64
    x = oracle.to_mont(oracle.FR, oracle.rand_field(rng, oracle.FR, n))
    y = oracle.ntt(x)
    assert np.array_equal(oracle.ntt(y, inverse=True), x)
    # direct evaluation of two outputs
    w = oracle.root_of_unity(n)
    xi = oracle.limbs_to_ints(oracle.from_mont(oracle.FR, x)); yi = oracle.limbs_to_ints(oracle.from_mont(oracle.FR, y))
    for k in (1, 37):
        assert yi[k] == sum(v * pow(w, k * j, oracle.R_MOD) for j, v in enumerate(xi)) % oracle.R_MOD
    # H of a satisfied system: (A.B - C) vanishes on the domain  =>  A*B - C == H * (X^n - 1) as polynomials
    ncons = 50
    a = oracle.to_mont(oracle.FR, oracle.rand_field(rng, oracle.FR, ncons)); b = oracle.to_mont(oracle.FR, oracle.rand_field(rng, oracle.FR, ncons))
    c = oracle.f_op(oracle.FR, "mul", a, b)
    h = oracle.limbs_to_ints(oracle.from_mont(oracle.FR, oracle.compute_h(a, b, c, n)))
    assert h[n - 1] == 0
    pad = lambda v: np.concatenate([v, np.zeros((n - ncons, 4), dtype=np.uint64)])
    coef = lambda v: oracle.limbs_to_ints(oracle.from_mont(oracle.FR, oracle.ntt(pad(v), inverse=True)))
    ca, cb, cc = coef(a), coef(b), coef(c)
    z = 123456789
    ev = lambda co: sum(v * pow(z, j, oracle.R_MOD) for j, v in enumerate(co)) % oracle.R_MOD
    assert (ev(ca) * ev(cb) - ev(cc)) % oracle.R_MOD == ev(h) * (pow(z, n, oracle.R_MOD) - 1) % oracle.R_MOD
