"""GPU parity tests for the AES-V2 circuits (BASELINE configs 2 and 3), through the C-ABI.

Keys: the reference ships no pk.aes128 / pk.aes256 (.MISSING_LARGE_BLOBS), so both sides use the keys of the oracle's
Setup restatement on the reference's r1cs.aes128 / r1cs.aes256 (tests/conftest.py::aes_keys). Bit-exact means bit-exact
with the in-repo CPU oracle for fixed (r, s, mask); AES parity with gnark's bytes is UNPINNED."""
import struct

import numpy as np
import pytest

from conftest import AES_KAT, AES_RSM, aes_keys

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gnark_symmetric_crypto_b200 as G
    return G


@pytest.fixture(scope="module")
def ctx128(G):
    pk, vk, r1 = aes_keys(128)
    c = G.Groth16Context(pk, r1, device=0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx256(G):
    pk, vk, r1 = aes_keys(256)
    c = G.Groth16Context(pk, r1, device=0)
    yield c
    c.close()


def rsm_bytes(r, s, m):
    return r.to_bytes(32, "big") + s.to_bytes(32, "big") + m.to_bytes(32, "big")


def signals_of(ct, nonce, counter, pt):   # verifiers.go:110-119 (counter big-endian for AES, core_test.go:205,249)
    return ct + nonce + struct.pack(">I", counter) + pt


def test_aes_witness_and_commitment_hash(G, oracle):
    """provers.go:184-210 on the device (AES-CTR keystream + byte-valued witness) and the BSB22 commitment hash
    (hash_to_field over SHA-256) against the oracle."""
    from oracle import setup as S
    rng = np.random.default_rng(3)
    for klen in (16, 32):
        n = 37
        keys = [rng.bytes(klen) for _ in range(n)]; nonces = [rng.bytes(12) for _ in range(n)]
        ctrs = [int(rng.integers(0, 1 << 32)) for _ in range(n)]; ins = [rng.bytes(64) for _ in range(n)]
        ctrs[0], ctrs[1], ctrs[2] = 0, 0xFFFFFFFC, 0xFFFFFFFE       # the last one carries into the nonce (cipher.NewCTR)
        nonces[2] = bytes([0xFF]) * 12
        cts, wit = G.aes_witness(keys, nonces, ctrs, ins)
        for j in range(n):
            a, ct = S.aes_assignment(keys[j], nonces[j], ctrs[j], ins[j])
            assert cts[j] == ct, (klen, j)
            assert np.array_equal(wit[j], oracle.to_mont(1, oracle.ints_to_limbs(a))), (klen, j)
    key = bytes.fromhex("7E24067817FAE0D743D6CE1F32539163")           # RFC 3686 test vector #2
    nonce = bytes.fromhex("006CB6DBC0543B59DA48D90B")
    cts, _ = G.aes_witness([key], [nonce], [1], [bytes(range(32)) + bytes(32)], with_witness=False)
    assert cts[0][:32].hex().upper() == "5104A106168A72D9790D41EE8EDAD388EB2E1EFC46DA57C8FCE630DF9141BE28"
    pts = oracle.g1_fixed_base(oracle.rand_field(rng, 1, 65)); pts[7] = 0
    ref = oracle.to_mont(1, oracle.ints_to_limbs([S.hash_to_fr(S.g1_uncompressed(p), b"bsb22-commitment") for p in pts]))
    assert np.array_equal(G.bsb22_challenge(pts), ref)


def check_config(G, ctx, orc, oracle, bits):
    from oracle import setup as S
    k = AES_KAT[bits]
    r, s, m = AES_RSM
    ref_proof, ref_ct, ref = orc.prove(k["key"], k["nonce"], k["counter"], k["input"], r, s, m, detail=True)
    assert ctx.proof_bytes == 196 and ctx.nb_commitments == 1 and ctx.supported
    # solver: every wire and every constraint evaluation (incl. the commitment challenge wire, lookups, divisions)
    inputs, ct = S.aes_assignment(k["key"], k["nonce"], k["counter"], k["input"])
    wit = oracle.to_mont(1, oracle.ints_to_limbs(inputs[1:]))
    W, A, B, Cc = ctx.solve(wit, 1, masks=[m])
    nc = A.shape[1]
    assert np.array_equal(W[0], ref["W"])
    assert np.array_equal(A[0], ref["A"][:nc]) and np.array_equal(B[0], ref["B"][:nc]) and np.array_equal(Cc[0], ref["C"][:nc])
    # gnark-shaped entry with every intermediate: H coefficient for coefficient, the five MSM points (G2 B included)
    proof, det = ctx.prove_witness(wit, rsm_bytes(r, s, m), detail=True)
    assert np.array_equal(det["h"], ref["h"][oracle.bitrev_perm(ctx.n)])
    for name in ("msmA", "msmB1", "msmK", "msmZ", "msmB2"):
        assert np.array_equal(det[name], ref[name]), name
    assert proof == ref_proof
    assert proof[32:96] == oracle.g2_compress(ref["Bs"].reshape(1, 16))
    # library-shaped entry
    proofs, cts = ctx.prove_aes_batch([k["key"]], [k["nonce"]], [k["counter"]], [k["input"]], [rsm_bytes(r, s, m)])
    assert proofs[0] == ref_proof and cts[0] == ref_ct == ct
    pub = S.aes_public_from_signals(signals_of(ct, k["nonce"], k["counter"], k["input"]))
    assert orc.verify(proofs[0], pub)
    # fresh randomness: different bytes, still accepted
    p2, _ = ctx.prove_aes_batch([k["key"]], [k["nonce"]], [k["counter"]], [k["input"]], None)
    assert p2[0] != ref_proof and orc.verify(p2[0], pub)


def test_aes128_config2(G, ctx128, aes128_oracle, oracle):
    """BASELINE config 2: core_test.go:265 input, fixed r, s, mask: bit-exact proof vs the CPU oracle."""
    check_config(G, ctx128, aes128_oracle, oracle, 128)


def test_aes256_config3(G, ctx256, aes256_oracle, oracle):
    """BASELINE config 3: core_test.go:275 input; H polynomial (2^17 coefficients) and the G2 MSM compared
    coefficient for coefficient, then the whole proof."""
    check_config(G, ctx256, aes256_oracle, oracle, 256)


@pytest.mark.parametrize("bits", [128, 256])
def test_aes_batch_matches_oracle(G, ctx128, ctx256, aes128_oracle, aes256_oracle, oracle, bits):
    from oracle import setup as S
    ctx, orc = (ctx128, aes128_oracle) if bits == 128 else (ctx256, aes256_oracle)
    rng = np.random.default_rng(bits)
    n = 24
    keys = [rng.bytes(bits // 8) for _ in range(n)]; nonces = [rng.bytes(12) for _ in range(n)]
    ctrs = [int(rng.integers(0, (1 << 32) - 4)) for _ in range(n)]; ins = [rng.bytes(64) for _ in range(n)]
    ctrs[0] = 0xFFFFFFFB                                        # largest counter the circuit accepts (4 blocks)
    rsm = [tuple(int.from_bytes(rng.bytes(32), "big") % oracle.R_MOD for _ in range(3)) for _ in range(n)]
    proofs, cts = ctx.prove_aes_batch(keys, nonces, ctrs, ins, [rsm_bytes(*t) for t in rsm])
    assert len(set(proofs)) == n
    for j in range(n):
        assert cts[j] == S.aes_ctr(keys[j], nonces[j], ctrs[j], ins[j])
        assert proofs[j][128:132] == b"\x00\x00\x00\x01"
    for j in (0, 1, 7, 23):
        ref, _ = orc.prove(keys[j], nonces[j], ctrs[j], ins[j], *rsm[j])
        assert proofs[j] == ref, j
        assert orc.verify(proofs[j], S.aes_public_from_signals(signals_of(cts[j], nonces[j], ctrs[j], ins[j])))
    # batch position must not matter
    p7, _ = ctx.prove_aes_batch(keys[7:8], nonces[7:8], ctrs[7:8], ins[7:8], [rsm_bytes(*rsm[7])])
    assert p7[0] == proofs[7]
    # a counter within 3 of 2^32 fails the circuit's AssertIsLessOrEqual(counter, MaxUint32): groth16.Prove errors
    with pytest.raises(G.ProverError) as e:
        ctx.prove_aes_batch(keys[:2], nonces[:2], [5, 0xFFFFFFFD], ins[:2], None)
    assert e.value.rc == 4   # G16_ERR_UNSAT
    with pytest.raises(G.ProverError):
        ctx.prove_aes_batch([bytes(32)] if bits == 128 else [bytes(16)], nonces[:1], ctrs[:1], ins[:1], None)   # other circuit's key


def test_libprove_abi_aes(G, aes128_oracle, aes256_oracle, oracle):
    """core_test.go:174-260 TestFullAES256 / TestFullAES128 through InitAlgorithm + Prove(JSON)."""
    from oracle import setup as S
    rng = np.random.default_rng(99)
    for alg, name, bits, orc in ((G.AES_128, "aes-128-ctr", 128, aes128_oracle), (G.AES_256, "aes-256-ctr", 256, aes256_oracle)):
        pk, vk, r1 = aes_keys(bits)
        assert G.InitAlgorithm(alg, pk, r1) is True
        key, nonce, pt, counter = rng.bytes(bits // 8), rng.bytes(12), rng.bytes(64), int(rng.integers(0, 1 << 31))
        out = G.OutputParams.from_json(G.Prove(G.InputParams(name, key, nonce, counter, pt).to_json()))
        assert len(out.proof_json) == 196
        assert out.public_signals == S.aes_ctr(key, nonce, counter, pt)
        assert orc.verify(out.proof_json, S.aes_public_from_signals(signals_of(out.public_signals, nonce, counter, pt)))
        with pytest.raises(RuntimeError, match="nonce length must be 12"):
            G.Prove(G.InputParams(name, key, nonce[:11], counter, pt).to_json())
        with pytest.raises(RuntimeError, match="key length must be 16 or 32"):
            G.Prove(G.InputParams(name, key + b"x", nonce, counter, pt).to_json())
