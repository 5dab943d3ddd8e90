"""CPU: the oracle's Groth16 Setup restatement and its AES (BSB22 commitment) prove / verify path (oracle/setup.py).

What pins it (the reference ships no pk.aes*, SURVEY.md §0.4 — byte-level AES parity with gnark stays UNPINNED):
  * expand_message_xmd against the RFC 9380 Appendix K.1 vectors;
  * Setup on the reference's r1cs.chacha20 reproduces the structure gnark recorded in the shipped pk.chacha20
    (query lengths, InfinityA / InfinityB masks, domain header) and yields keys under which the (pinned) ChaCha oracle
    prover's proofs verify;
  * Setup on the reference's r1cs.aes128/256 reproduces the structure of the shipped vk.aes128/256 (143 K points, one
    commitment without public-committed wires, one Pedersen key, 5008 bytes);
  * AES proofs for the reference's own benchmark inputs (libraries/core_test.go:265,275) verify; tampered ones do not;
  * the witness semantics satisfy every constraint of the reference's r1cs.aes* (the solver checks each R1C).
"""
import struct

import numpy as np
import pytest

from conftest import AES_KAT, AES_RSM, GOLDEN, aes_keys


def test_expand_message_xmd_rfc9380():
    from oracle import setup as S
    dst = b"QUUX-V01-CS02-with-expander-SHA256-128"
    assert S.expand_message_xmd(b"", dst, 0x20).hex() == "68a985b87eb6b46952128911f2a4412bbc302a9d759667f87f7a21d803f07235"
    assert S.expand_message_xmd(b"abc", dst, 0x20).hex() == "d8ccab23b5985ccea865c6c97b6e5b8350e794e603b4b97902f53a8a0d605615"
    assert len(S.expand_message_xmd(b"abc", b"bsb22-commitment", 48)) == 48


def test_setup_reproduces_shipped_chacha_key_structure(oracle, pk_bytes, r1cs_bytes, kat):
    from oracle import formats, setup as S
    cs = oracle.CircuitOracle(r1cs_bytes)
    pk, vk = S.setup(cs.r, b"g16-b200-chacha-selftest")
    ours, ref = formats.parse_pk_layout(pk), formats.parse_pk_layout(pk_bytes)
    assert len(pk) == len(pk_bytes) and len(vk) == (GOLDEN / "vk.chacha20").stat().st_size
    assert ours.n == ref.n and ours.hdr_fr == ref.hdr_fr
    assert ours.counts == ref.counts
    assert (ours.nb_wires, ours.nb_inf_a, ours.nb_inf_b) == (ref.nb_wires, ref.nb_inf_a, ref.nb_inf_b)
    assert np.array_equal(ours.inf_a, ref.inf_a) and np.array_equal(ours.inf_b, ref.inf_b)
    # a proof made with the self-generated pk verifies under the self-generated vk, and not under the shipped one
    prover = oracle.ChaChaOracleProver(pk, r1cs_bytes)
    proof, ct = prover.prove(kat["key"], kat["nonce"], kat["counter"], kat["input"], kat["r"], kat["s"])
    assert ct == kat["ct"] and proof != kat["proof"]
    inputs, _ = oracle.chacha_assignment(kat["key"], kat["nonce"], kat["counter"], kat["input"])
    pub = inputs[1:cs.r.n_public]
    assert oracle.VerifyingKeyOracle(vk).verify(proof, pub)
    assert not oracle.VerifyingKeyOracle((GOLDEN / "vk.chacha20").read_bytes()).verify(proof, pub)


@pytest.mark.parametrize("bits", [128, 256])
def test_aes_setup_prove_verify(oracle, bits):
    from oracle import setup as S
    pk, vk, r1 = aes_keys(bits)
    shipped_vk = (GOLDEN / f"vk.aes{bits}").read_bytes()
    # same structure as the vk gnark wrote for this circuit (different toxic waste, so different points)
    assert len(vk) == len(shipped_vk) == 5008
    assert vk[288:292] == shipped_vk[288:292] == struct.pack(">I", 143)
    tail = 292 + 143 * 32
    assert vk[tail:tail + 12] == shipped_vk[tail:tail + 12] == struct.pack(">III", 1, 0, 1)
    orc = S.AESOracleProver(r1, b"", keys=(pk, vk))
    k = AES_KAT[bits]
    r, s, m = AES_RSM
    proof, ct = orc.prove(k["key"], k["nonce"], k["counter"], k["input"], r, s, m)
    assert len(proof) == 196 and proof[128:132] == b"\x00\x00\x00\x01"
    assert ct == S.aes_ctr(k["key"], k["nonce"], k["counter"], k["input"])
    signals = ct + k["nonce"] + struct.pack(">I", k["counter"]) + k["input"]      # verifiers.go:110-119
    pub = S.aes_public_from_signals(signals)
    assert orc.verify(proof, pub)
    for pos in (3, 40, 100, 140, 170):
        bad = bytearray(proof); bad[pos] ^= 1
        assert not orc.verify(bytes(bad), pub)
    bad_pub = list(pub); bad_pub[20] ^= 1
    assert not orc.verify(proof, bad_pub)
    # the proof is a function of (input, r, s, mask): another mask moves the commitment, hence the challenge wire and
    # every point of the proof, and still verifies
    proof2, _ = orc.prove(k["key"], k["nonce"], k["counter"], k["input"], r, s, m + 1)
    assert proof2[132:164] != proof[132:164] and proof2[:32] != proof[:32] and orc.verify(proof2, pub)


def test_aes_witness_rejected_when_ciphertext_is_wrong(oracle, aes128_oracle):
    from oracle import setup as S
    k = AES_KAT[128]
    inputs, ct = S.aes_assignment(k["key"], k["nonce"], k["counter"], k["input"])
    aes128_oracle.solve(inputs, 5)
    inputs[1 + 12 + 1 + 64 + 7] ^= 1     # one ciphertext byte
    with pytest.raises(ValueError):
        aes128_oracle.solve(inputs, 5)


def test_aes_rfc3686_vectors(oracle, aes128_oracle):
    """RFC 3686 test vector #2 (the style circuits/aesV2/aes128_test.go:33-91 uses): AES-128-CTR, 32-byte plaintext;
    the oracle's CTR keystream equals the RFC's ciphertext and the assignment satisfies r1cs.aes128."""
    from oracle import setup as S
    key = bytes.fromhex("7E24067817FAE0D743D6CE1F32539163")
    nonce = bytes.fromhex("006CB6DB") + bytes.fromhex("C0543B59DA48D90B")
    pt = bytes(range(32)) + bytes(32)
    ct = S.aes_ctr(key, nonce, 1, pt)
    assert ct[:32].hex().upper() == "5104A106168A72D9790D41EE8EDAD388EB2E1EFC46DA57C8FCE630DF9141BE28"
    inputs, ct2 = S.aes_assignment(key, nonce, 1, pt)
    assert ct2 == ct
    aes128_oracle.solve(inputs, 7)
