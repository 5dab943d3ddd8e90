"""GPU parity tests of the batched groth16.Verify through the C-ABI: same accept / reject verdicts as
the oracle's pairing verifier (pinned by the reference's shipped vk.chacha20) on valid, tampered and malformed proofs."""
import struct

import numpy as np
import pytest

from conftest import AES_KAT, AES_RSM, aes_keys, batch_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gnark_symmetric_crypto_b200 as G
    return G


def test_chacha_verify_batch(G, gpu_ctx, oracle, oracle_vk, kat):
    """core_test.go:130-172 TestFullChaCha20 shape, batched: proofs from the GPU prover (and the committed KAT proof) are
    accepted under the reference's vk.chacha20; wrong public signals, spliced points and malformed encodings are rejected."""
    vk_bytes = open("tests/golden/vk.chacha20", "rb").read()
    ver = G.Groth16Verifier(vk_bytes)
    assert (ver.n_public, ver.n_commitments, ver.proof_bytes, ver.nK) == (1152, 0, 164, 1153)
    n = 6
    keys, nonces, ctrs, ins, rs = batch_inputs(n, b"g16-b200-verify")
    proofs, cts = gpu_ctx.prove_chacha_batch(keys, nonces, ctrs, ins, rs)
    pubs = [oracle.chacha_public_from_signals(cts[i] + nonces[i] + struct.pack("<I", ctrs[i]) + ins[i]) for i in range(n)]
    inputs, _ = oracle.chacha_assignment(kat["key"], kat["nonce"], kat["counter"], kat["input"])
    proofs.append(kat["proof"]); pubs.append(inputs[1:1153])
    cases, want = [], []
    for p, pub in zip(proofs, pubs):
        cases.append((p, pub)); want.append(True)
    bad_pub = list(pubs[0]); bad_pub[700] ^= 1
    cases.append((proofs[0], bad_pub)); want.append(False)                                    # one public bit flipped
    cases.append((proofs[1][:32] + proofs[2][32:], pubs[2])); want.append(False)                # Ar of another proof
    cases.append((proofs[2][:96] + proofs[3][96:128] + proofs[2][128:], pubs[2])); want.append(False)   # Krs of another proof
    cases.append((bytes([proofs[3][0] & 0x3F]) + proofs[3][1:], pubs[3])); want.append(False)  # flag bits 00: not a compressed point
    cases.append((proofs[4][:128] + b"\x00\x00\x00\x01" + proofs[4][132:], pubs[4])); want.append(False)   # commitment count
    x_not_on_curve = bytes([0x80]) + bytes(30) + bytes([5])   # x = 5: 5^3 + 3 = 128 is not a square mod p? checked by the oracle below
    cases.append((x_not_on_curve + proofs[5][32:], pubs[5])); want.append(False)
    # Bs on the twist but outside the r-torsion subgroup: gnark's G2 decoder (IsInSubGroup) rejects it, so must we
    off = None
    for t in range(2, 200):
        raw = bytearray(64); raw[63] = t; raw[31] = 1; raw[0] |= 0x80
        try:
            pt = G.decompress(2, bytes(raw))
        except G.ProverError:
            continue
        if pt.any():
            off = bytes(raw)
            assert G.g2_subgroup_check(pt).tolist() == [False]
            break
    assert off is not None
    assert G.g2_subgroup_check(oracle.g2_fixed_base(oracle.ints_to_limbs([3, 0, 77]))).tolist() == [True, True, True]
    cases.append((proofs[0][:32] + off + proofs[0][96:], pubs[0])); want.append(False)
    got = ver.verify_batch([c[0] for c in cases], [c[1] for c in cases])
    assert got.tolist() == want
    for (p, pub), w in zip(cases[:9], want[:9]):   # the oracle's verdicts on the well-formed cases
        assert oracle_vk.verify(p, pub) == w
    # gnark-shaped public witness (Montgomery fr.Element limbs) gives the same verdicts
    mont = np.stack([oracle.to_mont(1, oracle.ints_to_limbs(c[1])) for c in cases])
    assert ver.verify_batch([c[0] for c in cases], mont).tolist() == want
    assert ver.verify(kat["proof"], inputs[1:1153]) is True
    with pytest.raises(G.ProverError):
        G.Groth16Verifier(vk_bytes[:-1])
    ver.close()


@pytest.mark.parametrize("bits", [128, 256])
def test_aes_verify_batch(G, oracle, aes128_oracle, aes256_oracle, bits):
    """core_test.go:174-260 shape for the AES circuits (one BSB22 commitment: challenge public input, kSum += C, Pedersen
    proof-of-knowledge pairing), keys from the Setup restatement; verdicts equal the oracle's."""
    from oracle import setup as S
    pk, vk, r1 = aes_keys(bits)
    orc = aes128_oracle if bits == 128 else aes256_oracle
    ctx = G.Groth16Context(pk, r1, device=0)
    ver = G.Groth16Verifier(vk)
    assert (ver.n_public, ver.n_commitments, ver.proof_bytes, ver.nK) == (141, 1, 196, 143)
    rng = np.random.default_rng(bits + 1)
    n = 4
    keys = [rng.bytes(bits // 8) for _ in range(n)]; nonces = [rng.bytes(12) for _ in range(n)]
    ctrs = [int(rng.integers(0, 1 << 31)) for _ in range(n)]; ins = [rng.bytes(64) for _ in range(n)]
    proofs, cts = ctx.prove_aes_batch(keys, nonces, ctrs, ins, None)
    pubs = [S.aes_public_from_signals(cts[i] + nonces[i] + struct.pack(">I", ctrs[i]) + ins[i]) for i in range(n)]
    cases = [(proofs[i], pubs[i], True) for i in range(n)]
    bad_pub = list(pubs[0]); bad_pub[20] ^= 1
    cases.append((proofs[0], bad_pub, False))
    cases.append((proofs[1][:132] + proofs[2][132:164] + proofs[1][164:], pubs[1], False))   # commitment of another proof
    cases.append((proofs[2][:164] + proofs[3][164:], pubs[2], False))                          # PoK of another proof
    cases.append((proofs[3][:128] + b"\x00\x00\x00\x00" + proofs[3][132:], pubs[3], False))   # commitment count
    got = ver.verify_batch([c[0] for c in cases], [c[1] for c in cases])
    assert got.tolist() == [c[2] for c in cases]
    for p, pub, w in cases[:7]:
        assert orc.verify(p, pub) == w
    ver.close(); ctx.close()


def test_libverify_abi_round_trip(G, oracle, pk_bytes, r1cs_bytes):
    """core_test.go:130-172 TestFullChaCha20 and :174-260 TestFullAES128/256 end to end through the outer ABI of both
    libraries: InitAlgorithm + Prove(JSON) -> InitVerifier + Verify(JSON); core_test.go:120-128 TestPanic's payload and other
    broken inputs verify as false."""
    vk = open("tests/golden/vk.chacha20", "rb").read()
    assert G.InitAlgorithm(G.CHACHA20, pk_bytes, r1cs_bytes) is True
    assert G.InitVerifier(G.CHACHA20, vk) is True
    assert G.InitVerifier(G.CHACHA20, vk) is True
    assert G.InitVerifier(9, vk) is False
    rng = np.random.default_rng(5)
    key, nonce, pt, counter = rng.bytes(32), rng.bytes(12), rng.bytes(64), 7
    out = G.OutputParams.from_json(G.Prove(G.InputParams("chacha20", key, nonce, counter, pt).to_json()))
    signals = out.public_signals + nonce + struct.pack("<I", counter) + pt      # core_test.go:157-163
    assert G.Verify(G.InputVerifyParams("chacha20", out.proof_json, signals).to_json()) is True
    wrong = bytearray(signals); wrong[3] ^= 1
    assert G.Verify(G.InputVerifyParams("chacha20", out.proof_json, bytes(wrong)).to_json()) is False
    assert G.Verify(G.InputVerifyParams("chacha20", out.proof_json, signals[:-1]).to_json()) is False
    assert G.Verify(G.InputVerifyParams("chacha20", out.proof_json[:100], signals).to_json()) is False
    assert G.Verify(G.InputVerifyParams("chacha21", out.proof_json, signals).to_json()) is False
    assert G.Verify(b'{"cipher":"aes-256-ctr1","key":[0],"nonce":[0],"counter":[0,1],"input":[0]}') is False
    assert G.Verify(b'not json') is False
    for alg, name, bits in ((G.AES_128, "aes-128-ctr", 128), (G.AES_256, "aes-256-ctr", 256)):
        pk, avk, r1 = aes_keys(bits)
        assert G.InitAlgorithm(alg, pk, r1) is True
        assert G.InitVerifier(alg, avk) is True
        key = rng.bytes(bits // 8)
        out = G.OutputParams.from_json(G.Prove(G.InputParams(name, key, nonce, counter, pt).to_json()))
        signals = out.public_signals + nonce + struct.pack(">I", counter) + pt   # core_test.go:205,249: big-endian counter
        assert G.Verify(G.InputVerifyParams(name, out.proof_json, signals).to_json()) is True
        wrong = bytearray(signals); wrong[70] ^= 1
        assert G.Verify(G.InputVerifyParams(name, out.proof_json, bytes(wrong)).to_json()) is False
