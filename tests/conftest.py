import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pk_bytes():
    return (GOLDEN / "pk.chacha20").read_bytes()


@pytest.fixture(scope="session")
def r1cs_bytes():
    return (GOLDEN / "r1cs.chacha20").read_bytes()


@pytest.fixture(scope="session")
def vk_bytes():
    return (GOLDEN / "vk.chacha20").read_bytes()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def oracle_prover(oracle, pk_bytes, r1cs_bytes):
    return oracle.ChaChaOracleProver(pk_bytes, r1cs_bytes)


@pytest.fixture(scope="session")
def oracle_vk(oracle, vk_bytes):
    return oracle.VerifyingKeyOracle(vk_bytes)


def aes_keys(bits: int):
    """(pk, vk, r1cs) for AES-128/256. The reference ships no pk.aes* (.MISSING_LARGE_BLOBS), so the keys come from the
    oracle's Setup restatement on the reference's r1cs.aes*, toxic waste from the seed "g16-b200-aes<bits>" (SURVEY §8d
    config 2/3). Cached under tests/golden/_gen (git-ignored; regenerated in ~30 s when absent)."""
    from oracle import oracle as O, setup as S
    r1 = (GOLDEN / f"r1cs.aes{bits}").read_bytes()
    gen = GOLDEN / "_gen"
    pkp, vkp = gen / f"pk.aes{bits}", gen / f"vk.aes{bits}"
    if not (pkp.exists() and vkp.exists()):
        gen.mkdir(exist_ok=True)
        pk, vk = S.setup(O.CircuitOracle(r1).r, f"g16-b200-aes{bits}".encode())
        tmp = gen / f".pk.aes{bits}.{os.getpid()}"
        tmp.write_bytes(pk); tmp.replace(pkp)
        tmp.write_bytes(vk); tmp.replace(vkp)
    return pkp.read_bytes(), vkp.read_bytes(), r1


@pytest.fixture(scope="session")
def aes128_oracle(oracle):
    from oracle import setup as S
    pk, vk, r1 = aes_keys(128)
    return S.AESOracleProver(r1, b"", keys=(pk, vk))


@pytest.fixture(scope="session")
def aes256_oracle(oracle):
    from oracle import setup as S
    pk, vk, r1 = aes_keys(256)
    return S.AESOracleProver(r1, b"", keys=(pk, vk))


# libraries/core_test.go:265 / :275 — the reference's AES benchmark inputs (BASELINE configs 2, 3); fixed r, s, mask
AES_KAT = {
    128: dict(key=bytes([2]) * 16, nonce=bytes([3]) * 12, counter=2,
              input=bytes([183, 4, 206, 60, 254, 21, 117, 9, 150, 227, 246, 245, 71, 101, 56, 67, 79, 93, 44, 163, 22, 89, 128, 55, 214,
                           254, 228, 214, 89, 253, 176, 112, 138, 115, 93, 140, 194, 222, 104, 252, 49, 144, 91, 252]) + bytes(20)),
    256: dict(key=bytes([2]) * 32, nonce=bytes([3]) * 12, counter=10,
              input=bytes([189, 250, 225, 242, 6, 46, 173, 203, 7, 166, 62, 139, 67, 150, 1, 155, 64, 122, 211, 198, 184, 203, 124, 194,
                           99, 34, 127, 29, 236, 17, 232, 214, 154, 146, 78, 217, 254, 224, 208, 196, 55, 200, 23, 93, 90, 175, 240, 31,
                           31, 225, 26, 15, 219, 156, 123, 21, 103, 98, 205, 87, 197, 22, 245, 158])),
}
AES_RSM = (int("11" * 20, 16), int("22" * 20, 16), int("33" * 20, 16))


@pytest.fixture(scope="session")
def emu():
    """TEST-ONLY host-emulation build of the product sources (tests/emu). Never reachable through the package API."""
    from gnark_symmetric_crypto_b200 import _lib
    if os.environ.get("G16_EMU_SO"):   # scripts/emu_sanitize.sh: the ASan / TSan build of the same sources
        return _lib.bind(ROOT / os.environ["G16_EMU_SO"])
    subprocess.check_call(["make", "-C", str(ROOT / "tests" / "emu"), "-j", str(os.cpu_count() or 1), "-s"])
    return _lib.bind(ROOT / "tests" / "emu" / "_build" / "libg16emu.so")


@pytest.fixture(scope="session")
def gpu_ctx(pk_bytes, r1cs_bytes):
    import gnark_symmetric_crypto_b200 as G
    ctx = G.Groth16Context(pk_bytes, r1cs_bytes, device=0)
    yield ctx
    ctx.close()


# the reference's own benchmark inputs, libraries/core_test.go:285 (config 1) + SURVEY.md Appendix H
KAT = dict(
    key=bytes([2]) * 32, nonce=bytes([3]) * 12, counter=3,
    input=bytes.fromhex("a3f7e592aeda1507a7f51b35812dfc50a263d5a6d2df625e563b02e49c08bf30"
                        "d0e7483f5b13ff079532224ee8fbc31ab1899b18e453d36d9793a8355eb0dee9"),
    ct=bytes.fromhex("e11ef0b2e6d3e450ab1a3509c0a6a2c79ece1376a8a0a6c09603f26b15b106de"
                     "e60711d709ca21ac7e545f7d2c040f1ba1933d4eff4823a142da7aaffa483224"),
    r=int("11" * 20, 16), s=int("22" * 20, 16),
    proof=bytes.fromhex(
        "d73f52bc6800d1c4a07c95d0b21876d2ed029d442b2df690a2fe2a711b77f6e1"
        "95200aa0384e7f1ea31d47954f1fa672350b66ecf2608c5db062aa7f9ff7153b"
        "0f0896aa890cca5296834e8bf266931d6df1f412d99cfcf9d5786b7f3e7cb441"
        "ddfbcad67781ec5c223db4246c4c4e3860638f210422b7c296e145b1df4153f8"
        "00000000" "40" + "00" * 31),
)


@pytest.fixture(scope="session")
def kat():
    return KAT


def batch_inputs(n, seed=b"g16-b200-batch"):
    """BASELINE config 4 input stream: ChaCha20(key=SHA-256(seed), nonce=0) keystream cut into
    key(32) | nonce(12) | counter(4, LE) | input(64) | r(32) | s(32) per request; r, s reduced mod the group order."""
    import hashlib
    import struct
    from oracle import oracle as O
    k = hashlib.sha256(seed).digest()
    per = 32 + 12 + 4 + 64 + 64
    nblocks = (n * per + 63) // 64
    stream = b"".join(O.chacha20_block(k, i, bytes(12)) for i in range(nblocks))
    keys, nonces, ctrs, ins, rs = [], [], [], [], []
    for i in range(n):
        b = stream[i * per:(i + 1) * per]
        keys.append(b[:32]); nonces.append(b[32:44]); ctrs.append(struct.unpack("<I", b[44:48])[0]); ins.append(b[48:112])
        r = int.from_bytes(b[112:144], "big") % O.R_MOD
        s = int.from_bytes(b[144:176], "big") % O.R_MOD
        rs.append(r.to_bytes(32, "big") + s.to_bytes(32, "big"))
    return keys, nonces, ctrs, ins, rs
