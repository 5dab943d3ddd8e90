"""ORACLE — TEST INFRASTRUCTURE ONLY. Not imported by the product package (gnark_symmetric_crypto_b200/).

ctypes front-end for oracle/_build/liboracle.so plus the library-level logic of the reference restated in Python:
  * ChaCha20 witness assignment  — libraries/prover/impl/provers.go:79-158, utils/bytes.go:11-47
  * AES witness assignment       — libraries/prover/impl/provers.go:172-227
  * public-signal slicing        — libraries/verifier/impl/verifiers.go:50-152
Parity status: pinned to the reference through (1) pairing verification against the reference's shipped vk.chacha20,
(2) the SURVEY.md Appendix H proof KAT, (3) RFC 7539 vectors quoted in circuits/chachaV3/chacha_test.go:22-30,95-105.
gnark itself (go.mod:8-9) cannot run on this box.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
from pathlib import Path

import numpy as np

from . import formats
from .formats import R_MOD, P_MOD

_HERE = Path(__file__).resolve().parent
_LIB = None

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> Path:
    so = _HERE / "_build" / "liboracle.so"
    v3 = _HERE / "_build" / "liboracle_v3.so"
    srcs = [_HERE / n for n in ("oracle_core.cpp", "oracle_groth16.cpp", "bn254_field.hpp", "bn254_curve.hpp")]
    if force or not so.exists() or not v3.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs if s.exists()):
        subprocess.check_call(["make", "-C", str(_HERE), "-s"])
    return so


def _cpu_has(*flags) -> bool:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                have = set(line.split(":", 1)[1].split())
                return all(f in have for f in flags)
    except OSError:
        pass
    return False


BUILD_VARIANT = "x86-64-v2"


def lib():
    global _LIB, BUILD_VARIANT
    if _LIB is None:
        so = build()
        v3 = so.with_name("liboracle_v3.so")
        # the faster build (BMI2 mulx, ADX, AVX2) when this host supports it; ORACLE_PORTABLE=1 forces the portable one
        if v3.exists() and not os.environ.get("ORACLE_PORTABLE") and _cpu_has("avx2", "bmi2", "adx", "fma", "movbe", "abm"):
            so = v3
            BUILD_VARIANT = "x86-64-v3+adx"
        L = C.CDLL(str(so))
        L.orc_pk_parse.restype = C.c_void_p
        L.orc_pk_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_int]
        L.orc_pk_array.restype = C.c_void_p
        L.orc_pk_array.argtypes = [C.c_void_p, C.c_int]
        L.orc_pk_info.argtypes = [C.c_void_p, u64p]
        L.orc_pk_free.argtypes = [C.c_void_p]
        L.orc_vk_parse.restype = C.c_void_p
        L.orc_vk_parse.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_vk_array.restype = C.c_void_p
        L.orc_vk_array.argtypes = [C.c_void_p, C.c_int]
        L.orc_vk_nk.restype = C.c_size_t
        L.orc_vk_nk.argtypes = [C.c_void_p]
        L.orc_vk_free.argtypes = [C.c_void_p]
        L.orc_verify.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_init()
        e = (P_MOD ** 12 - 1) // R_MOD
        nl = (e.bit_length() + 63) // 64
        limbs = np.array([(e >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(nl)], dtype=np.uint64)
        L.orc_pairing_init(limbs.ctypes.data_as(u64p), C.c_size_t(nl))
        _LIB = L
    return _LIB


def _p(a, t=u64p):
    return a.ctypes.data_as(t)


NTHREADS = max(1, os.cpu_count() or 1)

# ----------------------------------------------------------------------------- field helpers
FP, FR = 0, 1


def ints_to_limbs(vals) -> np.ndarray:
    """python ints -> [n,4] u64 little-endian limbs"""
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for k in range(4):
            out[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(a: np.ndarray) -> list:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    return [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in a]


def to_mont(fld: int, canon: np.ndarray) -> np.ndarray:
    canon = np.ascontiguousarray(canon, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(canon)
    lib().orc_f_to_mont(fld, _p(canon), _p(out), C.c_size_t(len(canon)))
    return out


def from_mont(fld: int, mont: np.ndarray) -> np.ndarray:
    mont = np.ascontiguousarray(mont, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(mont)
    lib().orc_f_from_mont(fld, _p(mont), _p(out), C.c_size_t(len(mont)))
    return out


def f_op(fld: int, op: str, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
    code = {"add": 0, "sub": 1, "mul": 2, "inv": 3, "sqr": 4, "neg": 5}[op]
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(a)
    bp = None
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        bp = _p(b)
    lib().orc_f_op(fld, code, _p(a), bp, _p(out), C.c_size_t(len(a)))
    return out


def rand_field(rng: np.random.Generator, fld: int, n: int) -> np.ndarray:
    """n uniform field elements as canonical limbs [n,4]"""
    mod = P_MOD if fld == FP else R_MOD
    raw = rng.integers(0, 1 << 64, size=(n, 5), dtype=np.uint64)
    vals = [(int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 | int(r[4]) << 256) % mod for r in raw]
    return ints_to_limbs(vals)


# ----------------------------------------------------------------------------- points
G1_GEN = None
G2_GEN_INTS = (
    10857046999023057135944570762232829481370756359578518086990519993285655852781,   # x.a0
    11559732032986387107991004021392285783925812861821192530917403151452391805634,   # x.a1
    8495653923123431417604973247489272438418190587263600148770280649306958101930,    # y.a0
    4082367875863433681332203403145435568316851327593401208105741076214120093531,    # y.a1
)


def g1_gen() -> np.ndarray:
    return to_mont(FP, ints_to_limbs([1, 2])).reshape(8)


def g2_gen() -> np.ndarray:
    return to_mont(FP, ints_to_limbs(list(G2_GEN_INTS))).reshape(16)


def g1_decompress(raw: bytes | np.ndarray) -> np.ndarray:
    raw = np.frombuffer(bytes(raw), dtype=np.uint8) if not isinstance(raw, np.ndarray) else np.ascontiguousarray(raw)
    n = raw.size // 32
    out = np.empty((n, 8), dtype=np.uint64)
    rc = lib().orc_g1_decompress(_p(raw, u8p), _p(out), C.c_size_t(n), NTHREADS)
    if rc:
        raise ValueError(f"g1 decompress failed rc={rc}")
    return out


def g2_decompress(raw: bytes | np.ndarray) -> np.ndarray:
    raw = np.frombuffer(bytes(raw), dtype=np.uint8) if not isinstance(raw, np.ndarray) else np.ascontiguousarray(raw)
    n = raw.size // 64
    out = np.empty((n, 16), dtype=np.uint64)
    rc = lib().orc_g2_decompress(_p(raw, u8p), _p(out), C.c_size_t(n), NTHREADS)
    if rc:
        raise ValueError(f"g2 decompress failed rc={rc}")
    return out


def g1_compress(pts: np.ndarray) -> bytes:
    pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, 8)
    out = np.empty(32 * len(pts), dtype=np.uint8)
    lib().orc_g1_compress(_p(pts), _p(out, u8p), C.c_size_t(len(pts)))
    return out.tobytes()


def g2_compress(pts: np.ndarray) -> bytes:
    pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, 16)
    out = np.empty(64 * len(pts), dtype=np.uint8)
    lib().orc_g2_compress(_p(pts), _p(out, u8p), C.c_size_t(len(pts)))
    return out.tobytes()


def g1_msm(pts: np.ndarray, scalars_canon: np.ndarray, naive: bool = False, nthreads: int | None = None) -> np.ndarray:
    pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, 8)
    sc = np.ascontiguousarray(scalars_canon, dtype=np.uint64).reshape(-1, 4)
    assert len(pts) == len(sc)
    out = np.empty(8, dtype=np.uint64)
    lib().orc_g1_msm(_p(pts), _p(sc), C.c_size_t(len(pts)), nthreads or NTHREADS, int(naive), _p(out))
    return out


def g2_msm(pts: np.ndarray, scalars_canon: np.ndarray, naive: bool = False, nthreads: int | None = None) -> np.ndarray:
    pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, 16)
    sc = np.ascontiguousarray(scalars_canon, dtype=np.uint64).reshape(-1, 4)
    assert len(pts) == len(sc)
    out = np.empty(16, dtype=np.uint64)
    lib().orc_g2_msm(_p(pts), _p(sc), C.c_size_t(len(pts)), nthreads or NTHREADS, int(naive), _p(out))
    return out


def g1_fixed_base(ks_canon: np.ndarray, base: np.ndarray | None = None) -> np.ndarray:
    ks = np.ascontiguousarray(ks_canon, dtype=np.uint64).reshape(-1, 4)
    base = g1_gen() if base is None else np.ascontiguousarray(base, dtype=np.uint64)
    out = np.empty((len(ks), 8), dtype=np.uint64)
    lib().orc_g1_fixed_base(_p(base), _p(ks), C.c_size_t(len(ks)), NTHREADS, _p(out))
    return out


def g2_fixed_base(ks_canon: np.ndarray, base: np.ndarray | None = None) -> np.ndarray:
    ks = np.ascontiguousarray(ks_canon, dtype=np.uint64).reshape(-1, 4)
    base = g2_gen() if base is None else np.ascontiguousarray(base, dtype=np.uint64)
    out = np.empty((len(ks), 16), dtype=np.uint64)
    lib().orc_g2_fixed_base(_p(base), _p(ks), C.c_size_t(len(ks)), NTHREADS, _p(out))
    return out


def g1_add(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty(8, dtype=np.uint64)
    lib().orc_g1_add(_p(a), _p(b), _p(out))
    return out


def g2_add(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty(16, dtype=np.uint64)
    lib().orc_g2_add(_p(a), _p(b), _p(out))
    return out


def g1_mul(a, k: int):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    kk = ints_to_limbs([k % R_MOD])
    out = np.empty(8, dtype=np.uint64)
    lib().orc_g1_mul(_p(a), _p(kk), _p(out))
    return out


def g2_mul(a, k: int):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    kk = ints_to_limbs([k % R_MOD])
    out = np.empty(16, dtype=np.uint64)
    lib().orc_g2_mul(_p(a), _p(kk), _p(out))
    return out


def pairing_check(g1s: np.ndarray, g2s: np.ndarray) -> bool:
    g1s = np.ascontiguousarray(g1s, dtype=np.uint64).reshape(-1, 8)
    g2s = np.ascontiguousarray(g2s, dtype=np.uint64).reshape(-1, 16)
    return bool(lib().orc_pairing_check(_p(g1s), _p(g2s), C.c_size_t(len(g1s))))


# ----------------------------------------------------------------------------- NTT / H
def root_of_unity(n: int) -> int:
    root28 = 19103219067921713944291392827692070036145651957329286315305642004821462161904
    lg = n.bit_length() - 1
    assert 1 << lg == n and lg <= 28
    return pow(root28, 1 << (28 - lg), R_MOD)


def ntt(data_mont: np.ndarray, inverse: bool = False) -> np.ndarray:
    d = np.ascontiguousarray(data_mont, dtype=np.uint64).reshape(-1, 4).copy()
    n = len(d)
    w = to_mont(FR, ints_to_limbs([root_of_unity(n)]))
    lib().orc_ntt(_p(d), C.c_size_t(n), _p(w), int(inverse))
    return d


def bitrev_perm(n: int) -> np.ndarray:
    lg = n.bit_length() - 1
    idx = np.arange(n, dtype=np.uint64)
    rev = np.zeros(n, dtype=np.uint64)
    for b in range(lg):
        rev |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(lg - 1 - b)
    return rev.astype(np.int64)


def compute_h(a: np.ndarray, b: np.ndarray, c: np.ndarray, n: int, coset_gen: int = 5, nthreads: int | None = None) -> np.ndarray:
    """a,b,c: [ncons,4] Montgomery. returns h [n,4] Montgomery, NATURAL coefficient order."""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    c = np.ascontiguousarray(c, dtype=np.uint64).reshape(-1, 4)
    w = to_mont(FR, ints_to_limbs([root_of_unity(n)]))
    g = to_mont(FR, ints_to_limbs([coset_gen]))
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_compute_h(_p(a), _p(b), _p(c), C.c_size_t(len(a)), C.c_size_t(n), _p(w), _p(g), _p(out),
                        nthreads or NTHREADS)
    return out


# ----------------------------------------------------------------------------- ciphers (cipher oracles of the reference)
def _rotl(x, n):
    return ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF


def chacha20_block(key: bytes, counter: int, nonce: bytes) -> bytes:
    """RFC 7539 §2.3 block function (the reference uses golang.org/x/crypto/chacha20, provers.go:95-101)."""
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(struct.unpack("<8I", key)) + [counter & 0xFFFFFFFF] + \
        list(struct.unpack("<3I", nonce))
    x = st[:]

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return struct.pack("<16I", *[(x[i] + st[i]) & 0xFFFFFFFF for i in range(16)])


def chacha20_xor(key: bytes, nonce: bytes, counter: int, data: bytes) -> bytes:
    out = bytearray()
    for blk in range(0, len(data), 64):
        ks = chacha20_block(key, counter + blk // 64, nonce)
        out += bytes(a ^ b for a, b in zip(data[blk:blk + 64], ks))
    return bytes(out)


def _word_bits(words) -> list:
    bits = []
    for w in words:
        bits.extend((w >> i) & 1 for i in range(32))
    return bits


def chacha_assignment(key: bytes, nonce: bytes, counter: int, plaintext: bytes):
    """-> (input wire values as ints [ONE, public..., secret...], ciphertext). Mirrors provers.go:79-142.
    Witness order = public first in declaration order (Counter, Nonce, In, Out), then secret (Key): SURVEY §8(a5)."""
    if len(key) != 32:
        raise ValueError(f"key length must be 32: {len(key)}")
    if len(nonce) != 12:
        raise ValueError(f"nonce length must be 12: {len(nonce)}")
    if len(plaintext) != 64:
        raise ValueError(f"plaintext length must be 64: {len(plaintext)}")
    ct = chacha20_xor(key, nonce, counter, plaintext)
    b_pt = _word_bits(struct.unpack(">16I", plaintext))      # BytesToUint32BEBits
    b_ct = _word_bits(struct.unpack(">16I", ct))
    b_key = _word_bits(struct.unpack("<8I", key))            # BytesToUint32LEBits
    b_nonce = _word_bits(struct.unpack("<3I", nonce))
    b_ctr = _word_bits([counter & 0xFFFFFFFF])
    return [1] + b_ctr + b_nonce + b_pt + b_ct + b_key, ct


def chacha_public_from_signals(signals: bytes) -> list:
    """verifiers.go:50-87: 144-byte publicSignals = ct(64)|nonce(12)|counter(4, LE)|pt(64) -> public witness (no ONE)."""
    if len(signals) != 144:
        raise ValueError("public signals must be 144 bytes")
    ct, nonce, ctr, pt = signals[:64], signals[64:76], signals[76:80], signals[80:]
    return (_word_bits(struct.unpack("<1I", ctr)) + _word_bits(struct.unpack("<3I", nonce)) +
            _word_bits(struct.unpack(">16I", pt)) + _word_bits(struct.unpack(">16I", ct)))


# ----------------------------------------------------------------------------- solver
class CircuitOracle:
    """Parsed r1cs + the flattened views orc_solve consumes."""

    def __init__(self, r1cs_bytes: bytes):
        self.r = formats.parse_r1cs(r1cs_bytes)
        r = self.r
        self.kind = np.array([r.bp_kind[b] for b in r.bp_id], dtype=np.uint8)
        lk_ids = sorted(r.lookup_entries)
        self.lk_index = {b: i for i, b in enumerate(lk_ids)}
        self.ins_tab = np.array([self.lk_index.get(int(b), 0) for b in r.bp_id], dtype=np.uint32)
        tabs = np.zeros((max(1, len(lk_ids)), 256, 4), dtype=np.uint64)
        for b in lk_ids:
            ents = r.lookup_entries[b]
            assert len(ents) == 256
            for k, e in enumerate(ents):
                assert len(e) == 1 and e[0][1] == 0xFFFFFFFF, e
                tabs[self.lk_index[b], k] = r.coeffs[e[0][0]]
        self.tabs = tabs
        self.start = np.ascontiguousarray(r.start, dtype=np.uint64)
        self.wire_off = np.ascontiguousarray(r.wire_off, dtype=np.uint32)
        self.cons_off = np.ascontiguousarray(r.cons_off, dtype=np.uint32)
        self.calldata = np.ascontiguousarray(r.calldata, dtype=np.uint32)
        self.coeffs = np.ascontiguousarray(r.coeffs, dtype=np.uint64)

    def solve(self, inputs: list, randomize: int | None = None, bsb22=None):
        """inputs: python ints for wires [0, n_public+n_secret). -> (W, A, B, C) Montgomery [.,4] arrays"""
        r = self.r
        n_in = r.n_public + r.n_secret
        assert len(inputs) == n_in, (len(inputs), n_in)
        W = np.zeros((r.n_wires, 4), dtype=np.uint64)
        W[:n_in] = to_mont(FR, ints_to_limbs([v % R_MOD for v in inputs]))
        A = np.zeros((r.n_constraints, 4), dtype=np.uint64)
        B = np.zeros_like(A)
        Cc = np.zeros_like(A)
        rnd = None
        if randomize is not None:
            rnd = to_mont(FR, ints_to_limbs([randomize % R_MOD]))
        CB = C.CFUNCTYPE(C.c_int, C.c_void_p, u64p, C.c_size_t, u64p)
        cb = CB(bsb22) if bsb22 is not None else C.cast(None, CB)
        L = lib()
        L.orc_solve.restype = C.c_int
        rc = L.orc_solve(_p(self.calldata, u32p), C.c_size_t(r.n_instr), _p(self.kind, u8p), _p(self.start),
                         _p(self.wire_off, u32p), _p(self.cons_off, u32p), _p(self.ins_tab, u32p), _p(self.coeffs),
                         _p(self.tabs), _p(W), C.c_size_t(r.n_wires), C.c_size_t(n_in), _p(A), _p(B), _p(Cc),
                         _p(rnd) if rnd is not None else None, cb, None)
        if rc:
            raise ValueError(f"solver failed rc={rc}")
        return W, A, B, Cc


# ----------------------------------------------------------------------------- prover / verifier
class ProvingKeyOracle:
    def __init__(self, pk_bytes: bytes):
        self.h = lib().orc_pk_parse(pk_bytes, len(pk_bytes), NTHREADS)
        if not self.h:
            raise ValueError("pk parse failed")
        info = np.zeros(8, dtype=np.uint64)
        lib().orc_pk_info(self.h, _p(info))
        self.n, self.nA, self.nB, self.nZ, self.nK, self.nB2, self.nb_wires, self.n_commit = [int(x) for x in info]

    def array(self, name: str) -> np.ndarray:
        which = {"A": 0, "B": 1, "Z": 2, "K": 3, "B2": 4, "infA": 5, "infB": 6, "g1": 7, "g2": 8, "domain": 9}[name]
        ptr = lib().orc_pk_array(self.h, which)
        if name in ("infA", "infB"):
            return np.ctypeslib.as_array(C.cast(ptr, u8p), shape=(self.nb_wires,)).copy()
        shape = {"A": (self.nA, 8), "B": (self.nB, 8), "Z": (self.nZ, 8), "K": (self.nK, 8), "B2": (self.nB2, 16),
                 "g1": (3, 8), "g2": (2, 16), "domain": (5, 4)}[name]
        return np.ctypeslib.as_array(C.cast(ptr, u64p), shape=shape).copy()

    def prove(self, W, A, B, Cc, n_public: int, r: int, s: int, skip=(), nthreads: int | None = None,
              want_h: bool = False):
        """-> (proof bytes Ar|Bs|Krs (128 B), inter dict, h natural or None)"""
        W = np.ascontiguousarray(W, dtype=np.uint64); A = np.ascontiguousarray(A, dtype=np.uint64)
        B = np.ascontiguousarray(B, dtype=np.uint64); Cc = np.ascontiguousarray(Cc, dtype=np.uint64)
        out = np.zeros(128, dtype=np.uint8)
        inter = np.zeros(88, dtype=np.uint64)
        h = np.zeros((self.n, 4), dtype=np.uint64) if want_h else None
        skip_a = np.array(list(skip), dtype=np.uint32)
        rr = ints_to_limbs([r % R_MOD]); ss = ints_to_limbs([s % R_MOD])
        L = lib()
        L.orc_prove.restype = C.c_int
        rc = L.orc_prove(C.c_void_p(self.h), _p(W), _p(A), _p(B), _p(Cc), C.c_size_t(len(A)), C.c_size_t(n_public),
                         _p(skip_a, u32p), C.c_size_t(len(skip_a)), _p(rr), _p(ss), nthreads or NTHREADS,
                         _p(out, u8p), _p(inter), _p(h) if h is not None else None)
        if rc:
            raise ValueError(f"prove failed rc={rc}")
        names = [("msmA", 0, 8), ("msmB1", 8, 8), ("msmK", 16, 8), ("msmZ", 24, 8), ("msmB2", 32, 16), ("Ar", 48, 8),
                 ("Bs1", 56, 8), ("Krs", 64, 8), ("Bs", 72, 16)]
        return out.tobytes(), {k: inter[o:o + l].copy() for k, o, l in names}, h

    def __del__(self):
        try:
            lib().orc_pk_free(C.c_void_p(self.h))
        except Exception:
            pass


def serialize_proof_chacha(abc128: bytes) -> bytes:
    """gnark Proof.WriteTo with zero commitments: 128 B points | u32 0 | infinity PoK (32 B)."""
    return abc128 + b"\x00\x00\x00\x00" + b"\x40" + b"\x00" * 31


class VerifyingKeyOracle:
    def __init__(self, vk_bytes: bytes):
        self.h = lib().orc_vk_parse(vk_bytes, len(vk_bytes))
        if not self.h:
            raise ValueError("vk parse failed")
        self.nK = int(lib().orc_vk_nk(C.c_void_p(self.h)))

    def array(self, name):
        which = {"g1": 0, "g2": 1, "K": 2}[name]
        ptr = lib().orc_vk_array(C.c_void_p(self.h), which)
        shape = {"g1": (3, 8), "g2": (3, 16), "K": (self.nK, 8)}[name]
        return np.ctypeslib.as_array(C.cast(ptr, u64p), shape=shape).copy()

    def verify(self, proof: bytes, public_ints: list) -> bool:
        pub = to_mont(FR, ints_to_limbs([v % R_MOD for v in public_ints]))
        rc = lib().orc_verify(C.c_void_p(self.h), proof, len(proof), _p(pub), C.c_size_t(len(pub)))
        if rc < 0:
            raise ValueError(f"malformed proof rc={rc}")
        return rc == 1

    def __del__(self):
        try:
            lib().orc_vk_free(C.c_void_p(self.h))
        except Exception:
            pass


class ChaChaOracleProver:
    """End-to-end CPU restatement of `prover.Prove` for cipher "chacha20" (prove_impl.go:116-143, provers.go:79-158)."""

    def __init__(self, pk_bytes: bytes, r1cs_bytes: bytes):
        self.pk = ProvingKeyOracle(pk_bytes)
        self.cs = CircuitOracle(r1cs_bytes)

    def prove(self, key: bytes, nonce: bytes, counter: int, plaintext: bytes, r: int, s: int, nthreads=None, detail=False):
        inputs, ct = chacha_assignment(key, nonce, counter, plaintext)
        W, A, B, Cc = self.cs.solve(inputs)
        p128, inter, h = self.pk.prove(W, A, B, Cc, self.cs.r.n_public, r, s, nthreads=nthreads, want_h=detail)
        proof = serialize_proof_chacha(p128)
        if detail:
            return proof, ct, dict(W=W, A=A, B=B, C=Cc, h=h, **inter)
        return proof, ct
