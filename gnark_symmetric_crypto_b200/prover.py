"""Host-side mirror of the reference's prover package, over the C-ABI of libg16b200.so.

Same names, argument meaning and error behaviour as `libraries/prover/impl` of reclaimprotocol/gnark-symmetric-crypto:

    CHACHA20, AES_128, AES_256            prove_impl.go:15-19
    InitAlgorithm(id, pk, r1cs) -> bool   prove_impl.go:65-114   (False + message on stdout for bad ids / keys)
    Prove(params: bytes) -> bytes         prove_impl.go:116-143  (raises = the reference's panic, cf. TestPanic
                                                                  libraries/core_test.go:120-128)
    InputParams / OutputParams            provers.go:53-59, prove_impl.go:45-52

plus the batched entry the B200 backend adds: `Groth16Context.prove_chacha_batch`.
Everything computes on the GPU through ctypes; there is no CPU path in this module.
"""
from __future__ import annotations

import base64
import ctypes as C
import json
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import GoSlice, u8p, u32p, u64p, f32p

CHACHA20 = 0
AES_128 = 1
AES_256 = 2

_STATUS = {1: "bad argument", 2: "parse error", 3: "CUDA error", 4: "unsatisfied witness", 5: "unsupported", 6: "bad state"}


class ProverError(RuntimeError):
    def __init__(self, rc: int, msg: str):
        super().__init__(f"g16 status {rc} ({_STATUS.get(rc, '?')}): {msg}")
        self.rc = rc


def _check(rc: int):
    if rc:
        raise ProverError(rc, _lib.load().g16_last_error().decode(errors="replace"))


def _p8(a):
    return a.ctypes.data_as(u8p)


def _p64(a):
    return a.ctypes.data_as(u64p)


@dataclass
class InputParams:   # provers.go:53-59
    cipher: str
    key: bytes
    nonce: bytes
    counter: int
    input: bytes

    def to_json(self) -> bytes:
        return json.dumps({"cipher": self.cipher, "key": list(self.key), "nonce": list(self.nonce),
                           "counter": self.counter, "input": list(self.input)}).encode()


@dataclass
class OutputParams:   # prove_impl.go:45-52
    proof_json: bytes
    public_signals: bytes

    @staticmethod
    def from_json(data: bytes) -> "OutputParams":
        d = json.loads(data)
        return OutputParams(base64.b64decode(d["proof"]["proofJson"]), base64.b64decode(d["publicSignals"]))


def _slice(b: bytes):
    buf = (C.c_uint8 * len(b)).from_buffer_copy(b) if len(b) else (C.c_uint8 * 1)()
    return GoSlice(C.cast(buf, C.c_void_p), len(b), len(b)), buf


def InitAlgorithm(algorithm_id: int, proving_key: bytes, r1cs: bytes) -> bool:
    """libprove.go:20-23 / prove_impl.go:65-114 — through the library's own `InitAlgorithm` export."""
    L = _lib.load()
    if not 0 <= int(algorithm_id) <= 255:
        return False
    s1, k1 = _slice(proving_key)
    s2, k2 = _slice(r1cs)
    return bool(L.InitAlgorithm(int(algorithm_id), s1, s2))


def Prove(params: bytes) -> bytes:
    """libprove.go:30-47 — returns the JSON payload; a payload that is a bare JSON string is the reference's
    panic-as-payload convention and is raised here as RuntimeError (Go callers of impl.Prove see a panic)."""
    L = _lib.load()
    s, keep = _slice(bytes(params))
    r = L.Prove(s)
    try:
        out = C.string_at(r.r0, r.r1)
    finally:
        L.Free(r.r0)
    _raise_if_panic(out)
    return out


def _raise_if_panic(out: bytes):
    """libprove.go:33-43: a recovered panic comes back as json.Marshal(value) — a JSON string for string panics
    (log.Panicf, fmt.Sprintf), the marshalled error VALUE (an object without "proof") for panic(err)."""
    if out[:1] == b'"':
        raise RuntimeError(json.loads(out))
    if out[:1] != b"{" or b'"proof"' not in out[:16]:
        raise RuntimeError("prover panicked with error value " + out.decode(errors="replace"))


def ProveBatch(params_list) -> list:
    """The batched twin of Prove: a list of InputParams JSON byte strings (or dicts) -> a list with,
    per request and in order, the OutputParams JSON bytes or a RuntimeError instance for a request Prove would panic on."""
    L = _lib.load()
    items = [json.loads(p) if isinstance(p, (bytes, bytearray, str)) else p for p in params_list]
    s, keep = _slice(json.dumps(items).encode())
    r = L.ProveBatch(s)
    try:
        out = C.string_at(r.r0, r.r1)
    finally:
        L.Free(r.r0)
    if out[:1] != b"[":
        _raise_if_panic(out)
    res = []
    for item in json.loads(out):
        if isinstance(item, dict) and "proof" in item:
            res.append(json.dumps(item).encode())
        else:
            res.append(RuntimeError(item if isinstance(item, str) else "prover panicked with error value " + json.dumps(item)))
    return res


@dataclass
class InputVerifyParams:   # libraries/verifier/impl/verify_impl.go:18-22
    cipher: str
    proof: bytes
    public_signals: bytes   # ciphertext(64) | nonce(12) | counter(4: LE for chacha20, BE for AES) | plaintext(64)

    def to_json(self) -> bytes:
        return json.dumps({"cipher": self.cipher, "proof": list(self.proof), "publicSignals": list(self.public_signals)}).encode()


def InitVerifier(algorithm_id: int, verifying_key: bytes) -> bool:
    """Hands one of the reference's embedded verifying keys (verify_impl.go:26-62) to the library."""
    L = _lib.load()
    if not 0 <= int(algorithm_id) <= 255:
        return False
    s1, k1 = _slice(verifying_key)
    return bool(L.InitVerifier(int(algorithm_id), s1))


def Verify(params: bytes) -> bool:
    """libraries/verifier/libverify.go:14-17 — every failure is False."""
    L = _lib.load()
    s, keep = _slice(bytes(params))
    return bool(L.Verify(s))


def Setup(r1cs: bytes, trapdoor=None, device: int = 0):
    """keygen.go:359-435 (groth16.Setup + WriteTo) on the GPU: -> (pk bytes, vk bytes) in the layouts InitAlgorithm /
    InitVerifier read. trapdoor: six ints (tau, alpha, beta, gamma, delta, sigma) for reproducible keys, or None to draw
    the toxic waste from the OS CSPRNG inside the library (what gnark does)."""
    L = _lib.load()
    td = None
    if trapdoor is not None:
        if len(trapdoor) != 6:
            raise ValueError("trapdoor = (tau, alpha, beta, gamma, delta, sigma)")
        td = np.frombuffer(b"".join(int(x).to_bytes(32, "big") for x in trapdoor), dtype=np.uint8).copy()
    pk, vk = C.c_void_p(), C.c_void_p()
    npk, nvk = C.c_size_t(0), C.c_size_t(0)
    _check(L.g16_setup(r1cs, len(r1cs), _p8(td) if td is not None else None, device, C.byref(pk), C.byref(npk), C.byref(vk), C.byref(nvk)))
    try:
        return C.string_at(pk, npk.value), C.string_at(vk, nvk.value)
    finally:
        L.Free(pk)
        L.Free(vk)


class Groth16Context:
    """One (pk, r1cs) pair resident on one GPU — or, with `devices=[...]`, on every GPU of the list (g16_init_multi):
    batches are then sharded request i -> devices[i mod G], proved concurrently and gathered in input order.
    The inner seam under gnark's groth16.Prove."""

    def __init__(self, pk: bytes, r1cs: bytes, device: int = 0, devices=None):
        self._L = _lib.load()
        self._h = C.c_void_p()
        if devices is None:
            _check(self._L.g16_init(pk, len(pk), r1cs, len(r1cs), device, C.byref(self._h)))
            self.devices = [device]
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            _check(self._L.g16_init_multi(pk, len(pk), r1cs, len(r1cs), arr, len(devices), C.byref(self._h)))
            self.devices = [int(d) for d in devices]
        info = np.zeros(16, dtype=np.uint64)
        _check(self._L.g16_info(self._h, _p64(info)))
        (self.n, self.nA, self.nB, self.nZ, self.nK, self.nB2, self.nb_wires, self.nb_public, self.nb_secret,
         self.nb_constraints, self.nb_instructions, self.nb_levels, self.nb_commitments, self.proof_bytes, self.device,
         self.supported) = [int(x) for x in info]

    def close(self):
        if self._h:
            self._L.g16_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- gnark-shaped: witness (nbPublic-1+nbSecret Montgomery Fr, [n,4] u64) -> proof bytes
    def prove_witness(self, witness: np.ndarray, rs: bytes | None = None, detail: bool = False):
        w = np.ascontiguousarray(witness, dtype=np.uint64).reshape(-1, 4)
        proof = np.zeros(self.proof_bytes, dtype=np.uint8)
        plen = C.c_size_t(0)
        rsb = np.frombuffer(rs, dtype=np.uint8).copy() if rs is not None else None
        if not detail:
            _check(self._L.g16_prove_witness(self._h, _p64(w), len(w), _p8(rsb) if rsb is not None else None, _p8(proof),
                                             C.byref(plen)))
            return proof[:plen.value].tobytes()
        g1 = np.zeros((4, 8), dtype=np.uint64)
        g2 = np.zeros(16, dtype=np.uint64)
        h = np.zeros((self.n, 4), dtype=np.uint64)
        _check(self._L.g16_prove_witness_detail(self._h, _p64(w), len(w), _p8(rsb) if rsb is not None else None,
                                                _p8(proof), C.byref(plen), _p64(g1), _p64(g2), _p64(h)))
        return proof[:plen.value].tobytes(), dict(msmA=g1[0], msmB1=g1[1], msmK=g1[2], msmZ=g1[3], msmB2=g2, h=h)

    # ---- library-shaped batch (cipher "chacha20")
    @staticmethod
    def _pack(keys, nonces, counters, inputs, rs):
        n = len(counters)
        k = np.frombuffer(b"".join(keys), dtype=np.uint8).copy()
        no = np.frombuffer(b"".join(nonces), dtype=np.uint8).copy()
        i = np.frombuffer(b"".join(inputs), dtype=np.uint8).copy()
        c = np.asarray(counters, dtype=np.uint32).copy()
        if k.size != 32 * n:
            raise ValueError(f"key length must be 32: {k.size // max(n, 1)}")
        if no.size != 12 * n:
            raise ValueError(f"nonce length must be 12: {no.size // max(n, 1)}")
        if i.size != 64 * n:
            raise ValueError(f"plaintext length must be 64: {i.size // max(n, 1)}")
        r = None
        if rs is not None:
            r = np.frombuffer(b"".join(rs) if not isinstance(rs, (bytes, bytearray)) else bytes(rs), dtype=np.uint8).copy()
            if r.size != 64 * n:
                raise ValueError("rs must hold 64 bytes per proof")
        return n, k, no, c, i, r

    def prove_chacha_batch(self, keys, nonces, counters, inputs, rs=None):
        """-> (list of 164-byte proofs, list of 64-byte ciphertexts)"""
        n, k, no, c, i, r = self._pack(keys, nonces, counters, inputs, rs)
        proofs = np.zeros(n * self.proof_bytes, dtype=np.uint8)
        cts = np.zeros(n * 64, dtype=np.uint8)
        _check(self._L.g16_prove_chacha_batch(self._h, n, _p8(k), _p8(no), c.ctypes.data_as(u32p), _p8(i),
                                              _p8(r) if r is not None else None, _p8(proofs), _p8(cts)))
        pb = self.proof_bytes
        return [proofs[j * pb:(j + 1) * pb].tobytes() for j in range(n)], [cts[j * 64:(j + 1) * 64].tobytes() for j in range(n)]

    def prove_aes_batch(self, keys, nonces, counters, inputs, rsm=None):
        """cipher "aes-128-ctr" / "aes-256-ctr": -> (list of 196-byte proofs, list of 64-byte ciphertexts).
        rsm: per proof r | s | mask (3 x 32-byte big-endian) or None for fresh randomness."""
        n = len(counters)
        key_len = len(keys[0])
        if key_len not in (16, 32):
            raise ValueError(f"key length must be 16 or 32: {key_len}")
        k = np.frombuffer(b"".join(keys), dtype=np.uint8).copy()
        no = np.frombuffer(b"".join(nonces), dtype=np.uint8).copy()
        i = np.frombuffer(b"".join(inputs), dtype=np.uint8).copy()
        c = np.asarray(counters, dtype=np.uint32).copy()
        if no.size != 12 * n:
            raise ValueError(f"nonce length must be 12: {no.size // max(n, 1)}")
        if i.size != 64 * n:
            raise ValueError(f"plaintext length must be 64: {i.size // max(n, 1)}")
        r = None
        if rsm is not None:
            r = np.frombuffer(b"".join(rsm), dtype=np.uint8).copy()
            if r.size != 96 * n:
                raise ValueError("rsm must hold 96 bytes per proof")
        proofs = np.zeros(n * self.proof_bytes, dtype=np.uint8)
        cts = np.zeros(n * 64, dtype=np.uint8)
        _check(self._L.g16_prove_aes_batch(self._h, n, _p8(k), key_len, _p8(no), c.ctypes.data_as(u32p), _p8(i),
                                           _p8(r) if r is not None else None, _p8(proofs), _p8(cts)))
        pb = self.proof_bytes
        return [proofs[j * pb:(j + 1) * pb].tobytes() for j in range(n)], [cts[j * 64:(j + 1) * 64].tobytes() for j in range(n)]

    # phase-split variant for benchmarks (arrays already packed by the caller)
    def stage(self, k, no, c, i, r):
        _check(self._L.g16_chacha_batch_stage(self._h, len(c), _p8(k), _p8(no), c.ctypes.data_as(u32p), _p8(i),
                                              _p8(r) if r is not None else None))

    def stage_aes(self, k, key_len, no, c, i, r):
        _check(self._L.g16_aes_batch_stage(self._h, len(c), _p8(k), key_len, _p8(no), c.ctypes.data_as(u32p), _p8(i),
                                           _p8(r) if r is not None else None))

    def run(self) -> float:
        ms = C.c_float(0)
        _check(self._L.g16_chacha_batch_run(self._h, C.byref(ms)))
        return ms.value

    def fetch(self, proofs: np.ndarray, cts: np.ndarray):
        _check(self._L.g16_chacha_batch_fetch(self._h, _p8(proofs), _p8(cts)))

    def batch_status(self, n: int) -> np.ndarray:
        """Per-request status words of the last batch (0 = proved; bit 0 unsatisfied constraint, bit 1 division by zero)."""
        st = np.zeros(n, dtype=np.uint32)
        _check(self._L.g16_last_batch_status(self._h, st.ctypes.data_as(u32p), n))
        return st

    def set_schedule(self, pipeline: bool, sub_batch: int = 0):
        """pipeline=True: sub-batches alternate between two streams (throughput schedule, total time only);
        False: one main stream with per-stage timers. Results are identical."""
        _check(self._L.g16_set_schedule(self._h, int(bool(pipeline)), int(sub_batch)))

    def stage_ms(self) -> dict:
        ms = np.zeros(8, dtype=np.float32)
        _check(self._L.g16_last_stage_ms(self._h, ms.ctypes.data_as(f32p)))
        names = ["solve", "compute_h", "msm_sort", "msm_accumulate", "msm_reduce", "assemble", "total", "launches"]
        return {k: float(v) for k, v in zip(names, ms)}

    def counters(self) -> dict:
        c = np.zeros(8, dtype=np.uint64)
        _check(self._L.g16_last_counters(self._h, _p64(c)))
        names = ["g1_madds", "g2_madds", "g1_acc_launches", "g2_acc_launches", "launches", "proofs", "g1_madds_main_stream"]
        out = {k: int(v) for k, v in zip(names, c)}
        out["sub_batch"] = int(c[7]) & 0xFFFFFFFF
        out["pipelined"] = bool((int(c[7]) >> 32) & 1)
        out["eval_basis_z"] = bool((int(c[7]) >> 33) & 1)   # Z query over the evaluation-basis tables (4 transforms, no H)
        e = np.zeros(16, dtype=np.uint64)
        _check(self._L.g16_last_counters_ex(self._h, _p64(e)))
        out["z_sorted_slots"] = int(e[8])     # Z-query entries + batch-affine padding
        out["z_batch_affine_levels"] = int(e[9])
        out["z_buckets"] = int(e[10])
        out["z_xyzz_entries"] = int(e[13])    # what the XYZZ accumulation of the Z query walks: group sums + direct leftovers
        out["bitq_state"] = int(e[11]) & 0xFF  # combination tables of the wire queries: 0 learning, 1 in use, 2 off
        out["bitq_live"] = bool((int(e[11]) >> 8) & 1)
        return out

    # ---- stage-level
    def solve(self, witness: np.ndarray, batch: int = 1, masks=None):
        """masks: list of ints (one per witness) for circuits with a hints.Randomize wire"""
        w = np.ascontiguousarray(witness, dtype=np.uint64).reshape(batch, -1, 4)
        nw = w.shape[1]
        W = np.zeros((batch, self.nb_wires, 4), dtype=np.uint64)
        A = np.zeros((batch, self.nb_constraints, 4), dtype=np.uint64)
        B = np.zeros_like(A)
        Cc = np.zeros_like(A)
        mb = None
        if masks is not None:
            mb = np.frombuffer(b"".join(int(m).to_bytes(32, "big") for m in masks), dtype=np.uint8).copy()
        _check(self._L.g16_solve_ex(self._h, _p64(w), nw, batch, _p8(mb) if mb is not None else None, _p64(W), _p64(A),
                                    _p64(B), _p64(Cc)))
        return W, A, B, Cc

    def compute_h(self, a, b, c) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
        c = np.ascontiguousarray(c, dtype=np.uint64)
        h = np.zeros((self.n, 4), dtype=np.uint64)
        _check(self._L.g16_compute_h(self._h, _p64(a), _p64(b), _p64(c), _p64(h)))
        return h


# ---------------------------------------------------------------------------------------------- stage-level free functions
def field_op(field: int, op: str, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
    code = {"add": 0, "sub": 1, "mul": 2, "inv": 3, "sqr": 4, "neg": 5, "to_mont": 6, "from_mont": 7}[op]
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(a)
    bp = None
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        bp = _p64(b)
    _check(_lib.load().g16_field_op(field, code, _p64(a), bp, _p64(out), len(a)))
    return out


def group_op(group: int, op: str, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
    code = {"add": 0, "mul": 1, "dbl": 2, "add_xyzz": 3, "add_team": 4, "dbl_add_team": 5}[op]
    w = 8 if group == 1 else 16
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, w)
    out = np.empty_like(a)
    bp = None
    if b is not None:
        b = np.ascontiguousarray(b, dtype=np.uint64)
        bp = _p64(b)
    _check(_lib.load().g16_group_op(group, code, _p64(a), bp, _p64(out), len(a)))
    return out


def decompress(group: int, raw: bytes) -> np.ndarray:
    sz = 32 if group == 1 else 64
    n = len(raw) // sz
    buf = np.frombuffer(raw, dtype=np.uint8).copy()
    out = np.zeros((n, 8 if group == 1 else 16), dtype=np.uint64)
    _check(_lib.load().g16_decompress(group, _p8(buf), _p64(out), n))
    return out


def aes_witness(keys, nonces, counters, inputs, with_witness: bool = True):
    """provers.go:172-227 witness assignment on the device: -> (ciphertexts [n] bytes, witness [n, 142+key_len, 4] u64
    Montgomery in solver order ONE | Nonce | Counter | Plaintext | Ciphertext | Key)."""
    n = len(counters)
    key_len = len(keys[0])
    k = np.frombuffer(b"".join(keys), dtype=np.uint8).copy()
    no = np.frombuffer(b"".join(nonces), dtype=np.uint8).copy()
    i = np.frombuffer(b"".join(inputs), dtype=np.uint8).copy()
    c = np.asarray(counters, dtype=np.uint32).copy()
    cts = np.zeros(n * 64, dtype=np.uint8)
    wit = np.zeros((n, 142 + key_len, 4), dtype=np.uint64) if with_witness else None
    _check(_lib.load().g16_aes_witness(_p8(k), key_len, _p8(no), c.ctypes.data_as(u32p), _p8(i), n, _p8(cts),
                                       _p64(wit) if with_witness else None))
    return [cts[j * 64:(j + 1) * 64].tobytes() for j in range(n)], wit


def bsb22_challenge(commitments: np.ndarray) -> np.ndarray:
    """gnark prove.go:84-108 commitment hash: affine G1 points [n, 8] u64 Montgomery -> challenges [n, 4] u64 Montgomery"""
    pts = np.ascontiguousarray(commitments, dtype=np.uint64).reshape(-1, 8)
    out = np.zeros((len(pts), 4), dtype=np.uint64)
    _check(_lib.load().g16_bsb22_challenge(_p64(pts), len(pts), _p64(out)))
    return out


class Groth16Verifier:
    """One verifying key resident on one GPU: batched groth16.Verify (libraries/verifier/impl/verifiers.go:87-99,133-145)."""

    def __init__(self, vk: bytes, device: int = 0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        _check(self._L.g16_verify_init(vk, len(vk), device, C.byref(self._h)))
        info = np.zeros(4, dtype=np.uint64)
        _check(self._L.g16_verify_info(self._h, _p64(info)))
        self.n_public, self.n_commitments, self.proof_bytes, self.nK = [int(x) for x in info]
        self.last_ms = 0.0

    def close(self):
        if self._h:
            self._L.g16_verify_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def verify_batch(self, proofs, public_inputs) -> np.ndarray:
        """proofs: list of proof byte strings. public_inputs: per proof either a list of ints (sent as 32-byte big-endian
        values) or, as one array [n, n_public, 4] of u64, gnark's in-memory Montgomery elements. -> bool array."""
        n = len(proofs)
        if any(len(p) != self.proof_bytes for p in proofs):
            raise ValueError(f"every proof must be {self.proof_bytes} bytes")
        pr = np.frombuffer(b"".join(proofs), dtype=np.uint8).copy()
        if isinstance(public_inputs, np.ndarray):
            pub = np.ascontiguousarray(public_inputs, dtype=np.uint64).reshape(n, self.n_public, 4)
            fmt = 0
        else:
            if any(len(v) != self.n_public for v in public_inputs):
                raise ValueError(f"every proof needs {self.n_public} public inputs")
            pub = np.frombuffer(b"".join(int(x).to_bytes(32, "big") for v in public_inputs for x in v), dtype=np.uint8).copy()
            fmt = 1
        ok = np.zeros(n, dtype=np.uint8)
        ms = C.c_float(0)
        _check(self._L.g16_verify_batch(self._h, n, _p8(pr), pub.ctypes.data_as(C.c_void_p), fmt, _p8(ok), C.byref(ms)))
        self.last_ms = ms.value
        return ok.astype(bool)

    def verify(self, proof: bytes, public_inputs) -> bool:
        return bool(self.verify_batch([proof], [public_inputs])[0])


def g2_subgroup_check(g2s: np.ndarray) -> np.ndarray:
    """gnark-crypto G2Affine.IsInSubGroup for affine Montgomery twist points ([n, 16] u64) -> bool array."""
    Q = np.ascontiguousarray(g2s, dtype=np.uint64).reshape(-1, 16)
    ok = np.zeros(len(Q), dtype=np.uint8)
    _check(_lib.load().g16_g2_subgroup_check(_p64(Q), len(Q), _p8(ok)))
    return ok.astype(bool)


def pairing_check(g1s: np.ndarray, g2s: np.ndarray, pairs_per_check: int | None = None) -> np.ndarray:
    """prod_j e(P_j, Q_j) == 1 for every group of `pairs_per_check` consecutive pairs (default: one check over all pairs).
    Points affine Montgomery ([n, 8] / [n, 16] u64). Returns a bool array, one entry per check."""
    P = np.ascontiguousarray(g1s, dtype=np.uint64).reshape(-1, 8)
    Q = np.ascontiguousarray(g2s, dtype=np.uint64).reshape(-1, 16)
    if len(P) != len(Q) or not len(P):
        raise ValueError("need as many G1 as G2 points (at least one)")
    ppc = len(P) if pairs_per_check is None else int(pairs_per_check)
    if ppc < 1 or len(P) % ppc:
        raise ValueError("pairs_per_check must divide the number of pairs")
    ok = np.zeros(len(P) // ppc, dtype=np.uint8)
    _check(_lib.load().g16_pairing_check(_p64(P), _p64(Q), ppc, len(ok), _p8(ok)))
    return ok.astype(bool)


def msm(group: int, points: np.ndarray, scalars: np.ndarray, scalars_mont: bool = False, window: int = 0):
    """-> (affine result, [total, accumulate, sort, reduce] ms)"""
    w = 8 if group == 1 else 16
    pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, w)
    sc = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    assert len(pts) == len(sc)
    out = np.zeros(w, dtype=np.uint64)
    ms = np.zeros(4, dtype=np.float32)
    _check(_lib.load().g16_msm(group, _p64(pts), _p64(sc), int(scalars_mont), len(pts), window, _p64(out),
                               ms.ctypes.data_as(f32p)))
    return out, ms


class MsmPlan:
    """Device-resident MSM (points uploaded once) for the standalone sweeps of BASELINE config 5."""

    def __init__(self, group: int, points: np.ndarray, window: int = 0, device: int = 0, precompute: bool = False):
        """precompute=True: the bases are an SRS used by many MSMs — tabulate 2^(c w) P_i once (one bucket set per MSM
        instead of one per window, no doubling chain at the end: the mode the prover context uses for the pk queries)."""
        self._L = _lib.load()
        self.group = group
        w = 8 if group == 1 else 16
        pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, w)
        self.n = len(pts)
        self._h = C.c_void_p()
        _check(self._L.g16_msm_plan_create(group, _p64(pts), self.n, 0 if precompute else window, device, C.byref(self._h)))
        if precompute:
            _check(self._L.g16_msm_plan_precompute(self._h, window))

    def set_scalars(self, scalars: np.ndarray, mont: bool = False):
        sc = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        assert len(sc) == self.n
        _check(self._L.g16_msm_plan_set_scalars(self._h, _p64(sc), int(mont)))

    def run(self):
        out = np.zeros(8 if self.group == 1 else 16, dtype=np.uint64)
        ms = np.zeros(4, dtype=np.float32)
        _check(self._L.g16_msm_plan_run(self._h, _p64(out), ms.ctypes.data_as(f32p)))
        return out, ms

    def close(self):
        if self._h:
            self._L.g16_msm_plan_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ntt(data: np.ndarray, inverse: bool = False, coset: bool = False):
    d = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 4).copy()
    ms = C.c_float(0)
    _check(_lib.load().g16_ntt(_p64(d), len(d), int(inverse), int(coset), C.byref(ms)))
    return d, ms.value


def ntt_bench(n: int, batch: int, iters: int):
    """-> (ms per transform of the whole batch, mismatches after forward+inverse round trips)"""
    ms = C.c_float(0)
    chk = C.c_uint64(0)
    _check(_lib.load().g16_ntt_bench(n, batch, iters, C.byref(ms), C.byref(chk)))
    return ms.value, chk.value


def imad_peak():
    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    _check(_lib.load().g16_imad_peak(C.byref(a), C.byref(b), C.byref(c)))
    d = C.c_double(0)
    _check(_lib.load().g16_imad_chain_rate(C.byref(d)))
    return {"imad_per_s": a.value, "imad_wide_per_s": b.value, "modmul_per_s": c.value, "imad_wide_carry_per_s": d.value}
