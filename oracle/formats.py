"""ORACLE — TEST INFRASTRUCTURE ONLY. Not imported by the product package.

Python decoders for the gnark v0.11.0 binary artefacts the reference feeds to its prover
(`libraries/prover/impl/prove_impl.go:86-91,102-107`: ProvingKey.ReadFrom / NewCS(...).ReadFrom) and verifier
(`libraries/verifier/impl/verify_impl.go:36-58`). gnark itself is not on this box; the layouts are the ones decoded
from the reference's own fixtures in SURVEY.md Appendices A-D. The product has its own, independently written C++
decoders (gnark_symmetric_crypto_b200/csrc/r1cs_parse.cpp); tests diff the two.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
P_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47

HINT_NBITS = 4115454955
HINT_COUNT = 2138922168
HINT_RANDOMIZE = 1774611027
HINT_BSB22 = 4156202267

TAG_HINT = 5309735
TAG_R1C = 5309736
TAG_LOOKUP = 5309741
TAG_COMMIT = 5309742


# ----------------------------------------------------------------------------- intcomp streams (Appendix D)
def _unzigzag(v: np.ndarray, bits: int) -> np.ndarray:
    one = v.dtype.type(1)
    return (v >> one) ^ (-(v & one)).astype(v.dtype)


def _unpack_bits(words: np.ndarray, bitlen: int, count: int, wbits: int) -> np.ndarray:
    """`count` values of `bitlen` bits packed LSB-first across little-endian words of `wbits` bits."""
    dt = np.uint32 if wbits == 32 else np.uint64
    if bitlen == 0:
        return np.zeros(count, dtype=dt)
    # go through python ints: robust and the blocks are tiny
    big = 0
    for i, w in enumerate(words.tolist()):
        big |= int(w) << (wbits * i)
    mask = (1 << bitlen) - 1
    return np.array([(big >> (bitlen * k)) & mask for k in range(count)], dtype=dt)


def decode_stream(words: np.ndarray, wbits: int) -> np.ndarray:
    """Decode one `ronanh/intcomp` delta-bitpacked stream (u32 or u64 flavour) -> array of values."""
    dt = np.uint32 if wbits == 32 else np.uint64
    sub = wbits                      # values per sub-block: 32 or 64
    group = 4 * sub                  # values per group: 128 or 256
    n = len(words)
    out = []
    pos = 0
    mod = 1 << wbits
    if n == 0:
        return np.zeros(0, dtype=dt)
    body_end = n - 1                 # last word is the trailer
    # ---- optional bit-packed block
    if wbits == 32:
        first_nints = int(words[0])
    else:
        first_nints = int(words[0]) & 0xFFFFFFFF
    if first_nints >= group and first_nints % group == 0:
        if wbits == 32:
            nints, nwords, init = int(words[0]), int(words[1]), int(words[2])
            p = 3
        else:
            nints, nwords, init = int(words[0]) & 0xFFFFFFFF, int(words[0]) >> 32, int(words[1])
            p = 2
        blk_end = nwords
        prev = init
        vals = []
        for _g in range(nints // group):
            hdr = int(words[p]) & 0xFFFFFFFF
            p += 1
            for sb in range(4):
                b = (hdr >> (8 * (3 - sb))) & 0xFF
                zz, bitlen = b >> 7, b & 0x7F
                raw = _unpack_bits(words[p:p + bitlen], bitlen, sub, wbits)
                p += bitlen
                for v in raw.tolist():
                    d = ((v >> 1) ^ -(v & 1)) if zz else v
                    prev = (prev + d) % mod
                    vals.append(prev)
        assert p == blk_end, (p, blk_end)
        out.extend(vals)
        pos = blk_end
    # ---- optional varbyte block
    if pos < body_end:
        if wbits == 32:
            nints, nwords = int(words[pos]), int(words[pos + 1])
            p = pos + 2
        else:
            nints, nwords = int(words[pos]) & 0xFFFFFFFF, int(words[pos]) >> 32
            p = pos + 1
        payload = words[p:pos + nwords]
        by = payload.astype(">u4" if wbits == 32 else ">u8").tobytes()   # MSB-first within each word
        prev = 0
        k = 0
        for _ in range(nints):
            v = 0
            shift = 0
            while True:
                c = by[k]
                k += 1
                v |= (c & 0x7F) << shift
                shift += 7
                if not (c & 0x80):
                    break
            prev = (prev + v) % mod
            out.append(prev)
        pos += nwords
    assert pos == body_end, (pos, body_end, n)
    return np.array(out, dtype=dt)


# ----------------------------------------------------------------------------- r1cs (Appendix D)
@dataclass
class R1CS:
    levels: list            # list of np.uint32 arrays (instruction indices)
    bp_id: np.ndarray       # per instruction
    cons_off: np.ndarray
    wire_off: np.ndarray
    start: np.ndarray       # u64, index into calldata
    calldata: np.ndarray    # u32
    body: dict
    coeffs: np.ndarray      # [ncoef, 4] u64 LE limbs, Montgomery
    bp_kind: list = field(default_factory=list)        # per blueprint: 0 R1C, 1 hint, 2 lookup
    lookup_entries: dict = field(default_factory=dict)  # blueprint id -> list of (cid, wid) pairs per entry
    n_public: int = 0       # incl. the ONE wire
    n_secret: int = 0
    n_internal: int = 0
    n_constraints: int = 0
    commitments: list = field(default_factory=list)

    @property
    def n_wires(self):
        return self.n_public + self.n_secret + self.n_internal

    @property
    def n_instr(self):
        return len(self.bp_id)


def _leb128_all(buf: bytes, count: int) -> np.ndarray:
    out = np.empty(count, dtype=np.uint32)
    k = 0
    for i in range(count):
        v = 0
        shift = 0
        while True:
            c = buf[k]
            k += 1
            v |= (c & 0x7F) << shift
            shift += 7
            if not (c & 0x80):
                break
        out[i] = v
    assert k == len(buf), (k, len(buf))
    return out


def parse_r1cs(data: bytes) -> R1CS:
    import cbor2

    total, z0, minor, z1 = struct.unpack_from("<4Q", data, 0)
    assert total + 32 == len(data), (total, len(data))
    lv_len, ins_len, cd_len, body_len = struct.unpack_from("<4Q", data, 32)
    off = 64
    # levels
    sec = data[off:off + lv_len]
    off += lv_len
    (nlev,) = struct.unpack_from("<Q", sec, 0)
    p = 8
    levels = []
    for _ in range(nlev):
        (nw,) = struct.unpack_from("<Q", sec, p)
        p += 8
        words = np.frombuffer(sec, dtype="<u4", count=nw, offset=p)
        p += 4 * nw
        levels.append(decode_stream(words, 32))
    assert p == lv_len
    # instructions: 3 u32 columns + 1 u64 column
    sec = data[off:off + ins_len]
    off += ins_len
    p = 0
    cols = []
    for c in range(4):
        (nw,) = struct.unpack_from("<Q", sec, p)
        p += 8
        if c < 3:
            words = np.frombuffer(sec, dtype="<u4", count=nw, offset=p)
            p += 4 * nw
            cols.append(decode_stream(words, 32))
        else:
            words = np.frombuffer(sec, dtype="<u8", count=nw, offset=p)
            p += 8 * nw
            cols.append(decode_stream(words, 64))
    assert p == ins_len
    # calldata
    sec = data[off:off + cd_len]
    off += cd_len
    (ncd,) = struct.unpack_from("<Q", sec, 0)
    calldata = _leb128_all(sec[8:], ncd)
    # body
    body = cbor2.loads(data[off:off + body_len])
    off += body_len
    (ncoef,) = struct.unpack_from("<Q", data, off)
    off += 8
    coeffs = np.frombuffer(data, dtype="<u8", count=4 * ncoef, offset=off).reshape(ncoef, 4).copy()
    off += 32 * ncoef
    assert off == len(data), (off, len(data))

    r = R1CS(levels=levels, bp_id=cols[0], cons_off=cols[1], wire_off=cols[2], start=cols[3], calldata=calldata,
             body=body, coeffs=coeffs)
    for i, bp in enumerate(body["Blueprints"]):
        tag = bp.tag
        if tag == TAG_R1C:
            r.bp_kind.append(0)
        elif tag == TAG_HINT:
            r.bp_kind.append(1)
        elif tag == TAG_LOOKUP:
            r.bp_kind.append(2)
            ec = list(bp.value["EntriesCalldata"])
            ents = []
            q = 0
            while q < len(ec):
                nt = ec[q]
                q += 1
                ents.append([(ec[q + 2 * t], ec[q + 2 * t + 1]) for t in range(nt)])
                q += 2 * nt
            r.lookup_entries[i] = ents
        else:
            raise ValueError(f"unknown blueprint tag {tag}")
    r.n_public = len(body["Public"])
    r.n_secret = len(body["Secret"])
    r.n_internal = int(body["NbInternalVariables"])
    r.n_constraints = int(body["NbConstraints"])
    ci = body.get("CommitmentInfo")
    if ci is not None:
        val = ci.value if hasattr(ci, "value") else ci
        r.commitments = list(val) if val else []
    return r


# ----------------------------------------------------------------------------- pk / vk headers (Appendices A, B)
@dataclass
class PKLayout:
    n: int
    hdr_fr: list                     # 5 canonical ints: n^-1, omega, omega^-1, g, g^-1
    off_alpha: int
    counts: dict                     # A,B,Z,K,B2 -> count
    offs: dict                       # A,B,Z,K,B2,beta2 -> byte offset of first point
    nb_wires: int
    nb_inf_a: int
    nb_inf_b: int
    inf_a: np.ndarray
    inf_b: np.ndarray
    n_commit_keys: int


def parse_pk_layout(data: bytes) -> PKLayout:
    (n,) = struct.unpack_from(">Q", data, 0)
    hdr = [int.from_bytes(data[8 + 32 * i:40 + 32 * i], "big") for i in range(5)]
    off = 8 + 160 + 1
    off_alpha = off
    off += 96
    counts, offs = {}, {}
    for name in ("A", "B", "Z", "K"):
        (c,) = struct.unpack_from(">I", data, off)
        off += 4
        counts[name], offs[name] = c, off
        off += 32 * c
    offs["beta2"] = off
    off += 128
    (c,) = struct.unpack_from(">I", data, off)
    off += 4
    counts["B2"], offs["B2"] = c, off
    off += 64 * c
    nbw, nia, nib = struct.unpack_from(">3Q", data, off)
    off += 24
    inf_a = np.frombuffer(data, dtype=np.uint8, count=nbw, offset=off).copy()
    off += nbw
    inf_b = np.frombuffer(data, dtype=np.uint8, count=nbw, offset=off).copy()
    off += nbw
    (nck,) = struct.unpack_from(">I", data, off)
    off += 4
    if nck == 0:
        assert off == len(data), (off, len(data))
    return PKLayout(n, hdr, off_alpha, counts, offs, nbw, nia, nib, inf_a, inf_b, nck)
