// See host_parse.hpp. Formats: SURVEY.md Appendix A (pk) and Appendix D (r1cs), decoded from the reference's fixtures.
#include "host_parse.hpp"
#include <cstring>

namespace g16 {

namespace {

struct Rd {
    const uint8_t* p;
    size_t len, off = 0;
    const uint8_t* take(size_t n, const char* what) {
        if (n > len - off || off > len) throw ParseError(std::string("truncated input while reading ") + what);
        const uint8_t* r = p + off;
        off += n;
        return r;
    }
    uint64_t be64(const char* w) { const uint8_t* b = take(8, w); uint64_t v = 0; for (int i = 0; i < 8; i++) v = (v << 8) | b[i]; return v; }
    uint32_t be32(const char* w) { const uint8_t* b = take(4, w); uint32_t v = 0; for (int i = 0; i < 4; i++) v = (v << 8) | b[i]; return v; }
    uint64_t le64(const char* w) { const uint8_t* b = take(8, w); uint64_t v = 0; for (int i = 7; i >= 0; i--) v = (v << 8) | b[i]; return v; }
};

}  // namespace

PkFile parse_pk(const uint8_t* data, size_t len) {
    Rd r{data, len};
    PkFile pk;
    pk.n = r.be64("domain cardinality");
    if (pk.n == 0 || (pk.n & (pk.n - 1)) || pk.n > (1ull << 28)) throw ParseError("pk: domain cardinality is not a power of two");
    for (int i = 0; i < 5; i++) memcpy(pk.fr_hdr[i], r.take(32, "domain header"), 32);
    r.take(1, "withPrecompute");
    pk.g1_abd = r.take(96, "G1 alpha/beta/delta");
    pk.nA = r.be32("len(G1.A)"); pk.A = r.take((size_t)pk.nA * 32, "G1.A");
    pk.nB = r.be32("len(G1.B)"); pk.B = r.take((size_t)pk.nB * 32, "G1.B");
    pk.nZ = r.be32("len(G1.Z)"); pk.Z = r.take((size_t)pk.nZ * 32, "G1.Z");
    pk.nK = r.be32("len(G1.K)"); pk.K = r.take((size_t)pk.nK * 32, "G1.K");
    pk.g2_bd = r.take(128, "G2 beta/delta");
    pk.nB2 = r.be32("len(G2.B)"); pk.B2 = r.take((size_t)pk.nB2 * 64, "G2.B");
    pk.nb_wires = r.be64("nbWires");
    pk.nb_inf_a = r.be64("NbInfinityA");
    pk.nb_inf_b = r.be64("NbInfinityB");
    if (pk.nb_wires > (1ull << 31)) throw ParseError("pk: implausible nbWires");
    pk.inf_a = r.take(pk.nb_wires, "InfinityA");
    pk.inf_b = r.take(pk.nb_wires, "InfinityB");
    pk.n_commit_keys = r.be32("len(CommitmentKeys)");
    if (pk.n_commit_keys > 16) throw ParseError("pk: implausible number of commitment keys");
    for (uint32_t k = 0; k < pk.n_commit_keys; k++) {
        PkFile::Ped ped;
        ped.n_basis = r.be32("len(Pedersen.Basis)");
        ped.basis = r.take((size_t)ped.n_basis * 32, "Pedersen.Basis");
        ped.n_sigma = r.be32("len(Pedersen.BasisExpSigma)");
        ped.basis_sigma = r.take((size_t)ped.n_sigma * 32, "Pedersen.BasisExpSigma");
        if (ped.n_basis != ped.n_sigma) throw ParseError("pk: Pedersen basis sizes differ");
        pk.ped.push_back(ped);
    }
    if (r.off != len) throw ParseError("pk: trailing bytes");
    if (pk.nZ + 1 != pk.n) throw ParseError("pk: len(G1.Z) != n-1");
    if (pk.nB != pk.nB2) throw ParseError("pk: len(G1.B) != len(G2.B)");
    if (pk.nA + pk.nb_inf_a != pk.nb_wires || pk.nB + pk.nb_inf_b != pk.nb_wires) throw ParseError("pk: infinity counts inconsistent");
    return pk;
}

VkFile parse_vk(const uint8_t* data, size_t len) {
    Rd r{data, len};
    VkFile vk;
    vk.alpha = r.take(32, "G1.Alpha");
    vk.beta1 = r.take(32, "G1.Beta");
    vk.beta2 = r.take(64, "G2.Beta");
    vk.gamma2 = r.take(64, "G2.Gamma");
    vk.delta1 = r.take(32, "G1.Delta");
    vk.delta2 = r.take(64, "G2.Delta");
    vk.nK = r.be32("len(G1.K)");
    if (vk.nK == 0 || vk.nK > (1u << 24)) throw ParseError("vk: implausible len(G1.K)");
    vk.K = r.take((size_t)vk.nK * 32, "G1.K");
    uint32_t outer = r.be32("len(PublicAndCommitmentCommitted)");
    if (outer > 16) throw ParseError("vk: implausible number of commitments");
    for (uint32_t i = 0; i < outer; i++) {
        uint32_t inner = r.be32("len(PublicAndCommitmentCommitted[i])");
        if (inner > vk.nK) throw ParseError("vk: implausible committed list");
        std::vector<uint64_t> v(inner);
        for (uint32_t k = 0; k < inner; k++) v[k] = r.be64("committed index");
        vk.public_and_commitment_committed.push_back(v);
    }
    uint32_t nkeys = r.be32("len(CommitmentKeys)");
    if (nkeys != outer) throw ParseError("vk: commitment keys do not match the commitment list");
    for (uint32_t i = 0; i < nkeys; i++) {
        VkFile::PedVk k;
        k.g = r.take(64, "Pedersen vk G");
        k.g_root_sigma_neg = r.take(64, "Pedersen vk GRootSigmaNeg");
        vk.ped.push_back(k);
    }
    if (r.off != len) throw ParseError("vk: trailing bytes");
    return vk;
}

// ------------------------------------------------------------------------------------------------ intcomp streams
namespace {

template <class W>
uint64_t get_bits(const W* words, size_t nwords, size_t bitoff, int bl) {
    const int WB = sizeof(W) * 8;
    size_t wi = bitoff / WB;
    int sh = (int)(bitoff % WB);
    if (wi >= nwords) throw ParseError("intcomp: bit-packed block overruns its words");
    unsigned __int128 v = words[wi];
    if (sh + bl > WB) {
        if (wi + 1 >= nwords) throw ParseError("intcomp: bit-packed block overruns its words");
        v |= (unsigned __int128)words[wi + 1] << WB;
    }
    v >>= sh;
    uint64_t mask = bl >= 64 ? ~0ull : ((1ull << bl) - 1);
    return (uint64_t)v & mask;
}

// ronanh/intcomp delta-bitpacking stream: [bitpack block]? [varbyte block]? trailer
template <class W>
void decode_stream(const W* w, size_t n, std::vector<W>& out) {
    const int WB = sizeof(W) * 8;
    const size_t SUB = WB, GROUP = 4 * SUB;
    out.clear();
    if (n == 0) return;
    const size_t end = n - 1;   // trailer word
    size_t pos = 0;
    uint64_t first_nints = WB == 32 ? (uint64_t)w[0] : ((uint64_t)w[0] & 0xFFFFFFFFull);
    if (first_nints >= GROUP && first_nints % GROUP == 0) {
        uint64_t nints, nwords;
        W prev;
        size_t p;
        if (WB == 32) {
            if (n < 4) throw ParseError("intcomp: short bitpack header");
            nints = w[0]; nwords = w[1]; prev = w[2]; p = 3;
        } else {
            if (n < 3) throw ParseError("intcomp: short bitpack header");
            nints = (uint64_t)w[0] & 0xFFFFFFFFull; nwords = (uint64_t)w[0] >> 32; prev = w[1]; p = 2;
        }
        if (nwords > end) throw ParseError("intcomp: bitpack block longer than the stream");
        out.reserve(nints + GROUP);
        for (uint64_t g = 0; g < nints / GROUP; g++) {
            if (p >= nwords) throw ParseError("intcomp: bitpack group header out of range");
            uint32_t hdr = (uint32_t)w[p++];
            for (int sb = 0; sb < 4; sb++) {
                uint32_t b = (hdr >> (8 * (3 - sb))) & 0xFF;
                int zz = b >> 7, bl = b & 0x7F;
                if (bl > WB) throw ParseError("intcomp: bit length exceeds word size");
                if (p + bl > nwords) throw ParseError("intcomp: sub-block out of range");
                for (size_t i = 0; i < SUB; i++) {
                    uint64_t v = bl ? get_bits<W>(w + p, bl, i * bl, bl) : 0;
                    W d = zz ? (W)((v >> 1) ^ (uint64_t)(-(int64_t)(v & 1))) : (W)v;
                    prev = (W)(prev + d);
                    out.push_back(prev);
                }
                p += bl;
            }
        }
        if (p != nwords) throw ParseError("intcomp: bitpack block length mismatch");
        pos = nwords;
    }
    if (pos < end) {
        uint64_t nints, nwords;
        size_t p;
        if (WB == 32) {
            if (pos + 2 > end) throw ParseError("intcomp: short varbyte header");
            nints = w[pos]; nwords = w[pos + 1]; p = pos + 2;
        } else {
            nints = (uint64_t)w[pos] & 0xFFFFFFFFull; nwords = (uint64_t)w[pos] >> 32; p = pos + 1;
        }
        if (pos + nwords > end) throw ParseError("intcomp: varbyte block longer than the stream");
        size_t nbytes = (pos + nwords - p) * sizeof(W), k = 0;
        auto byte_at = [&](size_t idx) -> uint8_t {   // most-significant byte of each word first
            size_t wi = idx / sizeof(W), bi = idx % sizeof(W);
            return (uint8_t)(w[p + wi] >> (8 * (sizeof(W) - 1 - bi)));
        };
        W prev = 0;
        for (uint64_t i = 0; i < nints; i++) {
            uint64_t v = 0;
            int shift = 0;
            for (;;) {
                if (k >= nbytes) throw ParseError("intcomp: varbyte payload exhausted");
                uint8_t c = byte_at(k++);
                v |= (uint64_t)(c & 0x7F) << shift;
                shift += 7;
                if (!(c & 0x80)) break;
                if (shift > 63) throw ParseError("intcomp: varbyte value too long");
            }
            prev = (W)(prev + (W)v);
            out.push_back(prev);
        }
        pos += nwords;
    }
    if (pos != end) throw ParseError("intcomp: trailing words");
}

// ------------------------------------------------------------------------------------------------ minimal CBOR reader
struct Cbor {
    const uint8_t* p;
    size_t len, off = 0;
    uint8_t peek() { if (off >= len) throw ParseError("cbor: truncated"); return p[off]; }
    // reads the head; returns major type, stores the argument (or 31 -> indefinite flag)
    int head(uint64_t& arg, bool& indef) {
        uint8_t b = peek();
        off++;
        int major = b >> 5, ai = b & 31;
        indef = false;
        if (ai < 24) arg = ai;
        else if (ai == 24) { need(1); arg = p[off]; off += 1; }
        else if (ai == 25) { need(2); arg = ((uint64_t)p[off] << 8) | p[off + 1]; off += 2; }
        else if (ai == 26) { need(4); arg = 0; for (int i = 0; i < 4; i++) arg = (arg << 8) | p[off + i]; off += 4; }
        else if (ai == 27) { need(8); arg = 0; for (int i = 0; i < 8; i++) arg = (arg << 8) | p[off + i]; off += 8; }
        else if (ai == 31) { indef = true; arg = 0; }
        else throw ParseError("cbor: reserved additional info");
        return major;
    }
    void need(size_t n) { if (n > len - off) throw ParseError("cbor: truncated"); }
    void skip() {
        uint64_t a; bool ind;
        int m = head(a, ind);
        switch (m) {
            case 0: case 1: return;
            case 2: case 3:
                if (ind) { while (peek() != 0xFF) skip(); off++; }
                else { need(a); off += a; }
                return;
            case 4:
                if (ind) { while (peek() != 0xFF) skip(); off++; }
                else for (uint64_t i = 0; i < a; i++) skip();
                return;
            case 5:
                if (ind) { while (peek() != 0xFF) { skip(); skip(); } off++; }
                else for (uint64_t i = 0; i < a; i++) { skip(); skip(); }
                return;
            case 6: skip(); return;
            default: return;   // simple values / floats: argument already consumed
        }
    }
    uint64_t read_uint() {
        uint64_t a; bool ind;
        int m = head(a, ind);
        if (m != 0) throw ParseError("cbor: expected unsigned integer");
        return a;
    }
    std::string read_text() {
        uint64_t a; bool ind;
        int m = head(a, ind);
        if (m != 3 || ind) throw ParseError("cbor: expected text string");
        need(a);
        std::string s((const char*)p + off, a);
        off += a;
        return s;
    }
    // positions the reader on the value of `key` inside the map starting at the current offset; false if absent
    // (reader then sits after the map)
    bool map_find(const std::string& key) {
        uint64_t a; bool ind;
        int m = head(a, ind);
        if (m != 5 || ind) throw ParseError("cbor: expected definite map");
        for (uint64_t i = 0; i < a; i++) {
            uint8_t b = peek();
            if ((b >> 5) == 3) {
                std::string k = read_text();
                if (k == key) return true;
            } else {
                skip();
            }
            skip();
        }
        return false;
    }
    uint64_t array_len() {
        uint64_t a; bool ind;
        int m = head(a, ind);
        if (m == 7 && a == 22) return 0;   // null
        if (m != 4 || ind) throw ParseError("cbor: expected definite array");
        return a;
    }
    bool is_null() { return peek() == 0xF6; }
};

uint64_t body_uint(const uint8_t* body, size_t len, const char* key) {
    Cbor c{body, len};
    if (!c.map_find(key)) throw ParseError(std::string("r1cs body: missing ") + key);
    return c.read_uint();
}
uint64_t body_array_len(const uint8_t* body, size_t len, const char* key) {
    Cbor c{body, len};
    if (!c.map_find(key)) throw ParseError(std::string("r1cs body: missing ") + key);
    return c.array_len();
}
void read_u32_array(Cbor& c, std::vector<uint32_t>& out) {
    out.clear();
    if (c.is_null()) { c.off++; return; }
    uint64_t n = c.array_len();
    out.reserve(n);
    for (uint64_t i = 0; i < n; i++) out.push_back((uint32_t)c.read_uint());
}

const uint64_t TAG_HINT = 5309735, TAG_R1C = 5309736, TAG_LOOKUP = 5309741, TAG_COMMIT = 5309742;

}  // namespace

R1csFile parse_r1cs(const uint8_t* data, size_t len) {
    Rd r{data, len};
    R1csFile f;
    uint64_t total = r.le64("outer header");
    r.le64("outer header"); r.le64("outer header"); r.le64("outer header");
    if (total + 32 != len) throw ParseError("r1cs: outer length mismatch");
    uint64_t lv_len = r.le64("block header"), ins_len = r.le64("block header"), cd_len = r.le64("block header"),
             body_len = r.le64("block header");
    // levels
    {
        Rd s{r.take(lv_len, "levels"), (size_t)lv_len};
        uint64_t nlev = s.le64("nbLevels");
        if (nlev > lv_len) throw ParseError("r1cs: implausible level count");
        f.levels.resize(nlev);
        for (uint64_t i = 0; i < nlev; i++) {
            uint64_t nw = s.le64("level word count");
            const uint8_t* b = s.take(nw * 4, "level stream");
            std::vector<uint32_t> w(nw);
            memcpy(w.data(), b, nw * 4);
            decode_stream<uint32_t>(w.data(), nw, f.levels[i]);
        }
        if (s.off != lv_len) throw ParseError("r1cs: levels section length mismatch");
    }
    // instruction columns
    {
        Rd s{r.take(ins_len, "instructions"), (size_t)ins_len};
        std::vector<uint32_t>* cols[3] = {&f.bp_id, &f.cons_off, &f.wire_off};
        for (int c = 0; c < 3; c++) {
            uint64_t nw = s.le64("column word count");
            const uint8_t* b = s.take(nw * 4, "column stream");
            std::vector<uint32_t> w(nw);
            memcpy(w.data(), b, nw * 4);
            decode_stream<uint32_t>(w.data(), nw, *cols[c]);
        }
        uint64_t nw = s.le64("column word count");
        const uint8_t* b = s.take(nw * 8, "column stream");
        std::vector<uint64_t> w(nw);
        memcpy(w.data(), b, nw * 8);
        decode_stream<uint64_t>(w.data(), nw, f.start);
        if (s.off != ins_len) throw ParseError("r1cs: instruction section length mismatch");
        if (f.bp_id.size() != f.cons_off.size() || f.bp_id.size() != f.wire_off.size() || f.bp_id.size() != f.start.size())
            throw ParseError("r1cs: instruction columns differ in length");
    }
    // calldata
    {
        Rd s{r.take(cd_len, "calldata"), (size_t)cd_len};
        uint64_t cnt = s.le64("calldata count");
        if (cnt > cd_len) throw ParseError("r1cs: implausible calldata count");
        f.calldata.resize(cnt);
        for (uint64_t i = 0; i < cnt; i++) {
            uint64_t v = 0;
            int shift = 0;
            for (;;) {
                uint8_t c = *s.take(1, "calldata varint");
                v |= (uint64_t)(c & 0x7F) << shift;
                shift += 7;
                if (!(c & 0x80)) break;
                if (shift > 35) throw ParseError("r1cs: calldata varint too long");
            }
            f.calldata[i] = (uint32_t)v;
        }
        if (s.off != cd_len) throw ParseError("r1cs: calldata section length mismatch");
    }
    // body
    const uint8_t* body = r.take(body_len, "body");
    f.n_public = body_array_len(body, body_len, "Public");
    f.n_secret = body_array_len(body, body_len, "Secret");
    f.n_internal = body_uint(body, body_len, "NbInternalVariables");
    f.n_constraints = body_uint(body, body_len, "NbConstraints");
    {
        Cbor c{body, (size_t)body_len};
        if (!c.map_find("Blueprints")) throw ParseError("r1cs body: missing Blueprints");
        uint64_t nb = c.array_len();
        for (uint64_t i = 0; i < nb; i++) {
            uint64_t tag; bool ind;
            if (c.head(tag, ind) != 6) throw ParseError("r1cs body: blueprint is not tagged");
            std::vector<uint32_t> entries;
            if (tag == TAG_R1C) { f.bp_kind.push_back(INS_R1C); c.skip(); }
            else if (tag == TAG_HINT) { f.bp_kind.push_back(INS_HINT); c.skip(); }
            else if (tag == TAG_LOOKUP) {
                f.bp_kind.push_back(INS_LOOKUP);
                size_t save = c.off;
                Cbor m{body, (size_t)body_len};
                m.off = save;
                if (!m.map_find("EntriesCalldata")) throw ParseError("r1cs body: lookup blueprint without EntriesCalldata");
                read_u32_array(m, entries);
                c.skip();
            } else throw ParseError("r1cs body: unknown blueprint tag " + std::to_string(tag));
            f.bp_lookup_entries.push_back(std::move(entries));
        }
    }
    {
        Cbor c{body, (size_t)body_len};
        if (c.map_find("CommitmentInfo")) {
            uint64_t tag; bool ind;
            int m = c.head(tag, ind);
            if (m == 6 && tag == TAG_COMMIT) {
                uint64_t nc = c.array_len();
                for (uint64_t i = 0; i < nc; i++) {
                    CommitmentInfo ci;
                    size_t item = c.off;
                    const char* keys[4] = {"CommitmentIndex", "NbPublicCommitted", "PrivateCommitted", "PublicAndCommitmentCommitted"};
                    for (int k = 0; k < 4; k++) {
                        Cbor m2{body, (size_t)body_len};
                        m2.off = item;
                        if (!m2.map_find(keys[k])) continue;
                        if (k == 0) ci.commitment_index = m2.read_uint();
                        else if (k == 1) ci.nb_public_committed = m2.read_uint();
                        else if (k == 2) read_u32_array(m2, ci.private_committed);
                        else read_u32_array(m2, ci.public_and_commitment_committed);
                    }
                    c.off = item;
                    c.skip();
                    f.commitments.push_back(std::move(ci));
                }
            }
        }
    }
    // coefficient table
    uint64_t ncoef = r.le64("coefficient count");
    const uint8_t* cb = r.take(ncoef * 32, "coefficient table");
    f.coeffs.resize(ncoef * 4);
    memcpy(f.coeffs.data(), cb, ncoef * 32);
    if (r.off != len) throw ParseError("r1cs: trailing bytes");
    for (size_t i = 0; i < f.bp_id.size(); i++)
        if (f.bp_id[i] >= f.bp_kind.size()) throw ParseError("r1cs: instruction references unknown blueprint");
    return f;
}

}  // namespace g16
