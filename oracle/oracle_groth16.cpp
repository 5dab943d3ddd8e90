// ORACLE — TEST INFRASTRUCTURE ONLY (see bn254_field.hpp header). Never linked into the product library.
//
// R1CS solver, Groth16 prover and pairing verifier restated from SURVEY.md Appendices A-F.
// Reference call sites: libraries/prover/impl/provers.go:144-157 (NewWitness -> groth16.Prove -> Proof.WriteTo),
// libraries/verifier/impl/verifiers.go:50-104 (groth16.Verify). The arithmetic lives in gnark v0.11.0
// (constraint/bn254/solver.go, backend/groth16/bn254/{prove,verify,marshal}.go) which is NOT on this box;
// this restatement is pinned by the reference's own shipped vk.chacha20 (pairing check) — see oracle/README.md.
#include "bn254_curve.hpp"
#include <functional>
#include <string>

void parallel_for(size_t n, int nthreads, const std::function<void(size_t, size_t)>& fn);
void ntt_inplace(Fr* a, size_t n, const Fr& w);
extern "C" void orc_init();
extern "C" void orc_compute_h(const u64*, const u64*, const u64*, size_t, size_t, const u64*, const u64*, u64*, int);

// =====================================================================================================
// R1CS solver  (SURVEY.md Appendix D "Instruction i", Appendix E)
// =====================================================================================================
static const uint32_t HINT_NBITS = 4115454955u, HINT_COUNT = 2138922168u, HINT_RANDOMIZE = 1774611027u,
                      HINT_BSB22 = 4156202267u;
static const uint32_t WIRE_CONST = 0xFFFFFFFFu;

struct SolveCtx {
    const uint32_t* cd;
    const Fr* coeffs;
    Fr* W;
    std::vector<uint8_t> solved;
};

static inline Fr term_value(const SolveCtx& s, uint32_t cid, uint32_t wid) {
    if (wid == WIRE_CONST) return s.coeffs[cid];
    return s.coeffs[cid] * s.W[wid];
}

// evaluates a hint/lookup input linear expression starting at cd[pos]; advances pos
static Fr eval_le(const SolveCtx& s, size_t& pos) {
    uint32_t nt = s.cd[pos++];
    Fr acc = Fr::zero();
    for (uint32_t t = 0; t < nt; t++) {
        uint32_t cid = s.cd[pos++], wid = s.cd[pos++];
        acc = acc + term_value(s, cid, wid);
    }
    return acc;
}

typedef int (*bsb22_cb)(void* user, const u64* inputs_mont, size_t n_inputs, u64* out_mont);

extern "C" int orc_solve(const uint32_t* calldata, size_t n_instr, const uint8_t* ins_kind, const uint64_t* ins_start,
                         const uint32_t* ins_wire_off, const uint32_t* ins_cons_off, const uint32_t* ins_lookup_tab,
                         const u64* coeffs_mont, const u64* lookup_tabs_mont /* [ntab][256][4] */,
                         u64* W_mont, size_t n_wires, size_t n_inputs /* public + secret, incl ONE */,
                         u64* A_out, u64* B_out, u64* C_out, const u64* randomize_mont, bsb22_cb bsb, void* bsb_user) {
    orc_init();
    SolveCtx s;
    s.cd = calldata;
    s.coeffs = (const Fr*)coeffs_mont;
    s.W = (Fr*)W_mont;
    s.solved.assign(n_wires, 0);
    for (size_t i = 0; i < n_inputs; i++) s.solved[i] = 1;
    Fr* A = (Fr*)A_out; Fr* B = (Fr*)B_out; Fr* C = (Fr*)C_out;
    for (size_t ins = 0; ins < n_instr; ins++) {
        size_t base = ins_start[ins];
        uint32_t len = calldata[base];
        (void)len;
        if (ins_kind[ins] == 0) {   // generic R1C: [len, nL, nR, nO, terms...]
            uint32_t n[3] = {calldata[base + 1], calldata[base + 2], calldata[base + 3]};
            size_t pos = base + 4;
            Fr sum[3], ucoef[3];
            int uside = -1;
            uint32_t uwire = 0;
            for (int side = 0; side < 3; side++) {
                sum[side] = Fr::zero();
                ucoef[side] = Fr::zero();
                for (uint32_t t = 0; t < n[side]; t++) {
                    uint32_t cid = calldata[pos++], wid = calldata[pos++];
                    if (wid != WIRE_CONST && !s.solved[wid]) {
                        if (uside >= 0 && (uside != side || uwire != wid)) return -10;   // two unknowns
                        uside = side; uwire = wid;
                        ucoef[side] = ucoef[side] + s.coeffs[cid];
                    } else {
                        sum[side] = sum[side] + term_value(s, cid, wid);
                    }
                }
            }
            if (uside >= 0) {
                Fr w;
                if (uside == 2) w = (sum[0] * sum[1] - sum[2]) * ucoef[2].inv();
                else if (uside == 0) {
                    if (sum[1].is_zero()) return -11;
                    w = (sum[2] * sum[1].inv() - sum[0]) * ucoef[0].inv();
                } else {
                    if (sum[0].is_zero()) return -11;
                    w = (sum[2] * sum[0].inv() - sum[1]) * ucoef[1].inv();
                }
                s.W[uwire] = w;
                s.solved[uwire] = 1;
                sum[uside] = sum[uside] + ucoef[uside] * w;
            } else if (sum[0] * sum[1] != sum[2]) {
                return -(int)(100 + 0);   // unsatisfied constraint
            }
            size_t ci = ins_cons_off[ins];
            A[ci] = sum[0]; B[ci] = sum[1]; C[ci] = sum[2];
        } else if (ins_kind[ins] == 1) {   // generic hint: [len, hintID, nIn, inputs..., outStart, outEnd]
            uint32_t hid = calldata[base + 1], nin = calldata[base + 2];
            size_t pos = base + 3;
            std::vector<Fr> in(nin);
            for (uint32_t k = 0; k < nin; k++) in[k] = eval_le(s, pos);
            uint32_t o0 = calldata[pos], o1 = calldata[pos + 1];
            if (hid == HINT_NBITS) {
                u64 v[4]; in[0].to_canon(v);
                for (uint32_t k = 0; k < o1 - o0; k++) {
                    int bit = k < 256 ? (int)((v[k / 64] >> (k % 64)) & 1) : 0;
                    s.W[o0 + k] = bit ? Fr::one() : Fr::zero();
                }
            } else if (hid == HINT_COUNT) {
                // inputs: nbRows, rowWidth, nbRows*rowWidth table values, then queries (rowWidth each)
                u64 t[4]; in[0].to_canon(t); size_t rows = t[0];
                in[1].to_canon(t); size_t width = t[0];
                size_t nq = (nin - 2 - rows * width) / width;
                std::vector<u64> cnt(rows, 0);
                for (size_t q = 0; q < nq; q++) {
                    const Fr* qv = &in[2 + rows * width + q * width];
                    for (size_t r = 0; r < rows; r++) {
                        bool eq = true;
                        for (size_t c = 0; c < width; c++) if (in[2 + r * width + c] != qv[c]) { eq = false; break; }
                        if (eq) { cnt[r]++; break; }
                    }
                }
                for (uint32_t k = 0; k < o1 - o0; k++) s.W[o0 + k] = Fr::from_u64(k < rows ? cnt[k] : 0);
            } else if (hid == HINT_RANDOMIZE) {
                if (!randomize_mont) return -20;
                memcpy(s.W[o0].l, randomize_mont, 32);
            } else if (hid == HINT_BSB22) {
                if (!bsb) return -21;
                Fr out;
                int rc = bsb(bsb_user, (const u64*)in.data(), nin, out.l);
                if (rc) return rc;
                s.W[o0] = out;
            } else {
                return -22;
            }
            for (uint32_t k = o0; k < o1; k++) s.solved[k] = 1;
        } else {   // lookup: [len, nbEntries, nIn, inputs...] ; outputs at WireOffset + k
            uint32_t nent = calldata[base + 1], nin = calldata[base + 2];
            size_t pos = base + 3;
            const Fr* tab = (const Fr*)lookup_tabs_mont + (size_t)ins_lookup_tab[ins] * 256;
            for (uint32_t k = 0; k < nin; k++) {
                Fr v = eval_le(s, pos);
                u64 c[4]; v.to_canon(c);
                if (c[1] | c[2] | c[3] || c[0] >= nent) return -30;
                uint32_t w = ins_wire_off[ins] + k;
                s.W[w] = tab[c[0]];
                s.solved[w] = 1;
            }
        }
    }
    for (size_t i = 0; i < n_wires; i++) if (!s.solved[i]) return -40;
    return 0;
}

// =====================================================================================================
// Proving key and prover (Appendix F.1, F.2)
// =====================================================================================================
struct Reader {
    const uint8_t* p; size_t len, off = 0; bool ok = true;
    const uint8_t* take(size_t n) {
        if (off + n > len) { ok = false; return nullptr; }
        const uint8_t* r = p + off; off += n; return r;
    }
    u64 be64() { const uint8_t* b = take(8); if (!b) return 0; u64 v = 0; for (int i = 0; i < 8; i++) v = (v << 8) | b[i]; return v; }
    uint32_t be32() { const uint8_t* b = take(4); if (!b) return 0; uint32_t v = 0; for (int i = 0; i < 4; i++) v = (v << 8) | b[i]; return v; }
};

struct OrcPK {
    u64 n;
    Fr n_inv, omega, omega_inv, coset_g, coset_g_inv;
    G1A alpha, beta, delta;
    std::vector<G1A> A, B, Z, K;
    G2A beta2, delta2;
    std::vector<G2A> B2;
    u64 nb_wires, nb_inf_a, nb_inf_b;
    std::vector<uint8_t> inf_a, inf_b;
    uint32_t n_commit_keys;
};

static bool read_g1_vec(Reader& r, std::vector<G1A>& out, int nthreads) {
    uint32_t cnt = r.be32();
    const uint8_t* raw = r.take((size_t)cnt * 32);
    if (!r.ok) return false;
    out.resize(cnt);
    bool good = true;
    parallel_for(cnt, nthreads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) if (g1_decompress(raw + 32 * i, out[i])) good = false;
    });
    return good;
}
static bool read_g2_vec(Reader& r, std::vector<G2A>& out, int nthreads) {
    uint32_t cnt = r.be32();
    const uint8_t* raw = r.take((size_t)cnt * 64);
    if (!r.ok) return false;
    out.resize(cnt);
    bool good = true;
    parallel_for(cnt, nthreads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) if (g2_decompress(raw + 64 * i, out[i])) good = false;
    });
    return good;
}

extern "C" void* orc_pk_parse(const uint8_t* data, size_t len, int nthreads) {
    orc_init();
    Reader r{data, len};
    OrcPK* pk = new OrcPK();
    pk->n = r.be64();
    Fr* hdr[5] = {&pk->n_inv, &pk->omega, &pk->omega_inv, &pk->coset_g, &pk->coset_g_inv};
    for (int i = 0; i < 5; i++) { const uint8_t* b = r.take(32); if (!b) { delete pk; return nullptr; } *hdr[i] = Fr::from_be(b); }
    r.take(1);   // withPrecompute
    G1A* g1s[3] = {&pk->alpha, &pk->beta, &pk->delta};
    for (int i = 0; i < 3; i++) { const uint8_t* b = r.take(32); if (!b || g1_decompress(b, *g1s[i])) { delete pk; return nullptr; } }
    bool ok = read_g1_vec(r, pk->A, nthreads) && read_g1_vec(r, pk->B, nthreads) && read_g1_vec(r, pk->Z, nthreads) &&
              read_g1_vec(r, pk->K, nthreads);
    if (ok) {
        const uint8_t* b = r.take(64); ok = b && !g2_decompress(b, pk->beta2);
        b = r.take(64); ok = ok && b && !g2_decompress(b, pk->delta2);
    }
    ok = ok && read_g2_vec(r, pk->B2, nthreads);
    if (ok) {
        pk->nb_wires = r.be64(); pk->nb_inf_a = r.be64(); pk->nb_inf_b = r.be64();
        const uint8_t* ia = r.take(pk->nb_wires); const uint8_t* ib = r.take(pk->nb_wires);
        ok = r.ok;
        if (ok) { pk->inf_a.assign(ia, ia + pk->nb_wires); pk->inf_b.assign(ib, ib + pk->nb_wires); }
        pk->n_commit_keys = r.be32();
        ok = ok && r.ok;
    }
    if (!ok) { delete pk; return nullptr; }
    return pk;
}
extern "C" void orc_pk_free(void* p) { delete (OrcPK*)p; }
// info[0..8) = n, |A|, |B|, |Z|, |K|, |B2|, nbWires, nCommitKeys
extern "C" void orc_pk_info(void* p, u64* info) {
    OrcPK* pk = (OrcPK*)p;
    info[0] = pk->n; info[1] = pk->A.size(); info[2] = pk->B.size(); info[3] = pk->Z.size(); info[4] = pk->K.size();
    info[5] = pk->B2.size(); info[6] = pk->nb_wires; info[7] = pk->n_commit_keys;
}
// which: 0 A,1 B,2 Z,3 K (G1 affine, 8 u64 each) ; 4 B2 (G2 affine, 16 u64) ; 5 infA, 6 infB (bytes) ;
// 7 {alpha,beta,delta} ; 8 {beta2,delta2} ; 9 {n_inv, omega, omega_inv, g, g_inv}
extern "C" const void* orc_pk_array(void* p, int which) {
    OrcPK* pk = (OrcPK*)p;
    switch (which) {
        case 0: return pk->A.data(); case 1: return pk->B.data(); case 2: return pk->Z.data(); case 3: return pk->K.data();
        case 4: return pk->B2.data(); case 5: return pk->inf_a.data(); case 6: return pk->inf_b.data();
        case 7: return &pk->alpha; case 8: return &pk->beta2; case 9: return &pk->n_inv;
    }
    return nullptr;
}

static void canon_vec(const Fr* in, size_t n, std::vector<u64>& out) {
    out.resize(4 * n);
    for (size_t i = 0; i < n; i++) in[i].to_canon(&out[4 * i]);
}
static size_t bitrev_sz(size_t x, int lg) { size_t r = 0; for (int i = 0; i < lg; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }

// W: full wire vector; A,B,C: per-constraint evaluations (ncons). wK selection: private wires [n_public, nbWires)
// except those listed in `skip` (committed + commitment wires; empty for ChaCha). r,s canonical LE limbs.
// out_proof: Ar(32) | Bs(64) | Krs(32).  inter (optional): affine points msmA, msmB1, msmK, msmZ (G1, 8 u64 each),
// then msmB2 (16 u64), then Ar, Bs1, Krs (8 each), Bs (16): total 4*8+16+3*8+16 = 88 u64.
// h_out (optional): n Fr coefficients in NATURAL order.
extern "C" int orc_prove(void* p, const u64* W_mont, const u64* A_ev, const u64* B_ev, const u64* C_ev, size_t ncons,
                         size_t n_public, const uint32_t* skip, size_t n_skip, const u64* r_canon, const u64* s_canon,
                         int nthreads, uint8_t* out_proof, u64* inter, u64* h_out) {
    OrcPK* pk = (OrcPK*)p;
    const Fr* W = (const Fr*)W_mont;
    size_t n = pk->n;
    int lg = 0; while (((size_t)1 << lg) < n) lg++;
    std::vector<Fr> h(n);
    orc_compute_h(A_ev, B_ev, C_ev, ncons, n, pk->omega.l, pk->coset_g.l, (u64*)h.data(), nthreads);
    if (h_out) memcpy(h_out, h.data(), n * 32);
    // gnark pairs natural coefficient j with stored Z[brev(j)]  (Appendix F.2)
    std::vector<Fr> hz(pk->Z.size());
    for (size_t i = 0; i < pk->Z.size(); i++) hz[i] = h[bitrev_sz(i, lg)];
    if (!h[n - 1].is_zero() && pk->Z.size() == n - 1 && !h[bitrev_sz(n - 1, lg)].is_zero()) return -1;
    std::vector<Fr> wA, wB, wK;
    for (size_t i = 0; i < pk->nb_wires; i++) {
        if (!pk->inf_a[i]) wA.push_back(W[i]);
        if (!pk->inf_b[i]) wB.push_back(W[i]);
    }
    std::vector<uint8_t> sk(pk->nb_wires, 0);
    for (size_t i = 0; i < n_skip; i++) sk[skip[i]] = 1;
    for (size_t i = n_public; i < pk->nb_wires; i++) if (!sk[i]) wK.push_back(W[i]);
    if (wA.size() != pk->A.size() || wB.size() != pk->B.size() || wK.size() != pk->K.size()) return -2;
    std::vector<u64> sA, sB, sK, sZ;
    canon_vec(wA.data(), wA.size(), sA); canon_vec(wB.data(), wB.size(), sB);
    canon_vec(wK.data(), wK.size(), sK); canon_vec(hz.data(), hz.size(), sZ);
    G1J mA = msm_pippenger<Fp>(pk->A.data(), sA.data(), wA.size(), nthreads);
    G1J mB1 = msm_pippenger<Fp>(pk->B.data(), sB.data(), wB.size(), nthreads);
    G1J mK = msm_pippenger<Fp>(pk->K.data(), sK.data(), wK.size(), nthreads);
    G1J mZ = msm_pippenger<Fp>(pk->Z.data(), sZ.data(), hz.size(), nthreads);
    G2J mB2 = msm_pippenger<Fp2>(pk->B2.data(), sB.data(), wB.size(), nthreads);
    Fr r = Fr::from_canon(r_canon), s = Fr::from_canon(s_canon);
    u64 rs_c[4]; (r * s).to_canon(rs_c);
    G1J d1 = G1J::from_aff(pk->delta);
    G1J Ar = mA.add_aff(pk->alpha).add(d1.mul(r_canon, 4));
    G1J Bs1 = mB1.add_aff(pk->beta).add(d1.mul(s_canon, 4));
    G2J Bs = mB2.add_aff(pk->beta2).add(G2J::from_aff(pk->delta2).mul(s_canon, 4));
    G1J Krs = mK.add(mZ).add(Ar.mul(s_canon, 4)).add(Bs1.mul(r_canon, 4)).add(d1.mul(rs_c, 4).neg());
    G1A aAr = Ar.to_aff(), aKrs = Krs.to_aff();
    G2A aBs = Bs.to_aff();
    g1_compress(aAr, out_proof);
    g2_compress(aBs, out_proof + 32);
    g1_compress(aKrs, out_proof + 96);
    if (inter) {
        G1A t;
        t = mA.to_aff(); memcpy(inter, &t, 64);
        t = mB1.to_aff(); memcpy(inter + 8, &t, 64);
        t = mK.to_aff(); memcpy(inter + 16, &t, 64);
        t = mZ.to_aff(); memcpy(inter + 24, &t, 64);
        G2A t2 = mB2.to_aff(); memcpy(inter + 32, &t2, 128);
        memcpy(inter + 48, &aAr, 64);
        t = Bs1.to_aff(); memcpy(inter + 56, &t, 64);
        memcpy(inter + 64, &aKrs, 64);
        memcpy(inter + 72, &aBs, 128);
    }
    return 0;
}

// =====================================================================================================
// Pairing verifier (Appendix F.4). Optimal ate pairing on BN254 written over the polynomial representation
// Fp12 = Fp[w]/(w^12 - 18 w^6 + 82) (u = w^6 - 9), affine Miller loop on the twist, plain square-and-multiply final
// exponentiation with the exponent (p^12-1)/r supplied by the caller (computed with Python big ints).
// Slow but short; it only has to agree with the reference's shipped vk.chacha20.
// =====================================================================================================
struct F12 {
    Fp c[12];
    static F12 one() { F12 r; for (auto& x : r.c) x = Fp::zero(); r.c[0] = Fp::one(); return r; }
    bool is_one() const { if (c[0] != Fp::one()) return false; for (int i = 1; i < 12; i++) if (!c[i].is_zero()) return false; return true; }
};
static Fp g_c18, g_c82, g_c9;
static F12 f12_mul(const F12& a, const F12& b) {
    Fp t[23];
    for (auto& x : t) x = Fp::zero();
    for (int i = 0; i < 12; i++) {
        if (a.c[i].is_zero()) continue;
        for (int j = 0; j < 12; j++) t[i + j] = t[i + j] + a.c[i] * b.c[j];
    }
    for (int i = 22; i >= 12; i--) {   // w^12 = 18 w^6 - 82
        t[i - 6] = t[i - 6] + t[i] * g_c18;
        t[i - 12] = t[i - 12] - t[i] * g_c82;
    }
    F12 r;
    for (int i = 0; i < 12; i++) r.c[i] = t[i];
    return r;
}
// e * w^k for e in Fp2, k < 6
static void f12_add_fp2_at(F12& f, const Fp2& e, int k) {
    f.c[k] = f.c[k] + (e.a0 - e.a1 * g_c9);
    f.c[k + 6] = f.c[k + 6] + e.a1;
}
// line through R with slope m (both on the twist, Fp2), evaluated at P in G1
static F12 line_eval(const G2A& R, const Fp2& m, const G1A& P) {
    F12 l;
    for (auto& x : l.c) x = Fp::zero();
    l.c[0] = P.y.neg();
    f12_add_fp2_at(l, m.mul_fp(P.x), 1);
    f12_add_fp2_at(l, R.y - m * R.x, 3);
    return l;
}
static Fp2 g_frob_x, g_frob_y;   // xi^((p-1)/3), xi^((p-1)/2)
static std::vector<u64> g_final_exp;
static const u64 ATE_LOOP_LO = 0x9d797039be763ba8ULL;   // 29793968203157093288 = 2^64 + this

static void div_small(u64 r[4], const u64 a[4], u64 d) {
    u128 rem = 0;
    for (int i = 3; i >= 0; i--) { u128 cur = (rem << 64) | a[i]; r[i] = (u64)(cur / d); rem = cur % d; }
}
extern "C" void orc_pairing_init(const u64* final_exp_limbs, size_t nlimbs) {
    orc_init();
    g_c18 = Fp::from_u64(18); g_c82 = Fp::from_u64(82); g_c9 = Fp::from_u64(9);
    g_final_exp.assign(final_exp_limbs, final_exp_limbs + nlimbs);
    u64 pm1[4], one[4] = {1, 0, 0, 0}, e3[4], e2[4];
    sub4(pm1, g_fp.M, one);
    div_small(e3, pm1, 3); div_small(e2, pm1, 2);
    Fp2 xi = {Fp::from_u64(9), Fp::one()};
    g_frob_x = xi.pow(e3, 4);
    g_frob_y = xi.pow(e2, 4);
}
static G2A g2_frob(const G2A& q) { return {q.x.conj() * g_frob_x, q.y.conj() * g_frob_y}; }

static F12 miller(const G2A& Q, const G1A& P) {
    F12 f = F12::one();
    if (Q.is_inf() || P.is_inf()) return f;
    G2A R = Q;
    auto dbl_step = [&]() {
        Fp2 m = (R.x.sqr().dbl() + R.x.sqr()) * R.y.dbl().inv();
        F12 l = line_eval(R, m, P);
        f = f12_mul(f12_mul(f, f), l);
        Fp2 x3 = m.sqr() - R.x.dbl();
        Fp2 y3 = m * (R.x - x3) - R.y;
        R = {x3, y3};
    };
    auto add_step = [&](const G2A& T) {
        Fp2 m = (T.y - R.y) * (T.x - R.x).inv();
        F12 l = line_eval(R, m, P);
        f = f12_mul(f, l);
        Fp2 x3 = m.sqr() - R.x - T.x;
        Fp2 y3 = m * (R.x - x3) - R.y;
        R = {x3, y3};
    };
    for (int i = 63; i >= 0; i--) {
        dbl_step();
        if ((ATE_LOOP_LO >> i) & 1) add_step(Q);
    }
    G2A Q1 = g2_frob(Q);
    G2A nQ2 = g2_frob(Q1).neg();
    add_step(Q1);
    // last line only (no point update needed)
    Fp2 m = (nQ2.y - R.y) * (nQ2.x - R.x).inv();
    f = f12_mul(f, line_eval(R, m, P));
    return f;
}
static F12 final_exp(const F12& f) {
    F12 r = F12::one();
    for (int i = (int)g_final_exp.size() * 64 - 1; i >= 0; i--) {
        r = f12_mul(r, r);
        if ((g_final_exp[i / 64] >> (i % 64)) & 1) r = f12_mul(r, f);
    }
    return r;
}
// prod e(P_i, Q_i) == 1 ?
extern "C" int orc_pairing_check(const u64* g1s, const u64* g2s, size_t n) {
    F12 f = F12::one();
    for (size_t i = 0; i < n; i++) {
        G1A P; G2A Q; memcpy(&P, g1s + 8 * i, 64); memcpy(&Q, g2s + 16 * i, 128);
        f = f12_mul(f, miller(Q, P));
    }
    return final_exp(f).is_one() ? 1 : 0;
}

struct OrcVK {
    G1A alpha, beta, delta;
    G2A beta2, gamma2, delta2;
    std::vector<G1A> K;
    F12 miller_alpha_beta;
    uint32_t n_commit;
};
extern "C" void* orc_vk_parse(const uint8_t* data, size_t len) {
    orc_init();
    Reader r{data, len};
    OrcVK* vk = new OrcVK();
    const uint8_t* b;
    bool ok = true;
    b = r.take(32); ok = ok && b && !g1_decompress(b, vk->alpha);
    b = r.take(32); ok = ok && b && !g1_decompress(b, vk->beta);
    b = r.take(64); ok = ok && b && !g2_decompress(b, vk->beta2);
    b = r.take(64); ok = ok && b && !g2_decompress(b, vk->gamma2);
    b = r.take(32); ok = ok && b && !g1_decompress(b, vk->delta);
    b = r.take(64); ok = ok && b && !g2_decompress(b, vk->delta2);
    ok = ok && read_g1_vec(r, vk->K, 1);
    if (!ok) { delete vk; return nullptr; }
    uint32_t outer = r.be32();
    for (uint32_t i = 0; i < outer; i++) { uint32_t inner = r.be32(); r.take((size_t)inner * 8); }
    vk->n_commit = r.be32();
    if (!r.ok) { delete vk; return nullptr; }
    vk->miller_alpha_beta = miller(vk->beta2, vk->alpha);
    return vk;
}
extern "C" void orc_vk_free(void* p) { delete (OrcVK*)p; }
extern "C" size_t orc_vk_nk(void* p) { return ((OrcVK*)p)->K.size(); }
// which: 0 {alpha,beta,delta} G1 ; 1 {beta2,gamma2,delta2} G2 ; 2 K
extern "C" const void* orc_vk_array(void* p, int which) {
    OrcVK* vk = (OrcVK*)p;
    return which == 0 ? (const void*)&vk->alpha : which == 1 ? (const void*)&vk->beta2 : (const void*)vk->K.data();
}

// proof = Ar(32)|Bs(64)|Krs(32)|u32 nCommit|... ; pub: n_pub Montgomery Fr (without the leading ONE).
// returns 1 = accepted, 0 = rejected, <0 = malformed. Only the commitment-free form (ChaCha) is handled here.
extern "C" int orc_verify(void* p, const uint8_t* proof, size_t proof_len, const u64* pub_mont, size_t n_pub) {
    OrcVK* vk = (OrcVK*)p;
    if (proof_len < 128) return -1;
    if (n_pub + 1 != vk->K.size()) return -2;
    G1A Ar, Krs; G2A Bs;
    if (g1_decompress(proof, Ar) || g2_decompress(proof + 32, Bs) || g1_decompress(proof + 96, Krs)) return -3;
    if (!g1_on_curve(Ar) || !g1_on_curve(Krs) || !g2_on_curve(Bs)) return -3;
    std::vector<u64> sc(4 * n_pub);
    for (size_t i = 0; i < n_pub; i++) { Fr v; memcpy(v.l, pub_mont + 4 * i, 32); v.to_canon(&sc[4 * i]); }
    G1J ks = msm_pippenger<Fp>(vk->K.data() + 1, sc.data(), n_pub, 1).add_aff(vk->K[0]);
    G1A kSum = ks.to_aff();
    // e(-Ar,Bs) * e(alpha,beta2) * e(kSum,gamma2) * e(Krs,delta2) == 1
    F12 f = miller(Bs, Ar.neg());
    f = f12_mul(f, vk->miller_alpha_beta);
    f = f12_mul(f, miller(vk->gamma2, kSum));
    f = f12_mul(f, miller(vk->delta2, Krs));
    return final_exp(f).is_one() ? 1 : 0;
}
