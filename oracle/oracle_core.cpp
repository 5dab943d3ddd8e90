// ORACLE — TEST INFRASTRUCTURE ONLY (see bn254_field.hpp header). Never linked into the product library.
//
// Field/curve/MSM/NTT part of the CPU restatement. Reference call chain being restated:
//   libraries/prover/impl/provers.go:148,216  groth16.Prove  ->  gnark v0.11.0 backend/groth16/bn254/prove.go
//   -> gnark-crypto v0.14.0 ecc/bn254 {multiexp.go, fr/fft/fft.go, g1.go, g2.go, marshal.go}  (none of it on this box).
#include "bn254_curve.hpp"
#include <algorithm>
#include <atomic>
#include <functional>

FieldParams g_fp, g_fr;
Fp g_b1;
Fp2 g_b2;

static const u64 P_MOD[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const u64 R_MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};

static u64 g_p_plus1_div4[4], g_p_minus3_div4[4], g_p_minus1_div2[4];
static bool g_inited = false;

static void shr4(u64 r[4], const u64 a[4], int k) {
    for (int i = 0; i < 4; i++) r[i] = (a[i] >> k) | (i < 3 ? (a[i + 1] << (64 - k)) : 0);
}

void parallel_for(size_t n, int nthreads, const std::function<void(size_t, size_t)>& fn) {
    if (nthreads <= 1 || n < 2) { fn(0, n); return; }
    std::vector<std::thread> th;
    size_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        size_t lo = t * per, hi = std::min(n, lo + per);
        if (lo >= hi) break;
        th.emplace_back(fn, lo, hi);
    }
    for (auto& t : th) t.join();
}

extern "C" void orc_init() {
    if (g_inited) return;
    g_fp.init(P_MOD);
    g_fr.init(R_MOD);
    g_b1 = Fp::from_u64(3);
    Fp2 xi = {Fp::from_u64(9), Fp::one()};
    g_b2 = Fp2{Fp::from_u64(3), Fp::zero()} * xi.inv();
    u64 one[4] = {1, 0, 0, 0}, three[4] = {3, 0, 0, 0}, t[4];
    add4(t, P_MOD, one);  shr4(g_p_plus1_div4, t, 2);
    sub4(t, P_MOD, three); shr4(g_p_minus3_div4, t, 2);
    sub4(t, P_MOD, one);  shr4(g_p_minus1_div2, t, 1);
    g_inited = true;
}

// ------------------------------------------------------------------ square roots
bool fp_sqrt(const Fp& a, Fp& out) {   // p = 3 mod 4
    Fp s = a.pow(g_p_plus1_div4);
    if (s.sqr() != a) return false;
    out = s;
    return true;
}
// Square root in Fp2 for p = 3 mod 4 (Adj & Rodriguez-Henriquez, "Square root computation over even extension fields")
bool fp2_sqrt(const Fp2& a, Fp2& out) {
    if (a.is_zero()) { out = a; return true; }
    Fp2 a1 = a.pow(g_p_minus3_div4, 4);
    Fp2 x0 = a1 * a;
    Fp2 alpha = a1 * x0;
    Fp2 minus_one = Fp2::one().neg();
    Fp2 x;
    if (alpha == minus_one) {
        x = Fp2{Fp::zero(), Fp::one()} * x0;
    } else {
        Fp2 b = (Fp2::one() + alpha).pow(g_p_minus1_div2, 4);
        x = b * x0;
    }
    if (x.sqr() != a) return false;
    out = x;
    return true;
}

// ------------------------------------------------------------------ (de)compression, SURVEY.md Appendix A
int g1_decompress(const uint8_t in[32], G1A& out) {
    uint8_t flag = in[0] & 0xC0;
    if (flag == 0x40) { out = {Fp::zero(), Fp::zero()}; return 0; }
    if (flag == 0x00) return -1;   // uncompressed form is not 32 bytes
    uint8_t buf[32];
    memcpy(buf, in, 32);
    buf[0] &= 0x3F;
    Fp x = Fp::from_be(buf);
    Fp y;
    if (!fp_sqrt(x.sqr() * x + g_b1, y)) return -2;
    bool want_largest = (flag == 0xC0);
    if (y.lex_largest() != want_largest) y = y.neg();
    out = {x, y};
    return 0;
}
int g2_decompress(const uint8_t in[64], G2A& out) {
    uint8_t flag = in[0] & 0xC0;
    if (flag == 0x40) { out = {Fp2::zero(), Fp2::zero()}; return 0; }
    if (flag == 0x00) return -1;
    uint8_t buf[64];
    memcpy(buf, in, 64);
    buf[0] &= 0x3F;
    Fp2 x = {Fp::from_be(buf + 32), Fp::from_be(buf)};   // X.A1 || X.A0
    Fp2 y;
    if (!fp2_sqrt(x.sqr() * x + g_b2, y)) return -2;
    bool want_largest = (flag == 0xC0);
    if (y.lex_largest() != want_largest) y = y.neg();
    out = {x, y};
    return 0;
}
void g1_compress(const G1A& p, uint8_t out[32]) {
    if (p.is_inf()) { memset(out, 0, 32); out[0] = 0x40; return; }
    p.x.to_be(out);
    out[0] |= p.y.lex_largest() ? 0xC0 : 0x80;
}
void g2_compress(const G2A& p, uint8_t out[64]) {
    if (p.is_inf()) { memset(out, 0, 64); out[0] = 0x40; return; }
    p.x.a1.to_be(out);
    p.x.a0.to_be(out + 32);
    out[0] |= p.y.lex_largest() ? 0xC0 : 0x80;
}

// ------------------------------------------------------------------ MSM
template <class F>
Jac<F> msm_naive(const Aff<F>* pts, const u64* scalars, size_t n) {
    Jac<F> acc = Jac<F>::inf();
    for (size_t i = 0; i < n; i++) acc = acc.add(Jac<F>::from_aff(pts[i]).mul(scalars + 4 * i, 4));
    return acc;
}

// mixed addition Jacobian += affine (EFD madd-2007-bl) with the exceptional cases handled exactly
template <class F>
static inline void jac_madd(Jac<F>& p, const Aff<F>& q) {
    if (q.is_inf()) return;
    if (p.is_inf()) { p = Jac<F>::from_aff(q); return; }
    F Z1Z1 = p.Z.sqr();
    F U2 = q.x * Z1Z1;
    F S2 = q.y * p.Z * Z1Z1;
    if (U2 == p.X) {
        if (S2 == p.Y) { p = p.dbl(); return; }
        p = Jac<F>::inf();
        return;
    }
    F H = U2 - p.X;
    F HH = H.sqr();
    F I = HH.dbl().dbl();
    F J = H * I;
    F r = (S2 - p.Y).dbl();
    F V = p.X * I;
    Jac<F> o;
    o.X = r.sqr() - J - V.dbl();
    o.Y = r * (V - o.X) - (p.Y * J).dbl();
    o.Z = (p.Z + H).sqr() - Z1Z1 - HH;
    p = o;
}

static int pick_window(size_t n) {
    if (n < 32) return 3;
    int lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    int c = lg - 2;              // close to the min of ceil(254/c)*(n+2^c) for the sizes used here
    if (c < 4) c = 4;
    if (c > 16) c = 16;
    return c;
}

// Signed-digit Pippenger (same algorithm family as gnark-crypto multiexp.go:_innerMsm / partitionScalars:
// digits in [-2^(c-1), 2^(c-1)], running-sum bucket reduction, windows processed in parallel).
template <class F>
Jac<F> msm_pippenger(const Aff<F>* pts, const u64* scalars, size_t n, int nthreads) {
    if (n == 0) return Jac<F>::inf();
    const int c = pick_window(n);
    const int nwin = (255 + c - 1) / c + 1;   // one spare window for the final carry
    const int half = 1 << (c - 1);
    // recode
    std::vector<int32_t> digits((size_t)nwin * n);
    parallel_for(n, nthreads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            const u64* k = scalars + 4 * i;
            int carry = 0;
            for (int w = 0; w < nwin; w++) {
                int bit = w * c;
                int64_t d = carry;
                if (bit < 256) {
                    int limb = bit / 64, off = bit % 64;
                    u64 v = k[limb] >> off;
                    if (off + c > 64 && limb < 3) v |= k[limb + 1] << (64 - off);
                    d += (int64_t)(v & (((u64)1 << c) - 1));
                }
                if (d > half) { d -= (1 << c); carry = 1; } else carry = 0;
                digits[(size_t)w * n + i] = (int32_t)d;
            }
        }
    });
    std::vector<Jac<F>> wsum(nwin);
    std::atomic<int> next(0);
    auto worker = [&]() {
        std::vector<Jac<F>> buckets(half);
        for (;;) {
            int w = next.fetch_add(1);
            if (w >= nwin) break;
            for (auto& b : buckets) b = Jac<F>::inf();
            const int32_t* dg = &digits[(size_t)w * n];
            for (size_t i = 0; i < n; i++) {
                int32_t d = dg[i];
                if (d > 0) jac_madd(buckets[d - 1], pts[i]);
                else if (d < 0) jac_madd(buckets[-d - 1], pts[i].neg());
            }
            Jac<F> run = Jac<F>::inf(), tot = Jac<F>::inf();
            for (int k = half - 1; k >= 0; k--) {
                run = run.add(buckets[k]);
                tot = tot.add(run);
            }
            wsum[w] = tot;
        }
    };
    int nt = std::max(1, std::min(nthreads, nwin));
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    Jac<F> acc = Jac<F>::inf();
    for (int w = nwin - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) acc = acc.dbl();
        acc = acc.add(wsum[w]);
    }
    return acc;
}

template Jac<Fp> msm_pippenger<Fp>(const Aff<Fp>*, const u64*, size_t, int);
template Jac<Fp2> msm_pippenger<Fp2>(const Aff<Fp2>*, const u64*, size_t, int);
template Jac<Fp> msm_naive<Fp>(const Aff<Fp>*, const u64*, size_t);
template Jac<Fp2> msm_naive<Fp2>(const Aff<Fp2>*, const u64*, size_t);

// ------------------------------------------------------------------ NTT (gnark-crypto fr/fft restated as a textbook radix-2 transform)
static inline size_t bitrev(size_t x, int lg) {
    size_t r = 0;
    for (int i = 0; i < lg; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
// in-place, natural order in -> natural order out ; w = primitive n-th root (or its inverse); no scaling
void ntt_inplace(Fr* a, size_t n, const Fr& w) {
    int lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev(i, lg);
        if (i < j) std::swap(a[i], a[j]);
    }
    std::vector<Fr> tw(n / 2);
    if (n >= 2) {
        tw[0] = Fr::one();
        for (size_t i = 1; i < n / 2; i++) tw[i] = tw[i - 1] * w;
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        size_t half = len / 2, step = n / len;
        for (size_t s = 0; s < n; s += len)
            for (size_t k = 0; k < half; k++) {
                Fr u = a[s + k], v = a[s + k + half] * tw[k * step];
                a[s + k] = u + v;
                a[s + k + half] = u - v;
            }
    }
}

// ------------------------------------------------------------------ C API: fields, points, MSM, NTT
extern "C" {

void orc_f_to_mont(int fld, const u64* in, u64* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (fld == 0) { Fp v = Fp::from_canon(in + 4 * i); memcpy(out + 4 * i, v.l, 32); }
        else { Fr v = Fr::from_canon(in + 4 * i); memcpy(out + 4 * i, v.l, 32); }
    }
}
void orc_f_from_mont(int fld, const u64* in, u64* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (fld == 0) { Fp v; memcpy(v.l, in + 4 * i, 32); v.to_canon(out + 4 * i); }
        else { Fr v; memcpy(v.l, in + 4 * i, 32); v.to_canon(out + 4 * i); }
    }
}
// op: 0 add, 1 sub, 2 mul, 3 inv(a), 4 sqr(a), 5 neg(a)
void orc_f_op(int fld, int op, const u64* a, const u64* b, u64* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (fld == 0) {
            Fp x, y = Fp::zero(), r; memcpy(x.l, a + 4 * i, 32);
            if (b) memcpy(y.l, b + 4 * i, 32);
            switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break;
                          case 3: r = x.inv(); break; case 4: r = x.sqr(); break; default: r = x.neg(); }
            memcpy(out + 4 * i, r.l, 32);
        } else {
            Fr x, y = Fr::zero(), r; memcpy(x.l, a + 4 * i, 32);
            if (b) memcpy(y.l, b + 4 * i, 32);
            switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break;
                          case 3: r = x.inv(); break; case 4: r = x.sqr(); break; default: r = x.neg(); }
            memcpy(out + 4 * i, r.l, 32);
        }
    }
}
void orc_modulus(int fld, u64* out) { memcpy(out, fld == 0 ? g_fp.M : g_fr.M, 32); }

int orc_g1_decompress(const uint8_t* in, u64* out, size_t n, int nthreads) {
    std::atomic<int> rc(0);
    parallel_for(n, nthreads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            G1A p;
            int r = g1_decompress(in + 32 * i, p);
            if (r) rc = r;
            memcpy(out + 8 * i, &p, 64);
        }
    });
    return rc;
}
int orc_g2_decompress(const uint8_t* in, u64* out, size_t n, int nthreads) {
    std::atomic<int> rc(0);
    parallel_for(n, nthreads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            G2A p;
            int r = g2_decompress(in + 64 * i, p);
            if (r) rc = r;
            memcpy(out + 16 * i, &p, 128);
        }
    });
    return rc;
}
void orc_g1_compress(const u64* in, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) { G1A p; memcpy(&p, in + 8 * i, 64); g1_compress(p, out + 32 * i); }
}
void orc_g2_compress(const u64* in, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) { G2A p; memcpy(&p, in + 16 * i, 128); g2_compress(p, out + 64 * i); }
}
int orc_g1_on_curve(const u64* in, size_t n) {
    for (size_t i = 0; i < n; i++) { G1A p; memcpy(&p, in + 8 * i, 64); if (!g1_on_curve(p)) return 0; }
    return 1;
}
int orc_g2_on_curve(const u64* in, size_t n) {
    for (size_t i = 0; i < n; i++) { G2A p; memcpy(&p, in + 16 * i, 128); if (!g2_on_curve(p)) return 0; }
    return 1;
}
// scalars: canonical LE limbs. mode 0 = Pippenger, 1 = naive double-and-add
void orc_g1_msm(const u64* pts, const u64* scalars, size_t n, int nthreads, int mode, u64* out_aff) {
    G1J r = mode ? msm_naive<Fp>((const G1A*)pts, scalars, n) : msm_pippenger<Fp>((const G1A*)pts, scalars, n, nthreads);
    G1A a = r.to_aff();
    memcpy(out_aff, &a, 64);
}
void orc_g2_msm(const u64* pts, const u64* scalars, size_t n, int nthreads, int mode, u64* out_aff) {
    G2J r = mode ? msm_naive<Fp2>((const G2A*)pts, scalars, n) : msm_pippenger<Fp2>((const G2A*)pts, scalars, n, nthreads);
    G2A a = r.to_aff();
    memcpy(out_aff, &a, 128);
}
// out = a + b (affine in, affine out) ; out = k*a
void orc_g1_add(const u64* a, const u64* b, u64* out) {
    G1A pa, pb; memcpy(&pa, a, 64); memcpy(&pb, b, 64);
    G1A r = G1J::from_aff(pa).add(G1J::from_aff(pb)).to_aff();
    memcpy(out, &r, 64);
}
void orc_g1_mul(const u64* a, const u64* k_canon, u64* out) {
    G1A pa; memcpy(&pa, a, 64);
    G1A r = G1J::from_aff(pa).mul(k_canon, 4).to_aff();
    memcpy(out, &r, 64);
}
void orc_g2_add(const u64* a, const u64* b, u64* out) {
    G2A pa, pb; memcpy(&pa, a, 128); memcpy(&pb, b, 128);
    G2A r = G2J::from_aff(pa).add(G2J::from_aff(pb)).to_aff();
    memcpy(out, &r, 128);
}
void orc_g2_mul(const u64* a, const u64* k_canon, u64* out) {
    G2A pa; memcpy(&pa, a, 128);
    G2A r = G2J::from_aff(pa).mul(k_canon, 4).to_aff();
    memcpy(out, &r, 128);
}
// P_i = k_i * G (fixed base, for building synthetic MSM inputs), G given affine
void orc_g1_fixed_base(const u64* g, const u64* ks_canon, size_t n, int nthreads, u64* out) {
    G1A pg; memcpy(&pg, g, 64);
    parallel_for(n, nthreads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            G1A r = G1J::from_aff(pg).mul(ks_canon + 4 * i, 4).to_aff();
            memcpy(out + 8 * i, &r, 64);
        }
    });
}
void orc_g2_fixed_base(const u64* g, const u64* ks_canon, size_t n, int nthreads, u64* out) {
    G2A pg; memcpy(&pg, g, 128);
    parallel_for(n, nthreads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            G2A r = G2J::from_aff(pg).mul(ks_canon + 4 * i, 4).to_aff();
            memcpy(out + 16 * i, &r, 128);
        }
    });
}

// data: n Montgomery Fr, natural order in and out. inverse!=0 -> uses w^-1 and scales by 1/n
void orc_ntt(u64* data, size_t n, const u64* omega_mont, int inverse) {
    Fr w; memcpy(w.l, omega_mont, 32);
    Fr* a = (Fr*)data;
    if (inverse) {
        ntt_inplace(a, n, w.inv());
        Fr ninv = Fr::from_u64(n).inv();
        for (size_t i = 0; i < n; i++) a[i] = a[i] * ninv;
    } else {
        ntt_inplace(a, n, w);
    }
}

// H = (A*B - C)/Z_H on the coset g*<w>, SURVEY.md Appendix F.2 (restates gnark prove.go:computeH).
// a,b,c: ncons Montgomery Fr each (zero-padded to n here). h_out: n coefficients in NATURAL order.
void orc_compute_h(const u64* a_in, const u64* b_in, const u64* c_in, size_t ncons, size_t n,
                   const u64* omega_mont, const u64* coset_gen_mont, u64* h_out, int nthreads) {
    Fr w; memcpy(w.l, omega_mont, 32);
    Fr g; memcpy(g.l, coset_gen_mont, 32);
    Fr winv = w.inv(), ninv = Fr::from_u64(n).inv();
    std::vector<Fr> v[3];
    const u64* src[3] = {a_in, b_in, c_in};
    std::vector<Fr> gp(n);
    gp[0] = Fr::one();
    for (size_t i = 1; i < n; i++) gp[i] = gp[i - 1] * g;
    auto one_vec = [&](size_t k) {
        v[k].assign(n, Fr::zero());
        memcpy(v[k].data(), src[k], ncons * 32);
        ntt_inplace(v[k].data(), n, winv);
        for (size_t i = 0; i < n; i++) v[k][i] = v[k][i] * ninv * gp[i];
        ntt_inplace(v[k].data(), n, w);
    };
    if (nthreads >= 3) {
        std::thread t0(one_vec, 0), t1(one_vec, 1);
        one_vec(2);
        t0.join(); t1.join();
    } else {
        for (size_t k = 0; k < 3; k++) one_vec(k);
    }
    // den = 1/(g^n - 1)
    Fr gn = gp[n - 1] * g;
    Fr den = (gn - Fr::one()).inv();
    std::vector<Fr>& h = v[0];
    for (size_t i = 0; i < n; i++) h[i] = (v[0][i] * v[1][i] - v[2][i]) * den;
    ntt_inplace(h.data(), n, winv);
    Fr ginv = g.inv();
    Fr s = ninv;
    for (size_t i = 0; i < n; i++) { h[i] = h[i] * s; s = s * ginv; }
    memcpy(h_out, h.data(), n * 32);
}

}  // extern "C"
