// Device kernels of the Groth16 prove path other than MSM / NTT:
//   * point decompression of the proving key            (replaces gnark marshal.go:311-348 ProvingKey.readFrom — a19)
//   * ChaCha20 witness assignment                        (replaces provers.go:79-142 + utils/bytes.go:11-47 — a3..a6)
//   (the batched R1CS solver lives in solver.cuh / k_solver.cu)
//   * proof assembly + serialisation                     (replaces prove.go:174-295 tail, marshal.go:32-59 — a16, a18)
#pragma once
#include "prover_api.hpp"

namespace g16 {

// ------------------------------------------------------------------------------------------------ decompression
FD Fp fp_from_be32(const uint8_t* b, bool mask_flags) {
    Fp v;
    for (int i = 0; i < 8; i++) {
        const uint8_t* q = b + 28 - 4 * i;
        v.l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    if (mask_flags) v.l[7] &= 0x3FFFFFFFu;
    return v.to_mont();
}
FD void fp_to_be32(const Fp& m, uint8_t* b) {
    Fp c = m.from_mont();
    for (int i = 0; i < 8; i++) {
        uint8_t* q = b + 28 - 4 * i;
        q[0] = (uint8_t)(c.l[i] >> 24); q[1] = (uint8_t)(c.l[i] >> 16); q[2] = (uint8_t)(c.l[i] >> 8); q[3] = (uint8_t)c.l[i];
    }
}
FD Fp fp_three() { Fp o = Fp::one(); return o + o + o; }
FD Fp2 g2_coeff_b() {   // 3/(9+u)
    Fp o = Fp::one();
    Fp three = o + o + o;
    Fp nine = three + three + three;
    Fp2 xi = {nine, o};
    return Fp2{three, Fp::zero()} * xi.inv();
}
// exponents derived from p: which = 0 (p+1)/4, 1 (p-3)/4, 2 (p-1)/2
FD void fp_exponent(int which, uint32_t e[8]) {
    uint32_t t[8];
    for (int i = 0; i < 8; i++) t[i] = FpParams::mod(i);
    int sh;
    if (which == 0) { t[0] += 1; sh = 2; }        // low limb ...47: no carry
    else if (which == 1) { t[0] -= 3; sh = 2; }
    else { t[0] -= 1; sh = 1; }
    for (int i = 0; i < 8; i++) e[i] = (t[i] >> sh) | (i < 7 ? (t[i + 1] << (32 - sh)) : 0u);
}

// in: 32-byte compressed G1 points; out: affine Montgomery. err: bit 0 = not on curve, bit 1 = bad flag
__global__ void decompress_g1_kernel(const uint8_t* __restrict__ in, uint32_t n, G1Affine* __restrict__ out,
                                     uint32_t* __restrict__ err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* b = in + (size_t)i * 32;
    uint8_t flag = b[0] & 0xC0;
    if (flag == 0x40) { out[i] = G1Affine::inf(); return; }
    if (flag == 0x00) { atomicOr(err, 2u); out[i] = G1Affine::inf(); return; }
    Fp x = fp_from_be32(b, true);
    Fp rhs = x.sqr() * x + fp_three();
    uint32_t e[8];
    fp_exponent(0, e);
    Fp y = rhs.pow(e);
    if (y.sqr() != rhs) { atomicOr(err, 1u); out[i] = G1Affine::inf(); return; }
    bool want_largest = flag == 0xC0;
    if (y.lex_largest() != want_largest) y = y.neg();
    out[i] = {x, y};
}
__global__ void decompress_g2_kernel(const uint8_t* __restrict__ in, uint32_t n, G2Affine* __restrict__ out,
                                     uint32_t* __restrict__ err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* b = in + (size_t)i * 64;
    uint8_t flag = b[0] & 0xC0;
    if (flag == 0x40) { out[i] = G2Affine::inf(); return; }
    if (flag == 0x00) { atomicOr(err, 2u); out[i] = G2Affine::inf(); return; }
    Fp2 x = {fp_from_be32(b + 32, false), fp_from_be32(b, true)};   // X.A1 || X.A0
    Fp2 a = x.sqr() * x + g2_coeff_b();
    Fp2 y;
    if (a.is_zero()) {
        y = a;
    } else {
        // square root in Fp2, p = 3 mod 4 (Adj & Rodriguez-Henriquez, alg. 9)
        uint32_t e[8];
        fp_exponent(1, e);
        Fp2 a1 = a.pow(e);
        Fp2 x0 = a1 * a;
        Fp2 alpha = a1 * x0;
        if (alpha == Fp2::one().neg()) {
            y = Fp2{Fp::zero(), Fp::one()} * x0;
        } else {
            fp_exponent(2, e);
            Fp2 bb = (Fp2::one() + alpha).pow(e);
            y = bb * x0;
        }
        if (y.sqr() != a) { atomicOr(err, 1u); out[i] = G2Affine::inf(); return; }
    }
    bool want_largest = flag == 0xC0;
    if (y.lex_largest() != want_largest) y = y.neg();
    out[i] = {x, y};
}

FD void g1_compress(const G1Affine& p, uint8_t* out) {
    if (p.is_inf()) { for (int i = 0; i < 32; i++) out[i] = 0; out[0] = 0x40; return; }
    fp_to_be32(p.x, out);
    out[0] |= p.y.lex_largest() ? 0xC0 : 0x80;
}
FD void g2_compress(const G2Affine& p, uint8_t* out) {
    if (p.is_inf()) { for (int i = 0; i < 64; i++) out[i] = 0; out[0] = 0x40; return; }
    fp_to_be32(p.x.a1, out);
    fp_to_be32(p.x.a0, out + 32);
    out[0] |= p.y.lex_largest() ? 0xC0 : 0x80;
}

// big-endian 32-byte canonical scalars -> Fr canonical limbs (NOT Montgomery)
__global__ void scalars_from_be_kernel(const uint8_t* __restrict__ in, uint32_t n, Fr* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* b = in + (size_t)i * 32;
    Fr v;
    for (int k = 0; k < 8; k++) {
        const uint8_t* q = b + 28 - 4 * k;
        v.l[k] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    // reduce into [0, r): inputs are < 2^256 < 6r
    for (int k = 0; k < 5; k++) v.reduce_once();
    out[i] = v;
}

// ------------------------------------------------------------------------------------------------ ChaCha20 witness
FD uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define G16_QR(a, b, c, d)                                   \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16);            \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);            \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);             \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);

FD uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
FD uint32_t be32(const uint8_t* p) { return (uint32_t)p[3] | ((uint32_t)p[2] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[0] << 24); }

// One thread per request: ciphertext = ChaCha20(key, nonce, counter) xor input (RFC 7539 block function; the reference
// calls x/crypto/chacha20, provers.go:95-101), then the witness in gnark's order: ONE | Counter[32] | Nonce[3][32] |
// In[16][32] | Out[16][32] | Key[8][32], every word LSB-first; In/Out words are loaded big-endian, Key/Nonce words
// little-endian (provers.go:106-142, utils/bytes.go:11-47).
__global__ void chacha_witness_kernel(const uint8_t* __restrict__ keys, const uint8_t* __restrict__ nonces,
                                      const uint32_t* __restrict__ counters, const uint8_t* __restrict__ inputs,
                                      uint32_t n, Fr* __restrict__ W, size_t w_stride, uint8_t* __restrict__ ct_out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* key = keys + (size_t)i * 32;
    const uint8_t* nonce = nonces + (size_t)i * 12;
    const uint8_t* in = inputs + (size_t)i * 64;
    uint32_t st[16], x[16];
    st[0] = 0x61707865u; st[1] = 0x3320646eu; st[2] = 0x79622d32u; st[3] = 0x6b206574u;
    for (int k = 0; k < 8; k++) st[4 + k] = le32(key + 4 * k);
    st[12] = counters[i];
    for (int k = 0; k < 3; k++) st[13 + k] = le32(nonce + 4 * k);
    for (int k = 0; k < 16; k++) x[k] = st[k];
    for (int r = 0; r < 10; r++) {
        G16_QR(0, 4, 8, 12) G16_QR(1, 5, 9, 13) G16_QR(2, 6, 10, 14) G16_QR(3, 7, 11, 15)
        G16_QR(0, 5, 10, 15) G16_QR(1, 6, 11, 12) G16_QR(2, 7, 8, 13) G16_QR(3, 4, 9, 14)
    }
    uint8_t ct[64];
    for (int k = 0; k < 16; k++) {
        uint32_t ks = x[k] + st[k];
        for (int b = 0; b < 4; b++) ct[4 * k + b] = in[4 * k + b] ^ (uint8_t)(ks >> (8 * b));
    }
    for (int k = 0; k < 64; k++) ct_out[(size_t)i * 64 + k] = ct[k];
    // wire-major layout: wire k of request i lives at W[k * w_stride + i] (coalesced across the requests of a warp)
    Fr* w = W + i;
    const Fr one = Fr::one(), zero = Fr::zero();
    size_t pos = 0;
    w[pos] = one;
    pos += w_stride;
    auto put_word = [&](uint32_t word) {
        for (int b = 0; b < 32; b++) { w[pos] = ((word >> b) & 1u) ? one : zero; pos += w_stride; }
    };
    put_word(counters[i]);
    for (int k = 0; k < 3; k++) put_word(le32(nonce + 4 * k));
    for (int k = 0; k < 16; k++) put_word(be32(in + 4 * k));
    for (int k = 0; k < 16; k++) put_word(be32(ct + 4 * k));
    for (int k = 0; k < 8; k++) put_word(le32(key + 4 * k));
}

// generic: wire 0 = 1, wires 1..nw = witness[i][..]   (wire-major output, see above)
__global__ void witness_copy_kernel(const Fr* __restrict__ witness, uint32_t n_witness, uint32_t batch, Fr* __restrict__ W,
                                    size_t w_stride) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)batch * (n_witness + 1)) return;
    uint32_t i = (uint32_t)(gid / (n_witness + 1)), k = (uint32_t)(gid % (n_witness + 1));
    W[(size_t)k * w_stride + i] = k == 0 ? Fr::one() : witness[(size_t)i * n_witness + (k - 1)];
}

// wire-major W -> one row of nb_wires values per witness (test / gnark-shaped output only)
__global__ void wires_to_rows_kernel(const Fr* __restrict__ W, size_t w_stride, uint32_t batch, uint32_t nb_wires,
                                     Fr* __restrict__ out) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)batch * nb_wires) return;
    uint32_t i = (uint32_t)(gid % batch), k = (uint32_t)(gid / batch);
    out[(size_t)i * nb_wires + k] = W[(size_t)k * w_stride + i];
}

// ------------------------------------------------------------------------------------------------ proof assembly
// 
//   Ar  = msmA + alpha + r*delta          Bs1 = msmB1 + beta + s*delta        Bs = msmB2 + beta2 + s*delta2
//   Krs = msmK + msmZ + s*Ar + r*Bs1 - (r*s)*delta
// delta, delta2 are fixed: their multiples j*16^i*delta (i < 64, 1 <= j <= 15) are tabulated at init, so k*delta is at most
// 64 mixed additions and no doubling. The work of one proof is spread over three launches x several roles (blockIdx.y)
// so that the longest serial chain is one 254-bit double-and-add instead of six.

static const int FB_WINDOWS = 64, FB_ENTRIES = 15;

// tab[i*15 + (j-1)] = j * 16^i * base  (affine). One thread per window.
template <class C>
__global__ void fixed_base_table_kernel(typename C::A base, typename C::A* __restrict__ tab) {
    typedef typename C::X X;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= FB_WINDOWS) return;
    X b = X::from_affine(base);
    for (int d = 0; d < 4 * i; d++) b = b.dbl();
    X acc = b;
    tab[i * FB_ENTRIES] = acc.to_affine();
    for (int j = 1; j < FB_ENTRIES; j++) {
        acc.add(b);
        tab[i * FB_ENTRIES + j] = acc.to_affine();
    }
}
template <class C>
FD typename C::X fixed_base_mul(const typename C::A* __restrict__ tab, Scalar256 k) {
    typename C::X acc = C::X::inf();
#pragma unroll
    for (int wi = 0; wi < 8; wi++) {
        uint32_t word = k.w[wi];
#pragma unroll 1
        for (int j = 0; j < 8; j++) {
            uint32_t nib = (word >> (4 * j)) & 15u;
            if (nib) acc.madd(tab[(wi * 8 + j) * FB_ENTRIES + (nib - 1)], false);
        }
    }
    return acc;
}
FD Scalar256 scalar_of(const Fr& k) {
    Scalar256 s;
    for (int i = 0; i < 8; i++) s.w[i] = k.l[i];
    return s;
}

// rs: canonical limbs, r at [2i], s at [2i+1].
// phase 1, role (blockIdx.y) 0: Ar ; 1: Bs1 ; 2: Bs (compressed straight into the proof)
__global__ void __launch_bounds__(64)
assemble_phase1_kernel(AssemblyKeys keys, uint32_t n, const G1XYZZ* __restrict__ mA, const G1XYZZ* __restrict__ mB1,
                       const G2XYZZ* __restrict__ mB2, const Fr* __restrict__ rs, G1XYZZ* __restrict__ Ar_out,
                       G1XYZZ* __restrict__ Bs1_out, uint8_t* __restrict__ out, size_t out_stride) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t role = blockIdx.y;
    if (role == 0) {
        G1XYZZ v = fixed_base_mul<G1>(keys.delta_tab, scalar_of(rs[2 * i]));
        v.add(mA[i]);
        v.madd(keys.alpha, false);
        Ar_out[i] = v;
    } else if (role == 1) {
        G1XYZZ v = fixed_base_mul<G1>(keys.delta_tab, scalar_of(rs[2 * i + 1]));
        v.add(mB1[i]);
        v.madd(keys.beta, false);
        Bs1_out[i] = v;
    } else {
        G2XYZZ v = fixed_base_mul<G2>(keys.delta2_tab, scalar_of(rs[2 * i + 1]));
        v.add(mB2[i]);
        v.madd(keys.beta2, false);
        g2_compress(v.to_affine(), out + (size_t)i * out_stride + 32);
    }
}
// phase 2, role 0: s*Ar ; 1: r*Bs1
__global__ void __launch_bounds__(64)
assemble_phase2_kernel(uint32_t n, const G1XYZZ* __restrict__ Ar, const G1XYZZ* __restrict__ Bs1, const Fr* __restrict__ rs,
                       G1XYZZ* __restrict__ sAr, G1XYZZ* __restrict__ rBs1) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (blockIdx.y == 0) sAr[i] = scalar_mul(Ar[i], scalar_of(rs[2 * i + 1]));
    else rBs1[i] = scalar_mul(Bs1[i], scalar_of(rs[2 * i]));
}
// phase 3: Krs, compression of Ar and Krs, trailer of Proof.WriteTo without commitments (Appendix C)
__global__ void __launch_bounds__(64)
assemble_phase3_kernel(AssemblyKeys keys, uint32_t n, const G1XYZZ* __restrict__ mK, const G1XYZZ* __restrict__ mZ,
                       const G1XYZZ* __restrict__ Ar, const G1XYZZ* __restrict__ sAr, const G1XYZZ* __restrict__ rBs1,
                       const Fr* __restrict__ rs, uint8_t* __restrict__ out, size_t out_stride) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t* o = out + (size_t)i * out_stride;
    if (blockIdx.y == 1) {
        g1_compress(Ar[i].to_affine(), o);
        o[128] = 0; o[129] = 0; o[130] = 0; o[131] = 0;
        o[132] = 0x40;
        for (int k = 133; k < 164; k++) o[k] = 0;
        return;
    }
    Fr r = rs[2 * i], s = rs[2 * i + 1];
    Fr rsm = (r.to_mont() * s.to_mont()).from_mont();
    G1XYZZ Krs = fixed_base_mul<G1>(keys.delta_tab, scalar_of(rsm)).neg();
    Krs.add(mK[i]);
    Krs.add(mZ[i]);
    Krs.add(sAr[i]);
    Krs.add(rBs1[i]);
    g1_compress(Krs.to_affine(), o + 96);
}

}  // namespace g16
