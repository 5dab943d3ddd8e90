// Device kernels of the Groth16 prove path other than MSM / NTT:
//   * point decompression of the proving key            (replaces gnark marshal.go:311-348 ProvingKey.readFrom — a19)
//   * ChaCha20 witness assignment                        (replaces provers.go:79-142 + utils/bytes.go:11-47 — a3..a6)
//   * batched R1CS solver                                (replaces gnark constraint/bn254 solver.go:177-586 — a9, a10)
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
    Fr* w = W + (size_t)i * w_stride;
    const Fr one = Fr::one(), zero = Fr::zero();
    uint32_t pos = 0;
    w[pos++] = one;
    auto put_word = [&](uint32_t word) {
        for (int b = 0; b < 32; b++) w[pos++] = ((word >> b) & 1u) ? one : zero;
    };
    put_word(counters[i]);
    for (int k = 0; k < 3; k++) put_word(le32(nonce + 4 * k));
    for (int k = 0; k < 16; k++) put_word(be32(in + 4 * k));
    for (int k = 0; k < 16; k++) put_word(be32(ct + 4 * k));
    for (int k = 0; k < 8; k++) put_word(le32(key + 4 * k));
}

// generic: W[i][0] = 1, W[i][1..1+nw) = witness[i][..]
__global__ void witness_copy_kernel(const Fr* __restrict__ witness, uint32_t n_witness, uint32_t batch, Fr* __restrict__ W,
                                    size_t w_stride) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)batch * (n_witness + 1)) return;
    uint32_t i = (uint32_t)(gid / (n_witness + 1)), k = (uint32_t)(gid % (n_witness + 1));
    W[(size_t)i * w_stride + k] = k == 0 ? Fr::one() : witness[(size_t)i * n_witness + (k - 1)];
}

// ------------------------------------------------------------------------------------------------ R1CS solver
FD void acc_term(Fr& acc, const SolverProgram& sp, const Fr* W, uint32_t cid, uint32_t wid) {
    if (wid == WIRE_CONST) { acc = acc + sp.coeffs[cid]; return; }
    Fr w = W[wid];
    if (sp.fast_coeffs && cid <= 4) {   // uniform across the warp: all lanes run the same instruction
        switch (cid) {
            case 0: break;
            case 1: acc = acc + w; break;
            case 2: acc = acc + w.dbl(); break;
            case 3: acc = acc - w; break;
            default: acc = acc - w.dbl(); break;
        }
    } else {
        acc = acc + sp.coeffs[cid] * w;
    }
}
FD Fr eval_le(const SolverProgram& sp, const Fr* W, uint32_t& pos) {
    uint32_t nt = sp.calldata[pos++];
    Fr acc = Fr::zero();
    for (uint32_t t = 0; t < nt; t++) {
        uint32_t cid = sp.calldata[pos], wid = sp.calldata[pos + 1];
        pos += 2;
        acc_term(acc, sp, W, cid, wid);
    }
    return acc;
}

// executes instruction `ins` for one instance. status bits: 1 unsatisfied constraint, 2 division by zero, 4 unsupported
__device__ __forceinline__ void solve_instruction(const SolverProgram& sp, uint32_t ins, Fr* W, Fr* A, Fr* B, Fr* C, uint32_t* status) {
    const InsMeta m = sp.meta[ins];
    uint32_t kind = m.kind & 0xFF;
    uint32_t base = m.cd_start;
    if (kind == 0) {
        uint32_t n[3] = {sp.calldata[base + 1], sp.calldata[base + 2], sp.calldata[base + 3]};
        uint32_t pos = base + 4;
        uint32_t uside = (m.kind >> 8) & 0xFF;
        Fr sum[3], ucoef = Fr::zero();
        for (int side = 0; side < 3; side++) {
            Fr acc = Fr::zero();
            for (uint32_t t = 0; t < n[side]; t++) {
                uint32_t cid = sp.calldata[pos], wid = sp.calldata[pos + 1];
                pos += 2;
                if (wid == m.solve_wire && (uint32_t)side == uside) ucoef = ucoef + sp.coeffs[cid];
                else acc_term(acc, sp, W, cid, wid);
            }
            sum[side] = acc;
        }
        if (m.solve_wire != SOLVE_WIRE_NONE) {
            Fr w;
            Fr kinv = sp.ucoef_inv[ins];
            if (uside == 2) {
                w = (sum[0] * sum[1] - sum[2]) * kinv;
            } else {
                const Fr& other = sum[1 - uside];
                if (other.is_zero()) { atomicOr(status, 2u); w = Fr::zero(); }
                else w = (sum[2] * other.inv() - sum[uside]) * kinv;
            }
            W[m.solve_wire] = w;
            sum[uside] = sum[uside] + ucoef * w;
        } else if (sum[0] * sum[1] != sum[2]) {
            atomicOr(status, 1u);
        }
        A[m.cons_off] = sum[0];
        B[m.cons_off] = sum[1];
        C[m.cons_off] = sum[2];
    } else if (kind == 1) {
        uint32_t hid = sp.calldata[base + 1], nin = sp.calldata[base + 2];
        uint32_t pos = base + 3;
        if (hid == HINT_NBITS && nin == 1) {
            Fr v = eval_le(sp, W, pos).from_mont();
            uint32_t o0 = sp.calldata[pos], o1 = sp.calldata[pos + 1];
            const Fr one = Fr::one(), zero = Fr::zero();
            for (uint32_t k = 0; k < o1 - o0; k++) {
                uint32_t bit = k < 256 ? (v.l[k >> 5] >> (k & 31)) & 1u : 0u;
                W[o0 + k] = bit ? one : zero;
            }
        } else {
            atomicOr(status, 4u);
        }
    } else {
        // lookup: [len, nbEntries, nIn, inputs...] -> W[wire_off + k] = table[value(input k)]
        uint32_t nent = sp.calldata[base + 1], nin = sp.calldata[base + 2];
        uint32_t pos = base + 3;
        const Fr* tab = sp.lookup_tabs + (size_t)m.lookup_tab * 256;
        for (uint32_t k = 0; k < nin; k++) {
            Fr v = eval_le(sp, W, pos).from_mont();
            uint32_t hi = v.l[1] | v.l[2] | v.l[3] | v.l[4] | v.l[5] | v.l[6] | v.l[7];
            if (hi || v.l[0] >= nent || v.l[0] >= 256) { atomicOr(status, 1u); W[m.wire_off + k] = Fr::zero(); }
            else W[m.wire_off + k] = tab[v.l[0]];
        }
    }
}

// CTA = 32 instances (threadIdx.x) x SOLVER_WARPS instruction slots (threadIdx.y). All lanes of a warp execute the same
// instruction on different instances (no divergence); the warps of a CTA share out the instructions of a level and
// meet at a barrier between levels (the level schedule comes from the r1cs file).
__global__ void __launch_bounds__(32 * SOLVER_WARPS)
solver_kernel(SolverProgram sp, uint32_t batch, Fr* __restrict__ W, size_t w_stride, Fr* __restrict__ A, Fr* __restrict__ B,
              Fr* __restrict__ C, uint32_t* __restrict__ status) {
    uint32_t inst = blockIdx.x * 32 + threadIdx.x;
    bool active = inst < batch;
    Fr* w = W + (size_t)inst * w_stride;
    Fr* a = A + (size_t)inst * sp.n_dom;
    Fr* b = B + (size_t)inst * sp.n_dom;
    Fr* c = C + (size_t)inst * sp.n_dom;
    for (uint32_t lev = 0; lev < sp.nlevels; lev++) {
        uint32_t lo = sp.level_off[lev], hi = sp.level_off[lev + 1];
        if (active)
            for (uint32_t k = lo + threadIdx.y; k < hi; k += blockDim.y) solve_instruction(sp, sp.level_instr[k], w, a, b, c, status);
        __syncthreads();
    }
}

// per instruction: 1 / (sum of coefficients of the wire it solves)   (init-time)
__global__ void solver_ucoef_kernel(SolverProgram sp, uint32_t n_instr, Fr* __restrict__ out) {
    uint32_t ins = blockIdx.x * blockDim.x + threadIdx.x;
    if (ins >= n_instr) return;
    const InsMeta m = sp.meta[ins];
    Fr r = Fr::one();
    if ((m.kind & 0xFF) == 0 && m.solve_wire != SOLVE_WIRE_NONE) {
        uint32_t base = m.cd_start;
        uint32_t n[3] = {sp.calldata[base + 1], sp.calldata[base + 2], sp.calldata[base + 3]};
        uint32_t pos = base + 4, uside = (m.kind >> 8) & 0xFF;
        Fr uc = Fr::zero();
        for (int side = 0; side < 3; side++)
            for (uint32_t t = 0; t < n[side]; t++) {
                uint32_t cid = sp.calldata[pos], wid = sp.calldata[pos + 1];
                pos += 2;
                if (wid == m.solve_wire && (uint32_t)side == uside) uc = uc + sp.coeffs[cid];
            }
        r = uc.inv();
    }
    out[ins] = r;
}
// flag = 1 iff coefficient ids 0..4 are 0, 1, 2, -1, -2
__global__ void solver_check_coeffs_kernel(const Fr* __restrict__ coeffs, uint32_t ncoef, uint32_t* __restrict__ flag) {
    if (threadIdx.x || blockIdx.x) return;
    Fr one = Fr::one(), two = one + one;
    bool ok = ncoef >= 5 && coeffs[0].is_zero() && coeffs[1] == one && coeffs[2] == two && coeffs[3] == one.neg() &&
              coeffs[4] == two.neg();
    *flag = ok ? 1u : 0u;
}

// ------------------------------------------------------------------------------------------------ proof assembly
// One thread per proof:
//   Ar  = msmA + alpha + r*delta          Bs1 = msmB1 + beta + s*delta        Bs = msmB2 + beta2 + s*delta2
//   Krs = msmK + msmZ + s*Ar + r*Bs1 - (r*s)*delta
// rs: canonical limbs, r at [2i], s at [2i+1]. out: Proof.WriteTo bytes without commitments (164 B, Appendix C).
__global__ void __launch_bounds__(64)
assemble_kernel(AssemblyKeys keys, uint32_t n, const G1XYZZ* __restrict__ mA, const G1XYZZ* __restrict__ mB1,
                const G1XYZZ* __restrict__ mK, const G1XYZZ* __restrict__ mZ, const G2XYZZ* __restrict__ mB2,
                const Fr* __restrict__ rs, uint8_t* __restrict__ out, size_t out_stride) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr r = rs[2 * i], s = rs[2 * i + 1];
    Fr rsm = (r.to_mont() * s.to_mont()).from_mont();
    G1XYZZ d1 = G1XYZZ::from_affine(keys.delta);
    G1XYZZ Ar = mA[i];
    Ar.madd(keys.alpha, false);
    Ar.add(scalar_mul(d1, r));
    G1XYZZ Bs1 = mB1[i];
    Bs1.madd(keys.beta, false);
    Bs1.add(scalar_mul(d1, s));
    G1XYZZ Krs = mK[i];
    Krs.add(mZ[i]);
    Krs.add(scalar_mul(Ar, s));
    Krs.add(scalar_mul(Bs1, r));
    Krs.add(scalar_mul(d1, rsm).neg());
    uint8_t* o = out + (size_t)i * out_stride;
    g1_compress(Ar.to_affine(), o);
    g1_compress(Krs.to_affine(), o + 96);
    G2XYZZ Bs = mB2[i];
    Bs.madd(keys.beta2, false);
    Bs.add(scalar_mul(G2XYZZ::from_affine(keys.delta2), s));
    g2_compress(Bs.to_affine(), o + 32);
    o[128] = 0; o[129] = 0; o[130] = 0; o[131] = 0;
    o[132] = 0x40;
    for (int k = 133; k < 164; k++) o[k] = 0;
}

// gathers wire indices: out[j] = index of the j-th wire with flag[w] == 0   (host builds these; kept for reference)

}  // namespace g16
