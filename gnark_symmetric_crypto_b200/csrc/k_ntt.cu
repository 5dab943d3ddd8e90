// Hot translation unit: NTT passes, compute_h, and the integer-multiply microbenchmarks (field product inlined).
#include "ntt.cuh"

namespace g16 {

static std::vector<NttPass> ntt_plan(int k, bool dif) {
    std::vector<NttPass> ps;
    int m0 = k < NTT_MAX_TILE_LG ? k : NTT_MAX_TILE_LG;
    NttPass c;
    c.k = k; c.b_lo = 0; c.m = m0; c.q = 0; c.lg_tile = m0; c.dif = dif;
    int rem = k - m0;
    const int npass = (rem + 7) / 8;
    std::vector<NttPass> strided;
    int b = m0;
    for (int i = 0; i < npass; i++) {
        int m = rem / npass + (i < rem % npass ? 1 : 0);   // even split
        NttPass s;
        s.k = k; s.b_lo = b; s.m = m; s.q = NTT_MAX_TILE_LG - m; s.lg_tile = NTT_MAX_TILE_LG; s.dif = dif;
        if (s.q > b) { s.q = b; s.lg_tile = s.m + s.q; }
        strided.push_back(s);
        b += m;
    }
    if (dif) {
        for (int i = (int)strided.size() - 1; i >= 0; i--) ps.push_back(strided[i]);
        ps.push_back(c);
    } else {
        ps.push_back(c);
        for (auto& s : strided) ps.push_back(s);
    }
    return ps;
}

void ntt_domain_init(NttDomain& d, int k, const Fr& w, const Fr& g, cudaStream_t stream) {
    d.k = k;
    d.n = 1u << k;
    DevBuf<Fr> consts(5);
    G16_LAUNCH(ntt_domain_consts_kernel, 1, 1, 0, stream, false, w, g, d.n, consts.p);
    G16_CHECK_LAUNCH();
    Fr h[5];
    consts.download(h, 5, stream);
    G16_CUDA(cudaStreamSynchronize(stream));
    Fr winv = h[0], ninv = h[1], ginv = h[2], ninv_den = h[4];
    d.den = h[3];
    uint32_t half = d.n > 1 ? d.n / 2 : 1;
    d.tw_fwd.alloc(half);
    d.tw_inv.alloc(half);
    G16_LAUNCH(ntt_powers_kernel, div_up(half, 128), 128, 0, stream, false, w, half, d.tw_fwd.p);
    G16_LAUNCH(ntt_powers_kernel, div_up(half, 128), 128, 0, stream, false, winv, half, d.tw_inv.p);
    d.scale_ninv.alloc(d.n);
    d.scale_coset_fwd.alloc(d.n);
    d.scale_coset_inv.alloc(d.n);
    d.scale_coset_only.alloc(d.n);
    d.scale_ninv_den.alloc(d.n);
    d.scale_coset_inv_den.alloc(d.n);
    G16_LAUNCH(ntt_coset_table_kernel, div_up(d.n, 128), 128, 0, stream, false, Fr::one(), ninv, k, d.scale_ninv.p);
    G16_LAUNCH(ntt_coset_table_kernel, div_up(d.n, 128), 128, 0, stream, false, g, ninv, k, d.scale_coset_fwd.p);
    G16_LAUNCH(ntt_coset_table_kernel, div_up(d.n, 128), 128, 0, stream, false, ginv, ninv, k, d.scale_coset_inv.p);
    G16_LAUNCH(ntt_coset_table_kernel, div_up(d.n, 128), 128, 0, stream, false, g, Fr::one(), k, d.scale_coset_only.p);
    G16_LAUNCH(ntt_coset_table_kernel, div_up(d.n, 128), 128, 0, stream, false, Fr::one(), ninv_den, k, d.scale_ninv_den.p);
    G16_LAUNCH(ntt_coset_table_kernel, div_up(d.n, 128), 128, 0, stream, false, ginv, ninv_den, k, d.scale_coset_inv_den.p);
    G16_CHECK_LAUNCH();
    d.dif_passes = ntt_plan(k, true);
    d.dit_passes = ntt_plan(k, false);
    G16_CUDA(cudaStreamSynchronize(stream));
}

// 2^28-th root of unity of Fr, big-endian canonical
static const uint8_t ROOT28_BE[32] = {0x2a, 0x3c, 0x09, 0xf0, 0xa5, 0x8a, 0x7e, 0x85, 0x00, 0xe0, 0xa7, 0xeb, 0x8e, 0xf6, 0x2a, 0xbc,
                                      0x40, 0x2d, 0x11, 0x1e, 0x41, 0x11, 0x2e, 0xd4, 0x9b, 0xd6, 0x1b, 0x6e, 0x72, 0x5b, 0x19, 0xf0};

void ntt_standalone_domain(NttDomain& d, int k, cudaStream_t st) {
    DevBuf<uint8_t> rb;
    rb.upload(ROOT28_BE, 32, st);
    DevBuf<Fr> wg(2);
    G16_LAUNCH(ntt_root_kernel, 1, 1, 0, st, false, rb.p, k, wg.p);
    G16_CHECK_LAUNCH();
    Fr h[2];
    wg.download(h, 2, st);
    G16_CUDA(cudaStreamSynchronize(st));
    ntt_domain_init(d, k, h[0], h[1], st);
}

static void ntt_set_smem_attr() {
#if !defined(G16_EMU)
    // per device: the attribute belongs to the function in the current context
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !done[dev]) {
        G16_CUDA(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((size_t)32 << NTT_MAX_TILE_LG)));
        done[dev] = true;
    }
#endif
}

void ntt_run(NttDomain& d, Fr* data, size_t vec_stride, uint32_t batch, bool dif, bool inverse_root, const Fr* scale,
             cudaStream_t stream) {
    ntt_set_smem_attr();
    const std::vector<NttPass>& ps = dif ? d.dif_passes : d.dit_passes;
    const Fr* tw = inverse_root ? d.tw_inv.p : d.tw_fwd.p;
    for (const NttPass& p : ps) {
        if (p.m == 0) continue;
        dim3 grid(d.n >> p.lg_tile, batch);
        size_t smem = (size_t)32 << p.lg_tile;
        const Fr* sc = (p.b_lo == 0) ? scale : nullptr;
        G16_LAUNCH(ntt_pass_kernel, grid, NTT_THREADS, smem, stream, true, data, vec_stride, p, tw, sc);
        d.launches++;
    }
    G16_CHECK_LAUNCH();
}

// H = (A.B - C) / (x^n - 1) with SIX transforms instead of gnark's seven (prove.go:359-384 runs C through the coset as well).
// The witness satisfies the system (the solver checked it), so C interpolates a_i b_i on the domain: C = A.B mod (x^n - 1).
// Writing A.B = P_lo + x^n P_hi gives C = P_lo + P_hi and H = P_hi. On the coset g.w^i, x^n = g^n =: G, so the products
// d_i = A(g w^i) B(g w^i) are the evaluations of E = P_lo + G P_hi, and H = (E - C) / (G - 1) coefficient by coefficient:
//   coefficients of C (one inverse transform, no trip through the coset)  and  coefficients of E (inverse coset transform of d).
// Same polynomial, hence bit-identical coefficients; den = 1/(G - 1) and 1/n are folded into the scale tables of the two
// inverse transforms, so the last step is a subtraction.
void compute_h_run(NttDomain& d, Fr* a, Fr* b, Fr* c, size_t vec_stride, uint32_t batch, cudaStream_t stream) {
    if (vec_stride != d.n) throw std::invalid_argument("compute_h: vectors must be contiguous (stride == n)");
    Fr* v[2] = {a, b};
    for (int i = 0; i < 2; i++) {
        ntt_run(d, v[i], vec_stride, batch, true, true, d.scale_coset_fwd.p, stream);    // iNTT, then * g^i / n
        ntt_run(d, v[i], vec_stride, batch, false, false, nullptr, stream);              // evaluate on the coset
    }
    ntt_run(d, c, vec_stride, batch, true, true, d.scale_ninv_den.p, stream);            // den * coefficients of C (bit-reversed)
    size_t total = (size_t)batch * vec_stride;
    G16_LAUNCH(h_mul_kernel, div_up(total, 256), 256, 0, stream, false, a, (const Fr*)b, total);
    ntt_run(d, a, vec_stride, batch, true, true, d.scale_coset_inv_den.p, stream);       // den * coefficients of E (bit-reversed)
    G16_LAUNCH(h_sub_kernel, div_up(total, 256), 256, 0, stream, false, a, (const Fr*)c, total);
    d.launches += 2;
}

// The first four transforms of compute_h_run and the pointwise product: on return `a` holds d_i = A(g w^i) B(g w^i) in natural
// order; b is clobbered. The Z query can consume d (and the untouched evaluations of C) directly over the evaluation-basis
// tables built by ctx_build_eval_tables, so H is never materialised on that path.
void compute_d_run(NttDomain& d, Fr* a, Fr* b, size_t vec_stride, uint32_t batch, cudaStream_t stream) {
    if (vec_stride != d.n) throw std::invalid_argument("compute_d: vectors must be contiguous (stride == n)");
    Fr* v[2] = {a, b};
    for (int i = 0; i < 2; i++) {
        ntt_run(d, v[i], vec_stride, batch, true, true, d.scale_coset_fwd.p, stream);
        ntt_run(d, v[i], vec_stride, batch, false, false, nullptr, stream);
    }
    size_t total = (size_t)batch * vec_stride;
    G16_LAUNCH(h_mul_kernel, div_up(total, 256), 256, 0, stream, false, a, (const Fr*)b, total);
    d.launches += 1;
    G16_CHECK_LAUNCH();
}
// Column j0 + r of the linear map (d, c) -> h of compute_h_run, one column per row of `out` (rows x n, contiguous):
// which = 0: h as a function of d (inverse coset transform, den/n folded in); which = 1: minus h as a function of the
// evaluations of C (inverse transform) -- the sign is folded into the unit vector.
void compute_h_columns(NttDomain& d, Fr* out, uint32_t rows, uint32_t j0, int which, cudaStream_t stream) {
    G16_CUDA(cudaMemsetAsync(out, 0, (size_t)rows * d.n * sizeof(Fr), stream));
    G16_LAUNCH(unit_rows_kernel, div_up(rows, 128), 128, 0, stream, false, out, d.n, rows, j0, which);
    ntt_run(d, out, d.n, rows, true, true, which ? d.scale_ninv_den.p : d.scale_coset_inv_den.p, stream);
}

void ntt_bitrev(const Fr* in, Fr* out, int k, cudaStream_t stream) {
    G16_LAUNCH(ntt_bitrev_kernel, div_up((size_t)1 << k, 256), 256, 0, stream, false, in, out, k);
    G16_CHECK_LAUNCH();
}
void fr_be_to_mont(const uint8_t* d_in, uint32_t n, Fr* d_out, cudaStream_t stream) {
    G16_LAUNCH(fr_be_to_mont_kernel, div_up(n, 64), 64, 0, stream, false, d_in, n, d_out);
    G16_CHECK_LAUNCH();
}
void ntt_fill_pattern(Fr* out, size_t total, cudaStream_t stream) {
    G16_LAUNCH(fill_pattern_kernel, div_up(total, 256), 256, 0, stream, false, out, total);
    G16_CHECK_LAUNCH();
}
void ntt_count_mismatches(const Fr* a, const Fr* b, size_t total, uint32_t* d_count, cudaStream_t stream) {
    G16_LAUNCH(count_mismatch_kernel, div_up(total, 256), 256, 0, stream, false, a, b, total, d_count);
    G16_CHECK_LAUNCH();
}

// ------------------------------------------------------------------------------------------------ integer-multiply peak
#if !defined(G16_EMU)
template <int WIDE>
__global__ void __launch_bounds__(256) imad_peak_kernel(uint32_t* out, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 7;
    if (WIDE) {
        uint64_t c0 = a, c1 = b, c2 = a ^ b, c3 = a + b, c4 = a * 3, c5 = b * 5, c6 = a * 7, c7 = b * 9;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 16; u++) {
                asm volatile("mad.wide.u32 %0, %8, %9, %0;\n\tmad.wide.u32 %1, %8, %9, %1;\n\t"
                             "mad.wide.u32 %2, %8, %9, %2;\n\tmad.wide.u32 %3, %8, %9, %3;\n\t"
                             "mad.wide.u32 %4, %8, %9, %4;\n\tmad.wide.u32 %5, %8, %9, %5;\n\t"
                             "mad.wide.u32 %6, %8, %9, %6;\n\tmad.wide.u32 %7, %8, %9, %7;"
                             : "+l"(c0), "+l"(c1), "+l"(c2), "+l"(c3), "+l"(c4), "+l"(c5), "+l"(c6), "+l"(c7)
                             : "r"(a), "r"(b));
            }
        }
        uint64_t s = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
        if (s == 0x1234567ull) out[0] = (uint32_t)s;
    } else {
        uint32_t c0 = a, c1 = b, c2 = a ^ b, c3 = a + b, c4 = a * 3, c5 = b * 5, c6 = a * 7, c7 = b * 9;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 16; u++) {
                asm volatile("mad.lo.u32 %0, %8, %9, %0;\n\tmad.lo.u32 %1, %8, %9, %1;\n\t"
                             "mad.lo.u32 %2, %8, %9, %2;\n\tmad.lo.u32 %3, %8, %9, %3;\n\t"
                             "mad.lo.u32 %4, %8, %9, %4;\n\tmad.lo.u32 %5, %8, %9, %5;\n\t"
                             "mad.lo.u32 %6, %8, %9, %6;\n\tmad.lo.u32 %7, %8, %9, %7;"
                             : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7)
                             : "r"(a), "r"(b));
            }
        }
        uint32_t s = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
        if (s == 0x1234567u) out[0] = s;
    }
}
// carry-chain wide MADs (mad.lo.cc / madc.hi.cc pairs = IMAD.WIDE.U32.X with a predicate carry), four independent
// 8-limb accumulators per thread: the instruction mix of the Montgomery product itself
__global__ void __launch_bounds__(256) imad_chain_peak_kernel(uint32_t* out, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 7;
    uint32_t acc[4][8];
    for (int k = 0; k < 4; k++) for (int j = 0; j < 8; j++) acc[k][j] = a + 17 * k + j;
    uint32_t sink = 0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int k = 0; k < 4; k++) sink += mad4(acc[k], a, b, a ^ b, a + b, b + u);
        }
    }
    uint32_t s = sink;
    for (int k = 0; k < 4; k++) for (int j = 0; j < 8; j++) s ^= acc[k][j];
    if (s == 0x1234567u) out[0] = s;
}
// two independent dependent-chains of Fp products per thread
__global__ void __launch_bounds__(256) modmul_peak_kernel(Fp* out, int iters) {
    Fp a = Fp::one(), b = Fp::r2(), c = Fp::one(), d = Fp::r2();
    a.l[0] += threadIdx.x;
    b.l[0] += blockIdx.x;
    c.l[1] += threadIdx.x;
    d.l[1] += blockIdx.x;
    for (int i = 0; i < iters; i++) {
        a = a * b;
        c = c * d;
        b = b * a;
        d = d * c;
    }
    if (a.l[0] == 0x12345u && b.l[3] == 77u && c.l[1] == d.l[2]) out[0] = a;
}
#endif

static double g_chain_rate = 0;
double imad_chain_rate() { return g_chain_rate; }
void imad_peak_measure(double* imad_per_s, double* imad_wide_per_s, double* modmul_per_s) {
#if defined(G16_EMU)
    throw std::runtime_error("imad peak needs a GPU");
#else
    DevBuf<uint32_t> out(64);
    cudaEvent_t e0, e1;
    G16_CUDA(cudaEventCreate(&e0));
    G16_CUDA(cudaEventCreate(&e1));
    const int blocks = (int)sm_count() * 8, threads = 256, iters = 2000;
    auto time_it = [&](auto launch) {
        launch();   // warm-up
        G16_CUDA(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            G16_CUDA(cudaEventRecord(e0, 0));
            launch();
            G16_CUDA(cudaEventRecord(e1, 0));
            G16_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        return (double)best * 1e-3;
    };
    double ops = (double)blocks * threads * iters * 16.0 * 8.0;
    double t0 = time_it([&] { imad_peak_kernel<0><<<blocks, threads>>>(out.p, iters); });
    double t1 = time_it([&] { imad_peak_kernel<1><<<blocks, threads>>>(out.p, iters); });
    // 4 unrolled x 4 chains x 4 wide MADs per iteration
    double t3 = time_it([&] { imad_chain_peak_kernel<<<blocks, threads>>>(out.p, iters); });
    g_chain_rate = (double)blocks * threads * iters * 64.0 / t3;
    const int mm_iters = 400;
    double t2 = time_it([&] { modmul_peak_kernel<<<blocks, threads>>>((Fp*)out.p, mm_iters); });
    G16_CHECK_LAUNCH();
    *imad_per_s = ops / t0;
    *imad_wide_per_s = ops / t1;
    *modmul_per_s = (double)blocks * threads * mm_iters * 4.0 / t2;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
#endif
}

}  // namespace g16
