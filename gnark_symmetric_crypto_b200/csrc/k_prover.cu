// Cold translation unit (built with -DG16_COLD): key decompression, witness assignment, R1CS solver, proof assembly and
// the stage-level field / group test kernels. The Montgomery product stays out of line here (see field.cuh FD_MUL).
#include "prover_kernels.cuh"
#include "msm_types.hpp"
#include "team.cuh"

namespace g16 {

__global__ void field_op_kernel(int field, int op, const uint64_t* __restrict__ a, const uint64_t* __restrict__ b,
                                uint64_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (field == 0) {
        Fp x = ((const Fp*)a)[i], y = b ? ((const Fp*)b)[i] : Fp::zero(), r;
        switch (op) {
            case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break; case 3: r = x.inv(); break;
            case 4: r = x.sqr(); break; case 5: r = x.neg(); break; case 6: r = x.to_mont(); break; default: r = x.from_mont();
        }
        ((Fp*)out)[i] = r;
    } else {
        Fr x = ((const Fr*)a)[i], y = b ? ((const Fr*)b)[i] : Fr::zero(), r;
        switch (op) {
            case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break; case 3: r = x.inv(); break;
            case 4: r = x.sqr(); break; case 5: r = x.neg(); break; case 6: r = x.to_mont(); break; default: r = x.from_mont();
        }
        ((Fr*)out)[i] = r;
    }
}
// op 0: a + b via madd ; 1: k*a (k = 8 x u32 canonical limbs in b) ; 2: 2a ; 3: a + b via the general XYZZ addition
template <class C>
__global__ void group_op_kernel(int op, const typename C::A* __restrict__ a, const uint64_t* __restrict__ b,
                                typename C::A* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typedef typename C::X X;
    X acc = X::from_affine(a[i]);
    if (op == 0) {
        acc.madd(((const typename C::A*)b)[i], false);
    } else if (op == 1) {
        Scalar256 k;
        for (int j = 0; j < 8; j++) k.w[j] = ((const uint32_t*)b)[8 * i + j];
        X res = scalar_mul(acc, k);
        acc = res;
    } else if (op == 2) {
        acc = acc.dbl();
    } else {
        X o = X::from_affine(((const typename C::A*)b)[i]);
        // make the second operand non-trivially projective: (x,y,1,1) -> (4x, 8y, 4, 8)
        if (!o.is_inf()) {
            typename C::F two = C::F::one().dbl(), four = two.dbl(), eight = four.dbl();
            o.X = o.X * four; o.Y = o.Y * eight; o.ZZ = four; o.ZZZ = eight;
        }
        acc.add(o);
    }
    out[i] = acc.to_affine();
}

// the four-warp forms of team.cuh (G1): op 4: a + b with b made projective as in op 3 ; op 5: 4a + 2b (two team doublings of a,
// one of b, one team addition: both operands projective). Block = one team, lane = instance.
__global__ void __launch_bounds__(128)
group_op_team_kernel(int op, const G1Affine* __restrict__ a, const G1Affine* __restrict__ b, G1Affine* __restrict__ out, size_t n) {
    __shared__ Fp sm[TEAM4_SM_ELEMS];
    __shared__ uint32_t flag;
    Team4 T{sm, &flag, (int)(threadIdx.x >> 5), (int)(threadIdx.x & 31), 0};
    if (threadIdx.x == 0) flag = 0u;
    __syncthreads();
    const size_t i = (size_t)blockIdx.x * 32 + T.lane;
    const bool live = i < n;
    G1XYZZ acc = live ? G1XYZZ::from_affine(a[i]) : G1XYZZ::inf();
    G1XYZZ o = live ? G1XYZZ::from_affine(b[i]) : G1XYZZ::inf();
    if (op == 4) {
        if (!o.is_inf()) {
            Fp two = Fp::one().dbl(), four = two.dbl(), eight = four.dbl();
            o.X = o.X * four; o.Y = o.Y * eight; o.ZZ = four; o.ZZZ = eight;
        }
        acc = team_add(T, acc, o, false);
    } else {
        acc = team_dbl(T, team_dbl(T, acc));
        o = team_dbl(T, o);
        acc = team_add(T, acc, o, false);
    }
    if (live && T.w == 0) out[i] = acc.to_affine();
}

void launch_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n, cudaStream_t st) {
    G16_LAUNCH(field_op_kernel, div_up(n, 128), 128, 0, st, false, field, op, a, b, out, n);
    G16_CHECK_LAUNCH();
}
void launch_group_op(int group, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n, cudaStream_t st) {
    if (group == 1 && op >= 4) {
        G16_LAUNCH(group_op_team_kernel, div_up(n, 32), 128, 0, st, true, op, (const G1Affine*)a, (const G1Affine*)b, (G1Affine*)out, n);
    } else if (group == 1) {
        auto k = group_op_kernel<G1>;
        G16_LAUNCH(k, div_up(n, 64), 64, 0, st, false, op, (const G1Affine*)a, b, (G1Affine*)out, n);
    } else {
        auto k = group_op_kernel<G2>;
        G16_LAUNCH(k, div_up(n, 64), 64, 0, st, false, op, (const G2Affine*)a, b, (G2Affine*)out, n);
    }
    G16_CHECK_LAUNCH();
}

void launch_decompress_g1(const uint8_t* in, uint32_t n, G1Affine* out, uint32_t* err, cudaStream_t st) {
    G16_LAUNCH(decompress_g1_kernel, div_up(n, 128), 128, 0, st, false, in, n, out, err);
    G16_CHECK_LAUNCH();
}
void launch_decompress_g2(const uint8_t* in, uint32_t n, G2Affine* out, uint32_t* err, cudaStream_t st) {
    G16_LAUNCH(decompress_g2_kernel, div_up(n, 64), 64, 0, st, false, in, n, out, err);
    G16_CHECK_LAUNCH();
}
void launch_scalars_from_be(const uint8_t* in, uint32_t n, Fr* out, cudaStream_t st) {
    G16_LAUNCH(scalars_from_be_kernel, div_up(n, 128), 128, 0, st, false, in, n, out);
    G16_CHECK_LAUNCH();
}
void launch_chacha_witness(const uint8_t* keys, const uint8_t* nonces, const uint32_t* counters, const uint8_t* inputs,
                           uint32_t n, Fr* W, size_t w_stride, uint8_t* ct_out, cudaStream_t st) {
    G16_LAUNCH(chacha_witness_kernel, div_up(n, 64), 64, 0, st, false, keys, nonces, counters, inputs, n, W, w_stride, ct_out);
    G16_CHECK_LAUNCH();
}
void launch_witness_copy(const Fr* witness, uint32_t n_witness, uint32_t batch, Fr* W, size_t w_stride, cudaStream_t st) {
    G16_LAUNCH(witness_copy_kernel, div_up((size_t)batch * (n_witness + 1), 256), 256, 0, st, false, witness, n_witness,
               batch, W, w_stride);
    G16_CHECK_LAUNCH();
}
void launch_wires_to_rows(const Fr* W, size_t w_stride, uint32_t batch, uint32_t nb_wires, Fr* out, cudaStream_t st) {
    G16_LAUNCH(wires_to_rows_kernel, div_up((size_t)batch * nb_wires, 256), 256, 0, st, false, W, w_stride, batch, nb_wires, out);
    G16_CHECK_LAUNCH();
}
void launch_fixed_base_tables(const AssemblyKeys& keys, G1Affine* tab1, G2Affine* tab2, cudaStream_t st) {
    launch_fixed_base_table_g1(keys.delta, tab1, st);
    auto k2 = fixed_base_table_kernel<G2>;
    G16_LAUNCH(k2, 1, FB_WINDOWS, 0, st, false, keys.delta2, tab2);
    G16_CHECK_LAUNCH();
}
void launch_verify_unpack(const VerifyKeys& keys, const uint8_t* proofs, size_t stride, uint32_t n, G1Affine* P, G2Affine* Q,
                          G1Affine* P2, G2Affine* Q2, G1Affine* commit, const void* frob, uint32_t* bad, cudaStream_t st) {
    G16_LAUNCH(verify_unpack_kernel, dim3(div_up(n, 32), keys.n_commit ? 5 : 3), 32, 0, st, false, keys, proofs, stride, n, P, Q, P2, Q2,
               commit, (const Fp2*)frob, bad);
    G16_CHECK_LAUNCH();
}
void launch_g2_subgroup(const G2Affine* pts, uint32_t n, const void* frob, uint8_t* ok, cudaStream_t st) {
    G16_LAUNCH(g2_subgroup_kernel, div_up(n, 32), 32, 0, st, false, pts, n, (const Fp2*)frob, ok);
    G16_CHECK_LAUNCH();
}
void launch_verify_ksum(const G1XYZZ* msm, const G1Affine* commit, uint32_t n, G1Affine* P, cudaStream_t st) {
    G16_LAUNCH(verify_ksum_kernel, div_up(n, 32), 32, 0, st, false, msm, commit, n, P);
    G16_CHECK_LAUNCH();
}
void launch_verify_verdict(const uint8_t* ok1, const uint8_t* ok2, const uint32_t* bad, uint32_t n, uint8_t* out, cudaStream_t st) {
    G16_LAUNCH(verify_verdict_kernel, div_up(n, 128), 128, 0, st, false, ok1, ok2, bad, n, out);
    G16_CHECK_LAUNCH();
}
void launch_assemble_g2(const AssemblyKeys& keys, uint32_t n, const G2XYZZ* mB2, const Fr* rs, uint8_t* out, size_t out_stride,
                        cudaStream_t st) {
    G16_LAUNCH(assemble_g2_kernel, n, FB_WINDOWS, 0, st, true, keys, n, mB2, rs, out, out_stride);
    G16_CHECK_LAUNCH();
}
void launch_aes_witness(const uint8_t* keys, uint32_t key_len, const uint8_t* nonces, const uint32_t* counters,
                        const uint8_t* inputs, uint32_t n, Fr* W, size_t w_stride, uint8_t* ct_out, cudaStream_t st) {
    G16_LAUNCH(aes_witness_kernel, div_up(n, 64), 64, 0, st, false, keys, key_len, nonces, counters, inputs, n, W, w_stride, ct_out);
    G16_CHECK_LAUNCH();
}
void launch_bsb22_challenge(const G1XYZZ* commit, uint32_t n, Fr* W, size_t w_stride, uint32_t commit_wire, G1Affine* commit_aff,
                            cudaStream_t st) {
    G16_LAUNCH(bsb22_challenge_kernel, div_up(n, 64), 64, 0, st, false, commit, n, W, w_stride, commit_wire, commit_aff);
    G16_CHECK_LAUNCH();
}
void launch_g1_affine_to_xyzz(const G1Affine* in, uint32_t n, G1XYZZ* out, cudaStream_t st) {
    G16_LAUNCH(g1_affine_to_xyzz_kernel, div_up(n, 128), 128, 0, st, false, in, n, out);
    G16_CHECK_LAUNCH();
}
void launch_assemble_commitment(const G1Affine* commit_aff, const G1XYZZ* pok, uint32_t n, uint8_t* out, size_t out_stride,
                                cudaStream_t st) {
    G16_LAUNCH(assemble_commitment_kernel, div_up(n, 64), 64, 0, st, false, commit_aff, pok, n, out, out_stride);
    G16_CHECK_LAUNCH();
}
}  // namespace g16
