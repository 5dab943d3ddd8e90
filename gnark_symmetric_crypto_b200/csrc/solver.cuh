// Batched R1CS witness solver: many independent witnesses of the SAME constraint system solved in lock-step.
// Replaces: gnark v0.11.0 constraint/bn254 (*system).Solve system.go:72-99, (*solver).run
// solver.go:439-509, processInstruction :390-418, solveR1C :540-586, solveWithHint :205-229, accumulateInto :177-195,
// blueprints blueprint_r1cs.go:36-59 / blueprint_hint.go:10-30 / blueprint_logderivlookup.go:30-69 and the hint
// std/math/bits.nBits — reached from libraries/prover/impl/provers.go:148,216 (groth16.Prove -> r1cs.Solve).
//
// Mapping: warp = one instruction x 32 witnesses (lanes). All lanes run the same instruction, so control flow is uniform;
// the calldata of the instruction is read once per warp through the uniform path. Linear expressions are walked four
// terms at a time: the four wire loads and the four coefficient products are independent, only the final additions
// chain (the 33-bit adder recompositions of the ChaCha circuit have 130-term expressions, and they bound the level time).
#pragma once
#include "prover_api.hpp"

namespace g16 {

static const int SOLVER_UNROLL = 4;

// acc += sum of `nt` terms starting at calldata[pos]; terms on `skip_wire` (the wire this instruction solves) are not
// evaluated: their coefficients are summed into ucoef instead.
// W points at this witness' column of the wire-major array: wire k is W[k * ws].
__device__ __forceinline__ void acc_terms(const SolverProgram& sp, const Fr* __restrict__ W, size_t ws, uint32_t pos,
                                          uint32_t nt, uint32_t skip_wire, Fr& acc, Fr& ucoef) {
    for (uint32_t t0 = 0; t0 < nt; t0 += SOLVER_UNROLL) {
        uint32_t cid[SOLVER_UNROLL], wid[SOLVER_UNROLL];
        Fr w[SOLVER_UNROLL];
#pragma unroll
        for (int u = 0; u < SOLVER_UNROLL; u++) {
            bool ok = t0 + u < nt;
            cid[u] = ok ? sp.calldata[pos + 2 * (t0 + u)] : 0u;
            wid[u] = ok ? sp.calldata[pos + 2 * (t0 + u) + 1] : SOLVE_WIRE_NONE - 1;   // sentinel: "no term"
        }
#pragma unroll
        for (int u = 0; u < SOLVER_UNROLL; u++) {
            bool is_wire = wid[u] < sp.n_wires && wid[u] != skip_wire;
            w[u] = is_wire ? W[(size_t)wid[u] * ws] : Fr::zero();
        }
        // products first (independent), sums afterwards
        Fr term[SOLVER_UNROLL];
#pragma unroll
        for (int u = 0; u < SOLVER_UNROLL; u++) {
            term[u] = Fr::zero();
            if (wid[u] == SOLVE_WIRE_NONE - 1) continue;                       // past the end
            if (wid[u] == WIRE_CONST) { term[u] = sp.coeffs[cid[u]]; continue; }   // constant term
            if (wid[u] == skip_wire) continue;
            if (sp.fast_coeffs && cid[u] <= 4) {   // warp-uniform: 0, +1, +2, -1, -2
                switch (cid[u]) {
                    case 0: break;
                    case 1: term[u] = w[u]; break;
                    case 2: term[u] = w[u].dbl(); break;
                    case 3: term[u] = w[u].neg(); break;
                    default: term[u] = w[u].dbl().neg(); break;
                }
            } else {
                // bit-valued wires dominate these circuits (every ChaCha wire is 0, 1 or -1): c*0 and c*1 need no product.
                // Exact for any value: the Montgomery product runs whenever a lane holds something else.
                const Fr c = sp.coeffs[cid[u]];
                if (w[u].is_zero()) term[u] = Fr::zero();
                else if (w[u] == Fr::one()) term[u] = c;
                else term[u] = c * w[u];
            }
        }
#pragma unroll
        for (int u = 0; u < SOLVER_UNROLL; u++) {
            if (wid[u] == skip_wire && skip_wire != SOLVE_WIRE_NONE) ucoef = ucoef + sp.coeffs[cid[u]];
            else acc = acc + term[u];
        }
    }
}
__device__ __forceinline__ Fr eval_le(const SolverProgram& sp, const Fr* __restrict__ W, size_t ws, uint32_t& pos) {
    uint32_t nt = sp.calldata[pos++];
    Fr acc = Fr::zero(), dummy = Fr::zero();
    acc_terms(sp, W, ws, pos, nt, SOLVE_WIRE_NONE, acc, dummy);
    pos += 2 * nt;
    return acc;
}

// Tail of an R1C once the three linear expressions are known without the solved wire's terms (their coefficients are
// summed in ucoef): solve for the wire (solver.go:540-586), or check L*R == O, and emit the constraint's A/B/C values.
__device__ __forceinline__ void r1c_finish(const SolverProgram& sp, const InsMeta& m, uint32_t ins, Fr sL, Fr sR, Fr sO,
                                           const Fr& ucoef, Fr* __restrict__ W, size_t ws, Fr* __restrict__ A,
                                           Fr* __restrict__ B, Fr* __restrict__ C, uint32_t* status) {
    uint32_t uside = (m.kind >> 8) & 0xFF;
    if (m.solve_wire != SOLVE_WIRE_NONE) {
        Fr w;
        Fr kinv = sp.ucoef_inv[ins];
        if (uside == 2) {
            w = (sL * sR - sO) * kinv;
            sO = sO + ucoef * w;
        } else if (uside == 0) {
            if (sR.is_zero()) { atomicOr(status, 2u); w = Fr::zero(); }
            else w = (sO * sR.inv() - sL) * kinv;
            sL = sL + ucoef * w;
        } else {
            if (sL.is_zero()) { atomicOr(status, 2u); w = Fr::zero(); }
            else w = (sO * sL.inv() - sR) * kinv;
            sR = sR + ucoef * w;
        }
        W[(size_t)m.solve_wire * ws] = w;
    } else if (sL * sR != sO) {
        atomicOr(status, 1u);
    }
    A[m.cons_off] = sL;
    B[m.cons_off] = sR;
    C[m.cons_off] = sO;
}

// executes instruction `ins` for one witness. status bits: 1 unsatisfied constraint, 2 division by zero, 4 unsupported
__device__ __forceinline__ void solve_instruction(const SolverProgram& sp, uint32_t ins, Fr* __restrict__ W, size_t ws,
                                                  Fr* __restrict__ A, Fr* __restrict__ B, Fr* __restrict__ C,
                                                  uint32_t* status) {
    const InsMeta m = sp.meta[ins];
    uint32_t kind = m.kind & 0xFF;
    uint32_t base = m.cd_start;
    if (kind == 0) {
        uint32_t nL = sp.calldata[base + 1], nR = sp.calldata[base + 2], nO = sp.calldata[base + 3];
        uint32_t uside = (m.kind >> 8) & 0xFF;
        bool solves = m.solve_wire != SOLVE_WIRE_NONE;
        Fr sL = Fr::zero(), sR = Fr::zero(), sO = Fr::zero(), ucoef = Fr::zero(), unused = Fr::zero();
        uint32_t pos = base + 4;
        acc_terms(sp, W, ws, pos, nL, (solves && uside == 0) ? m.solve_wire : SOLVE_WIRE_NONE, sL, (uside == 0) ? ucoef : unused);
        pos += 2 * nL;
        acc_terms(sp, W, ws, pos, nR, (solves && uside == 1) ? m.solve_wire : SOLVE_WIRE_NONE, sR, (uside == 1) ? ucoef : unused);
        pos += 2 * nR;
        acc_terms(sp, W, ws, pos, nO, (solves && uside == 2) ? m.solve_wire : SOLVE_WIRE_NONE, sO, (uside == 2) ? ucoef : unused);
        r1c_finish(sp, m, ins, sL, sR, sO, ucoef, W, ws, A, B, C, status);
    } else if (kind == 1) {
        uint32_t hid = sp.calldata[base + 1], nin = sp.calldata[base + 2];
        uint32_t pos = base + 3;
        if (hid == HINT_NBITS && nin == 1) {
            Fr v = eval_le(sp, W, ws, pos).from_mont();
            uint32_t o0 = sp.calldata[pos], o1 = sp.calldata[pos + 1];
            const Fr one = Fr::one(), zero = Fr::zero();
            uint32_t k = 0;
#pragma unroll
            for (int wi = 0; wi < 8; wi++) {   // static limb index: no local-memory array
                uint32_t word = v.l[wi];
                for (int b = 0; b < 32 && k < o1 - o0; b++, k++) W[(size_t)(o0 + k) * ws] = ((word >> b) & 1u) ? one : zero;
            }
            for (; k < o1 - o0; k++) W[(size_t)(o0 + k) * ws] = zero;
        } else if (hid == HINT_RANDOMIZE) {
            // hints.Randomize: one uniformly random field element (the blinding mask of the BSB22 commitment)
            for (uint32_t k = 0; k < nin; k++) eval_le(sp, W, ws, pos);   // (no inputs in these circuits)
            uint32_t o0 = sp.calldata[pos];
            if (sp.randomize) W[(size_t)o0 * ws] = sp.randomize[0];
            else atomicOr(status, 4u);
        } else if (hid == HINT_BSB22) {
            // executed between two level ranges by the host pipeline (MSM + hash): nothing to do here
            if (ins != sp.bsb_ins) atomicOr(status, 4u);
        } else if (hid == HINT_COUNT) {
            // logderivarg.countHint: inputs [nbRows, rowWidth, nbRows x rowWidth table values, queries (rowWidth each)];
            // output k = number of queries equal to row k. Rows carry their own index in column 0 (logderivlookup tables),
            // so a query can only match the row named by its first value.
            uint32_t nrows = eval_le(sp, W, ws, pos).from_mont().l[0];
            uint32_t width = eval_le(sp, W, ws, pos).from_mont().l[0];
            uint32_t o_pos;
            if (nrows > 256 || width != 2 || nin < 2 + nrows * width) { atomicOr(status, 4u); return; }
            uint16_t cnt[256];
            for (uint32_t k = 0; k < 256; k++) cnt[k] = 0;
            // position of every table row in calldata: rows are single constant terms in these circuits, but walk them anyway
            uint32_t row_pos0 = pos;
            uint32_t p2 = pos;
            for (uint32_t k = 0; k < nrows * width; k++) { uint32_t nt = sp.calldata[p2]; p2 += 1 + 2 * nt; }
            uint32_t row_stride_words = nrows ? (p2 - row_pos0) / nrows : 0;   // uniform rows (checked below)
            uint32_t nq = (nin - 2 - nrows * width) / width;
            pos = p2;
            for (uint32_t q = 0; q < nq; q++) {
                Fr qi = eval_le(sp, W, ws, pos);
                Fr qv = eval_le(sp, W, ws, pos);
                Fr qc = qi.from_mont();
                uint32_t hi = qc.l[1] | qc.l[2] | qc.l[3] | qc.l[4] | qc.l[5] | qc.l[6] | qc.l[7];
                if (hi || qc.l[0] >= nrows) continue;
                uint32_t rp = row_pos0 + qc.l[0] * row_stride_words;
                Fr ri = eval_le(sp, W, ws, rp);
                Fr rv = eval_le(sp, W, ws, rp);
                if (ri == qi && rv == qv) cnt[qc.l[0]]++;
            }
            o_pos = pos;
            uint32_t o0 = sp.calldata[o_pos], o1 = sp.calldata[o_pos + 1];
            for (uint32_t k = 0; k < o1 - o0; k++) {
                Fr v = Fr::zero();
                v.l[0] = k < 256 ? cnt[k] : 0;
                W[(size_t)(o0 + k) * ws] = v.to_mont();
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
            Fr v = eval_le(sp, W, ws, pos).from_mont();
            uint32_t hi = v.l[1] | v.l[2] | v.l[3] | v.l[4] | v.l[5] | v.l[6] | v.l[7];
            if (hi || v.l[0] >= nent || v.l[0] >= 256) { atomicOr(status, 1u); W[(size_t)(m.wire_off + k) * ws] = Fr::zero(); }
            else W[(size_t)(m.wire_off + k) * ws] = tab[v.l[0]];
        }
    }
}

// One launch per level of the r1cs schedule (instructions inside a level are independent, level l
// only reads wires of levels < l). grid.x covers the instructions of the level, grid.y the groups of 32 witnesses.
__global__ void __launch_bounds__(32 * SOLVER_WARPS)
solver_level_kernel(SolverProgram sp, uint32_t lo, uint32_t hi, uint32_t batch, Fr* W, size_t w_stride, Fr* A, Fr* B, Fr* C,
                    uint32_t* status) {
    uint32_t inst = blockIdx.y * 32 + threadIdx.x;
    uint32_t k = lo + blockIdx.x * blockDim.y + threadIdx.y;
    if (inst >= batch || k >= hi) return;
    if (sp.randomize) sp.randomize += inst;   // sp is a by-value kernel parameter: per-thread view of the mask array
    solve_instruction(sp, sp.level_instr[k], W + inst, w_stride, A + (size_t)inst * sp.n_dom,
                      B + (size_t)inst * sp.n_dom, C + (size_t)inst * sp.n_dom, status + inst);   // one status word per witness
}

#if !defined(G16_EMU)
// ------------------------------------------------------------------------------------------------ small batches
// Latency path (a single Prove request): with fewer witnesses than lanes the kernel above leaves
// 31 lanes idle and walks the 130-term adder expressions serially. Here a warp owns one (instruction, witness) pair and
// its lanes split the TERMS; the partial sums meet in a shuffle butterfly. Field addition is exact and commutative, so
// the result is bit-identical to the serial walk.
__device__ __forceinline__ Fr warp_sum(Fr v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Fr o;
#pragma unroll
        for (int i = 0; i < 8; i++) o.l[i] = __shfl_xor_sync(0xffffffffu, v.l[i], off);
        v = v + o;
    }
    return v;
}
__device__ __forceinline__ Fr term_value(const SolverProgram& sp, const Fr* __restrict__ W, size_t ws, uint32_t cid, uint32_t wid) {
    if (wid == WIRE_CONST) return sp.coeffs[cid];
    if (wid >= sp.n_wires) return Fr::zero();
    Fr w = W[(size_t)wid * ws];
    if (sp.fast_coeffs && cid <= 4) {
        switch (cid) {
            case 0: return Fr::zero();
            case 1: return w;
            case 2: return w.dbl();
            case 3: return w.neg();
            default: return w.dbl().neg();
        }
    }
    const Fr c = sp.coeffs[cid];
    if (w.is_zero()) return Fr::zero();
    if (w == Fr::one()) return c;
    return c * w;
}
// linear expression at calldata[pos] (nt | nt x (cid, wid)), terms split over the lanes; every lane returns the sum
__device__ __forceinline__ Fr eval_le_warp(const SolverProgram& sp, const Fr* __restrict__ W, size_t ws, uint32_t& pos, uint32_t lane) {
    uint32_t nt = sp.calldata[pos++];
    Fr acc = Fr::zero();
    for (uint32_t t = lane; t < nt; t += 32) acc = acc + term_value(sp, W, ws, sp.calldata[pos + 2 * t], sp.calldata[pos + 2 * t + 1]);
    pos += 2 * nt;
    return warp_sum(acc);   // lanes without a term hold zero
}
FD uint32_t fr_limb(const Fr& v, uint32_t i) {   // static indexing only (no local-memory array)
    uint32_t r = 0;
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) r = (i == k) ? v.l[k] : r;
    return r;
}
// one warp executes instruction `ins` of witness `inst`
__device__ __forceinline__ void solve_item_warp(SolverProgram sp, uint32_t ins, uint32_t inst, Fr* W, size_t w_stride, Fr* A, Fr* B,
                                                Fr* C, uint32_t* status) {
    const uint32_t lane = threadIdx.x;
    if (sp.randomize) sp.randomize += inst;
    W += inst;
    status += inst;   // one status word per witness
    A += (size_t)inst * sp.n_dom; B += (size_t)inst * sp.n_dom; C += (size_t)inst * sp.n_dom;
    const size_t ws = w_stride;
    const InsMeta m = sp.meta[ins];
    const uint32_t kind = m.kind & 0xFF, base = m.cd_start;
    if (kind == 0) {
        const uint32_t nL = sp.calldata[base + 1], nR = sp.calldata[base + 2], nO = sp.calldata[base + 3];
        const uint32_t uside = (m.kind >> 8) & 0xFF;
        const bool solves = m.solve_wire != SOLVE_WIRE_NONE;
        Fr sL = Fr::zero(), sR = Fr::zero(), sO = Fr::zero(), ucoef = Fr::zero();
        const uint32_t total = nL + nR + nO;
        for (uint32_t t = lane; t < total; t += 32) {
            const uint32_t side = t < nL ? 0u : (t < nL + nR ? 1u : 2u);
            const uint32_t cid = sp.calldata[base + 4 + 2 * t], wid = sp.calldata[base + 5 + 2 * t];
            if (solves && side == uside && wid == m.solve_wire) { ucoef = ucoef + sp.coeffs[cid]; continue; }
            const Fr v = term_value(sp, W, ws, cid, wid);
            if (side == 0) sL = sL + v;
            else if (side == 1) sR = sR + v;
            else sO = sO + v;
        }
        if (nL) sL = warp_sum(sL);
        if (nR) sR = warp_sum(sR);
        if (nO) sO = warp_sum(sO);
        if (solves) ucoef = warp_sum(ucoef);
        if (lane == 0) r1c_finish(sp, m, ins, sL, sR, sO, ucoef, W, ws, A, B, C, status);
    } else if (kind == 1 && sp.calldata[base + 1] == HINT_NBITS && sp.calldata[base + 2] == 1) {
        uint32_t pos = base + 3;
        const Fr v = eval_le_warp(sp, W, ws, pos, lane).from_mont();
        const uint32_t o0 = sp.calldata[pos], o1 = sp.calldata[pos + 1];
        const Fr one = Fr::one(), zero = Fr::zero();
        for (uint32_t b = lane; b < o1 - o0; b += 32) {
            const uint32_t word = b < 256 ? fr_limb(v, b >> 5) : 0u;
            W[(size_t)(o0 + b) * ws] = ((word >> (b & 31)) & 1u) ? one : zero;
        }
    } else if (kind == 1 && sp.calldata[base + 1] == HINT_COUNT) {
        // logderivarg.countHint (see solve_instruction): the queries are independent, so the lanes take every 32nd one and
        // count into shared memory. Every lane walks the calldata (the expressions have variable length) but evaluates only
        // its own queries. The AES circuits have five of these with hundreds of queries each: serial, they were most of
        // the solve time of a single proof.
        __shared__ uint32_t cnt_s[SOLVER_WARPS][256];
        uint32_t* cnt = cnt_s[threadIdx.y];
        for (uint32_t k2 = lane; k2 < 256; k2 += 32) cnt[k2] = 0;
        __syncwarp();
        const uint32_t nin = sp.calldata[base + 2];
        uint32_t pos = base + 3;
        const uint32_t nrows = eval_le(sp, W, ws, pos).from_mont().l[0];
        const uint32_t width = eval_le(sp, W, ws, pos).from_mont().l[0];
        if (nrows > 256 || width != 2 || nin < 2 + nrows * width) {
            if (lane == 0) atomicOr(status, 4u);
            return;
        }
        const uint32_t row_pos0 = pos;
        uint32_t p2 = pos;
        for (uint32_t k2 = 0; k2 < nrows * width; k2++) { uint32_t nt = sp.calldata[p2]; p2 += 1 + 2 * nt; }
        const uint32_t row_stride_words = nrows ? (p2 - row_pos0) / nrows : 0;
        const uint32_t nq = (nin - 2 - nrows * width) / width;
        pos = p2;
        for (uint32_t q = 0; q < nq; q++) {
            if ((q & 31u) != lane) {   // skip both expressions of a query that belongs to another lane
                uint32_t nt = sp.calldata[pos]; pos += 1 + 2 * nt;
                nt = sp.calldata[pos]; pos += 1 + 2 * nt;
                continue;
            }
            Fr qi = eval_le(sp, W, ws, pos);
            Fr qv = eval_le(sp, W, ws, pos);
            Fr qc = qi.from_mont();
            uint32_t hi = qc.l[1] | qc.l[2] | qc.l[3] | qc.l[4] | qc.l[5] | qc.l[6] | qc.l[7];
            if (hi || qc.l[0] >= nrows) continue;
            uint32_t rp = row_pos0 + qc.l[0] * row_stride_words;
            Fr ri = eval_le(sp, W, ws, rp);
            Fr rv = eval_le(sp, W, ws, rp);
            if (ri == qi && rv == qv) atomicAdd(&cnt[qc.l[0]], 1u);
        }
        __syncwarp();
        const uint32_t o0 = sp.calldata[pos], o1 = sp.calldata[pos + 1];
        for (uint32_t k2 = lane; k2 < o1 - o0; k2 += 32) {
            Fr v = Fr::zero();
            v.l[0] = k2 < 256 ? cnt[k2] : 0u;
            W[(size_t)(o0 + k2) * ws] = v.to_mont();
        }
    } else if (lane == 0) {
        solve_instruction(sp, ins, W, ws, A, B, C, status);   // lookups, Randomize: serial on one lane
    }
}
// grid.x covers the instructions of the level (blockDim.y warps per block), grid.y the witnesses
__global__ void __launch_bounds__(32 * SOLVER_WARPS)
solver_level_small_kernel(SolverProgram sp, uint32_t lo, uint32_t hi, uint32_t batch, Fr* W, size_t w_stride, Fr* A, Fr* B,
                          Fr* C, uint32_t* status) {
    const uint32_t inst = blockIdx.y;
    const uint32_t k = lo + blockIdx.x * blockDim.y + threadIdx.y;
    if (inst >= batch || k >= hi) return;   // warp-uniform
    solve_item_warp(sp, sp.level_instr[k], inst, W, w_stride, A, B, C, status);
}
// The same kernel for launches with programmatic stream serialisation (the small-batch graph of k_solver.cu): the grid may be
// scheduled while the level before it is still running — it says so to the level after it at once, fetches its instruction
// (program data, written at init), and only then waits for the level before it to complete and flush its wires. A level of one
// request is a few microseconds of work behind a few microseconds of launch latency; this hides the latter behind the former.
__global__ void __launch_bounds__(32 * SOLVER_WARPS)
solver_level_small_pdl_kernel(SolverProgram sp, uint32_t lo, uint32_t hi, uint32_t batch, Fr* W, size_t w_stride, Fr* A, Fr* B,
                              Fr* C, uint32_t* status) {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.launch_dependents;");
#endif
    const uint32_t inst = blockIdx.y;
    const uint32_t k = lo + blockIdx.x * blockDim.y + threadIdx.y;
    const bool live = inst < batch && k < hi;   // warp-uniform
    const uint32_t ins = live ? sp.level_instr[k] : 0u;
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
    if (!live) return;
    solve_item_warp(sp, ins, inst, W, w_stride, A, B, C, status);
}
// (Measured on B200 and not adopted: all levels in ONE cooperative launch, warps striding over the items of a level and
// meeting in grid.sync(). One ChaCha proof: 3.3 ms against 1.5 ms for one launch per level — back-to-back launches overlap
// the tail of a level with the head of the next, a grid-wide barrier cannot.)
#endif

// ------------------------------------------------------------------------------------------------ countHint, block-parallel
// Init: one thread per countHint instruction walks its calldata once and records where every query starts, so that the
// solve-time kernel can hand queries to threads directly. Header: nq (or ~0 = unsupported shape), row_pos0,
// row_stride_words, nrows; then nq + 1 positions (the last one is where the output range is stored).
__global__ void solver_count_index_kernel(SolverProgram sp, const uint32_t* __restrict__ ids, uint32_t n, uint32_t* __restrict__ out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const InsMeta m = sp.meta[ids[t]];
    uint32_t* ix = out + m.lookup_tab;
    const uint32_t base = m.cd_start, nin = sp.calldata[base + 2];
    uint32_t pos = base + 3;
    // nbRows and rowWidth must be constants (single constant term): they are evaluated without a witness
    uint32_t hdr[2];
    bool ok = true;
    for (int k = 0; k < 2; k++) {
        ok = ok && sp.calldata[pos] == 1 && sp.calldata[pos + 2] == WIRE_CONST;
        Fr v = ok ? sp.coeffs[sp.calldata[pos + 1]].from_mont() : Fr::zero();
        ok = ok && (v.l[1] | v.l[2] | v.l[3] | v.l[4] | v.l[5] | v.l[6] | v.l[7]) == 0;
        hdr[k] = v.l[0];
        pos += 1 + 2 * sp.calldata[pos];
    }
    const uint32_t nrows = hdr[0], width = hdr[1];
    if (!ok || nrows > 256 || width != 2 || nin < 2 + nrows * width) { ix[0] = 0xFFFFFFFFu; return; }
    const uint32_t row_pos0 = pos;
    for (uint32_t k = 0; k < nrows * width; k++) pos += 1 + 2 * sp.calldata[pos];
    const uint32_t nq = (nin - 2 - nrows * width) / width;
    ix[0] = nq; ix[1] = row_pos0; ix[2] = nrows ? (pos - row_pos0) / nrows : 0; ix[3] = nrows;
    for (uint32_t q = 0; q < nq; q++) {
        ix[4 + q] = pos;
        pos += 1 + 2 * sp.calldata[pos];
        pos += 1 + 2 * sp.calldata[pos];
    }
    ix[4 + nq] = pos;
}
#if !defined(G16_EMU)
static const int SOLVER_COUNT_THREADS = 256;
// grid.x = the countHint instructions [lo, hi) of a level, grid.y = witnesses; the threads of a block split the queries
__global__ void __launch_bounds__(SOLVER_COUNT_THREADS)
solver_count_kernel(SolverProgram sp, uint32_t lo, uint32_t hi, uint32_t batch, Fr* W, size_t ws, uint32_t* status) {
    __shared__ uint32_t cnt[256];
    const uint32_t tid = threadIdx.x, inst = blockIdx.y;
    const uint32_t ins = sp.level_instr[lo + blockIdx.x];
    W += inst;
    status += inst;
    const uint32_t* ix = sp.count_index + sp.meta[ins].lookup_tab;
    const uint32_t nq = ix[0], row_pos0 = ix[1], row_stride_words = ix[2], nrows = ix[3];
    if (nq == 0xFFFFFFFFu) {
        if (tid == 0) atomicOr(status, 4u);
        return;
    }
    cnt[tid] = 0;
    __syncthreads();
    for (uint32_t q = tid; q < nq; q += SOLVER_COUNT_THREADS) {
        uint32_t pos = ix[4 + q];
        Fr qi = eval_le(sp, W, ws, pos);
        Fr qv = eval_le(sp, W, ws, pos);
        Fr qc = qi.from_mont();
        uint32_t hi2 = qc.l[1] | qc.l[2] | qc.l[3] | qc.l[4] | qc.l[5] | qc.l[6] | qc.l[7];
        if (hi2 || qc.l[0] >= nrows) continue;
        uint32_t rp = row_pos0 + qc.l[0] * row_stride_words;
        Fr ri = eval_le(sp, W, ws, rp);
        Fr rv = eval_le(sp, W, ws, rp);
        if (ri == qi && rv == qv) atomicAdd(&cnt[qc.l[0]], 1u);
    }
    __syncthreads();
    const uint32_t pos = ix[4 + nq];
    const uint32_t o0 = sp.calldata[pos], o1 = sp.calldata[pos + 1];
    for (uint32_t k = tid; k < o1 - o0; k += SOLVER_COUNT_THREADS) {
        Fr v = Fr::zero();
        v.l[0] = k < 256 ? cnt[k] : 0u;
        W[(size_t)(o0 + k) * ws] = v.to_mont();
    }
}
#endif

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

}  // namespace g16
