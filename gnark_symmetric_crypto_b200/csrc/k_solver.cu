// Hot translation unit: the batched R1CS solver (solver.cuh) with the Montgomery product inlined, so that the four
// independent coefficient products of an unrolled term group overlap.
#include "solver.cuh"
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

namespace g16 {

static size_t launch_solver_levels(const SolverProgram& sp, const uint32_t* h_level_off, const uint32_t* h_level_split,
                                   const uint32_t* h_level_split2, uint32_t lev_begin, uint32_t lev_end, uint32_t batch, Fr* W, size_t w_stride, Fr* A, Fr* B, Fr* C,
                                   uint32_t* status, cudaStream_t st);

// Small batches replay the level launches from a CUDA graph: the sequence is fixed for a given (level range, batch, buffers),
// and a level of a single request is a few microseconds of work, so the per-launch overhead matters (163 launches for
// ChaCha, 441 for AES). The cache belongs to one prover context (one solver program). G16_SOLVER_GRAPH=0: plain launches.
#if !defined(G16_EMU)
typedef std::tuple<const void*, const void*, const void*, const void*, const void*, const void*, uint32_t, size_t, uint32_t, uint32_t>
    SolverGraphKey;
struct SolverGraph { cudaGraphExec_t exec; size_t launches; };
struct SolverGraphCache {
    std::map<SolverGraphKey, SolverGraph> graphs;
    std::mutex mu;
    void clear() {
        for (auto& kv : graphs) cudaGraphExecDestroy(kv.second.exec);
        graphs.clear();
    }
};
SolverGraphCache* solver_graph_cache_create() { return new SolverGraphCache(); }
void solver_graph_cache_destroy(SolverGraphCache* c) {
    if (!c) return;
    c->clear();
    delete c;
}
#else
struct SolverGraphCache {};
SolverGraphCache* solver_graph_cache_create() { return nullptr; }
void solver_graph_cache_destroy(SolverGraphCache*) {}
#endif

size_t launch_solver(const SolverProgram& sp, const uint32_t* h_level_off, const uint32_t* h_level_split,
                     const uint32_t* h_level_split2, uint32_t lev_begin, uint32_t lev_end, uint32_t batch, Fr* W, size_t w_stride, Fr* A, Fr* B, Fr* C, uint32_t* status,
                     cudaStream_t st, SolverGraphCache* cache) {
#if !defined(G16_EMU)
    static const uint32_t graph_max = [] { const char* v = getenv("G16_SOLVER_GRAPH"); return (uint32_t)(v && *v ? atoi(v) : 16); }();
    if (cache && batch <= graph_max && lev_end > lev_begin) {
        SolverGraphKey key(sp.randomize, W, A, B, C, status, batch, w_stride, lev_begin, lev_end);
        std::lock_guard<std::mutex> lk(cache->mu);
        auto& graphs = cache->graphs;
        auto it = graphs.find(key);
        if (it == graphs.end()) {
            if (graphs.size() >= 64) cache->clear();   // buffers were reallocated many times: start over
            cudaGraph_t graph = nullptr;
            SolverGraph sg{nullptr, 0};
            G16_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            try {
                sg.launches = launch_solver_levels(sp, h_level_off, h_level_split, h_level_split2, lev_begin, lev_end, batch, W, w_stride, A, B, C, status, st);
            } catch (...) {
                cudaStreamEndCapture(st, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            G16_CUDA(cudaStreamEndCapture(st, &graph));
            cudaError_t e = cudaGraphInstantiate(&sg.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) throw CudaError(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
            it = graphs.emplace(key, sg).first;
        }
        G16_CUDA(cudaGraphLaunch(it->second.exec, st));
        return it->second.launches;
    }
#endif
    return launch_solver_levels(sp, h_level_off, h_level_split, h_level_split2, lev_begin, lev_end, batch, W, w_stride, A, B, C, status, st);
}

static size_t launch_solver_levels(const SolverProgram& sp, const uint32_t* h_level_off, const uint32_t* h_level_split,
                                   const uint32_t* h_level_split2, uint32_t lev_begin, uint32_t lev_end, uint32_t batch, Fr* W, size_t w_stride, Fr* A, Fr* B, Fr* C,
                                   uint32_t* status, cudaStream_t st) {
    const uint32_t groups = div_up(batch, 32);
    size_t launches = 0;
    if (lev_end > sp.nlevels) lev_end = sp.nlevels;
#if !defined(G16_EMU)
    // few witnesses: every instruction goes to the term-parallel kernel (G16_SOLVER_SMALL = largest such batch)
    static const uint32_t small_max = [] { const char* v = getenv("G16_SOLVER_SMALL"); return (uint32_t)(v && *v ? atoi(v) : 16); }();
    const bool small = batch <= small_max;
#endif
    for (uint32_t lev = lev_begin; lev < lev_end; lev++) {
        uint32_t lo = h_level_off[lev], hi = h_level_off[lev + 1];
        uint32_t split = hi;   // [lo, split) witness-parallel, [split, hi) term-parallel, countHints [end, hi) block-parallel
#if !defined(G16_EMU)
        const uint32_t end = hi;
        if (h_level_split2 && sp.count_index) hi = h_level_split2[lev];
        if (small) split = lo;
        else if (h_level_split) split = h_level_split[lev];
        if (split > hi) split = hi;
#endif
        if (split > lo) {
            dim3 grid(div_up(split - lo, SOLVER_WARPS), groups);
            G16_LAUNCH(solver_level_kernel, grid, dim3(32, SOLVER_WARPS), 0, st, false, sp, lo, split, batch, W, w_stride, A, B, C,
                       status);
            launches++;
        }
#if !defined(G16_EMU)
        if (hi > split) {
            dim3 grid(div_up(hi - split, SOLVER_WARPS), batch);
            // small batches: programmatic stream serialisation, so that the launch of a level overlaps the level before it
            // (G16_SOLVER_PDL=0: plain launches)
            static const int pdl = [] { const char* v = getenv("G16_SOLVER_PDL"); return v && *v ? atoi(v) : 1; }();
            if (small && pdl) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = grid; cfg.blockDim = dim3(32, SOLVER_WARPS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                G16_CUDA(cudaLaunchKernelEx(&cfg, solver_level_small_pdl_kernel, sp, split, hi, batch, W, w_stride, A, B, C, status));
            } else {
                G16_LAUNCH(solver_level_small_kernel, grid, dim3(32, SOLVER_WARPS), 0, st, false, sp, split, hi, batch, W, w_stride, A,
                           B, C, status);
            }
            launches++;
        }
        if (end > hi) {
            G16_LAUNCH(solver_count_kernel, dim3(end - hi, batch), SOLVER_COUNT_THREADS, 0, st, false, sp, hi, end, batch, W, w_stride,
                       status);
            launches++;
        }
#endif
    }
    G16_CHECK_LAUNCH();
    return launches;
}
void launch_solver_count_index(const SolverProgram& sp, const uint32_t* ids, uint32_t n, uint32_t* out, cudaStream_t st) {
    if (!n) return;
    G16_LAUNCH(solver_count_index_kernel, div_up(n, 32), 32, 0, st, false, sp, ids, n, out);
    G16_CHECK_LAUNCH();
}
int launch_solver_init(const SolverProgram& sp, uint32_t n_instr, uint32_t n_coeffs, Fr* ucoef_inv, cudaStream_t st) {
    DevBuf<uint32_t> flag(1);
    G16_LAUNCH(solver_check_coeffs_kernel, 1, 1, 0, st, false, sp.coeffs, n_coeffs, flag.p);
    G16_LAUNCH(solver_ucoef_kernel, div_up(n_instr, 128), 128, 0, st, false, sp, n_instr, ucoef_inv);
    G16_CHECK_LAUNCH();
    uint32_t hf = 0;
    flag.download(&hf, 1, st);
    G16_CUDA(cudaStreamSynchronize(st));
    return (int)hf;
}

}  // namespace g16
