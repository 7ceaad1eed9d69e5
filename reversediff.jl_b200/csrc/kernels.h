// Kernel launchers of the replay engine (device code in gemm.cu / kernels.cu).  Everything is
// asynchronous on the given stream and returns the launch status.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>
#include "program.h"

namespace rdb {

// kernels launched by the most recent gemm_f64/gemm_f32 call on this thread (slicing passes + MMA kernel), for launch accounting
extern thread_local int g_gemm_launches;

// ---- (c) dense contractions -------------------------------------------------------------------------
cudaError_t gemm_f64(cudaStream_t, bool ta, bool tb, int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                     int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc, int mode = 0, int slices = 0);
cudaError_t gemm_f32(cudaStream_t, bool ta, bool tb, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                     int64_t lda, const float *B, int64_t ldb, float beta, float *C, int64_t ldc);

// TMA-fed tcgen05 path (umma.cu): FP32 via 3xTF32; returns cudaErrorNotSupported for problems too small to pay off
bool umma_available();
cudaError_t gemm_f32_umma(cudaStream_t, bool ta, bool tb, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                          int64_t lda, const float *B, int64_t ldb, float beta, float *C, int64_t ldc);

// Operand identity hints for the slice cache of the int8 path: a non-zero tag says "this operand's contents are identified by
// (tag, pointer, shape)"; the engine derives it from (tape uid, buffer id, value version).  0 = do not cache (adjoints).
// Consumed (reset to 0) by the next gemm_f64 call on this thread.
extern thread_local uint64_t g_tag_a, g_tag_b;
// "The operand handed to the next gemm_f64 is rows [row0, row0 + rows) of a larger operand that starts at `full` (same leading
// dimension, same contraction length), and the tag names that larger operand": the int8 path then slices the whole operand once
// and serves every panel of a panelled GEMM from it.  rows = the m (A) or n (B) index.  Consumed like the tags.
struct OperandWindow { const void *full = nullptr; int64_t rows_full = 0, row0 = 0; };
extern thread_local OperandWindow g_win_a, g_win_b;

// ---- per-context state of the GEMM layer ---------------------------------------------------------------
// Everything the `*` kernels keep between calls lives here (one per rdb_ctx, i.e. per device + stream), not in process globals:
// two contexts - two streams, two devices, two host threads - never share a workspace or a cache entry.  The engine points
// g_gemm at the context's state at every C-ABI entry; code that runs without a context (none in the product) gets a
// process-wide default instance.
struct SliceEntry {              // one operand of the int8 path, sliced: [slice][rows][Kp] int8 + per-row scales
    uint64_t tag;
    const void *ptr;
    int64_t ld, rows, K;
    bool kmajor;
    int ns, bits;
    cudaEvent_t ready;           // recorded on the stream that (re)filled the entry; readers on another stream wait for it
    cudaStream_t filled_on;
    int8_t *q;
    double *sc;
    size_t bytes;
    uint64_t last_use;
};
constexpr int kGuardSlots = 2048;
struct GuardSummary { unsigned long long checked, flagged; double worst_ratio; unsigned long long pairs; };   // pairs: slice products issued (after plane trimming), summed over launches
struct GemmState {
    void *ws = nullptr;          // grow-only workspace of the tcgen05 paths (slices of uncached operands, scales)
    size_t ws_bytes = 0;
    double *gemv_scratch = nullptr;
    size_t gemv_scratch_elems = 0;
    uint64_t alloc_gen = 0;      // bumped whenever ws / gemv_scratch move: a CUDA graph captured before is stale
    int *side_e = nullptr;       // exponent scratch of the slicing-ahead stream
    int64_t side_e_n = 0;
    std::vector<SliceEntry> cache;
    uint64_t use = 0;
    size_t cache_bytes = 0;
    bool no_cache = false;       // CUDA-graph tapes: every operand is sliced into the workspace on every replay (a cache hit
                                 // at capture time would freeze stale slices into the graph)
    // accuracy certificate of the int8-slice path (umma.cu, "guard"): per launch (max |alpha op(A) op(B)|, max error bound)
    // device buffers the engine KNOWS to hold one value in every element right now (deriv(x) after the adjoint of sum/mean
    // overwrote it with a fill): the slicer writes their digit planes from element 0 instead of reading the matrix (umma.cu
    // k_slice_uniform - the same digits, scales and plane word the generic passes produce).  The engine adds an entry when it
    // issues the fill and removes it before anything else may write the buffer (engine.cu: dtake / dmaterialize / sweep start).
    std::vector<const void *> uniform;
    void uniform_set(const void *p) { for (auto q : uniform) if (q == p) return; uniform.push_back(p); }
    void uniform_clear(const void *p) { for (size_t i = 0; i < uniform.size(); ++i) if (uniform[i] == p) { uniform[i] = uniform.back(); uniform.pop_back(); return; } }
    bool uniform_has(const void *p) const { for (auto q : uniform) if (q == p) return true; return false; }
    unsigned long long *guard = nullptr;      // device, kGuardSlots x 2 bit patterns of non-negative doubles
    GuardSummary *guard_sum = nullptr;        // device: accumulated by guard_fold
    GuardSummary *guard_host = nullptr;       // pinned host copy
    int guard_used = 0;
    bool guard_pending = false;               // folded results wait in guard_sum
    void release();
};
extern thread_local GemmState *g_gemm;
GemmState &gemm_state();                      // *g_gemm, or the process-wide default
// fold the slots used so far into the summary and reset them (stream-ordered); guard_read copies the summary to the host,
// synchronises the stream and clears it
cudaError_t guard_fold(cudaStream_t st, GemmState &G);
cudaError_t guard_read(cudaStream_t st, GemmState &G, GuardSummary *out);

// FP64 through exact int8 slices on the tcgen05 INT8 pipe (Ozaki scheme); cudaErrorNotSupported when not applicable
int ozaki_default_slices();
cudaError_t gemm_f64_ozaki(cudaStream_t, bool ta, bool tb, int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                           int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc, int nslices,
                           uint64_t tag_a = 0, uint64_t tag_b = 0, OperandWindow wa = OperandWindow(), OperandWindow wb = OperandWindow());

// slice an operand ahead of its GEMM on another stream (umma.cu); a no-op when the operand is not eligible
cudaError_t ozaki_preslice(cudaStream_t side, const double *src, int64_t ld, int64_t rows, int64_t K, bool kmajor, int nslices, uint64_t tag,
                           int *launched = nullptr);

// ---- (a)/(b) elementwise sweeps ---------------------------------------------------------------------
constexpr int kReduceBlocks = 1184;   // 148 SMs x 8 resident CTAs: partial-sum slots of the two-pass reductions

// Broadcast-aware argument of an elementwise instruction (bound trick of propagation.jl:25-27 expressed as
// strides: a size-1 dimension gets stride 0).
struct ExprArg {
    const void *ptr;
    int64_t stride[kMaxDims];
    int32_t same;   // 1: same shape as the output (offset == linear index)
    int32_t pad;
};

struct ExprParams {
    int32_t n_args, n_ops;
    int32_t order;                 // 1: value + first partials; 2: also second partials (Hessian)
    int32_t reduce;                // 1: also emit per-block partial sums of the output (fused `sum`)
    int64_t n;                     // output elements
    int64_t dims[kMaxDims];        // output dims
    ExprArg arg[kMaxArgs];
    void *out;
    void *partial[kMaxArgs];       // may be NULL (untracked argument)
    void *second[kMaxArgs * (kMaxArgs + 1) / 2];   // (p<=q) packed, order 2 only
    const int32_t *ops;            // device, 4 x int32 per op
    const double *fconst;          // device
    void *block_sums;              // kReduceBlocks partial sums when reduce == 1
};

// NVRTC-specialised elementwise forward kernel (jit.cu); jit_expr_kernel returns nullptr when NVRTC is unavailable or the
// program cannot be specialised (the interpreter launch_expr_forward is used then)
void *jit_expr_kernel(const ExprParams &p, const int32_t *host_ops, const double *host_fconst, bool f32, bool reduce);
cudaError_t jit_expr_launch(cudaStream_t, void *fn, const ExprParams &p, bool reduce);
// a run of adjacent same-shape elementwise instructions as one NVRTC kernel; from_item[a] >= 0: argument a is the value that item
// of the run computed for the same element (never re-read from memory)
struct JitRunItem { ExprParams p; const int32_t *ops; int from_item[kMaxArgs]; };
void *jit_run_kernel(const JitRunItem *items, int n_items, const double *host_fconst, bool f32, bool reduce);
cudaError_t jit_run_launch(cudaStream_t, void *fn, const JitRunItem *items, int n_items, void *block_sums, bool reduce);

template <typename T> cudaError_t launch_fill(cudaStream_t, T *p, int64_t n, T v);
// out = a + sign*b   (array +/-: arithmetic.jl:7-12,69-79); if block_sums != NULL also writes per-block sums
template <typename T> cudaError_t launch_add(cudaStream_t, T *out, const T *a, const T *b, T sign, int64_t n, T *block_sums);
// block_sums <- per-block sums of x  (sum / mean forward, reductions.jl:23-27,81-85)
template <typename T> cudaError_t launch_block_sums(cudaStream_t, const T *x, int64_t n, T *block_sums);
// out[0] = scale * sum(block_sums[0..kReduceBlocks))
template <typename T> cudaError_t launch_finish_sum(cudaStream_t, const T *block_sums, T *out, T scale, bool accumulate = false);
// block_sums <- per-block sums of a[i]*b[i]
template <typename T> cudaError_t launch_dot_block_sums(cudaStream_t, const T *a, const T *b, int64_t n, T *block_sums);
// dst[i + r*n] += delta[r]*src[i]
template <typename T> cudaError_t launch_scal_acc(cudaStream_t, T *dst, const T *src, const T *delta, int64_t n, int64_t R, bool overwrite = false);
// dst[i + r*n] += delta[clamp(i) + r*delta_n], clamp given by per-dimension strides of delta (0 on reduced dims)
template <typename T> cudaError_t launch_bcast_acc(cudaStream_t, T *dst, const int64_t *dims, const int64_t *delta_stride, const T *delta, int64_t n, int64_t R, int64_t delta_n, bool overwrite = false);
// dst (+)= sign*src over n elements (increment_deriv!/decrement_deriv!, propagation.jl:33-73)
template <typename T> cudaError_t launch_axpy(cudaStream_t, T *dst, const T *src, T sign, int64_t n, bool overwrite);
// dst[i + r*n] (+)= scale*delta[r]   (sum/mean reverse: scalar adjoint broadcast, reductions.jl:15-21,73-79)
template <typename T> cudaError_t launch_acc_scalar(cudaStream_t, T *dst, int64_t n, int64_t R, const T *delta, T scale, bool overwrite);
// elementwise instruction forward (broadcast.jl:196-204 ; elementwise.jl:323-343)
template <typename T> cudaError_t launch_expr_forward(cudaStream_t, const ExprParams &p);
// dst[i + r*n] (+)= delta[i + r*n]*partial[i]   (diffresult_increment_deriv!, propagation.jl:82-88)
template <typename T> cudaError_t launch_mul_acc(cudaStream_t, T *dst, const T *delta, const T *partial, int64_t n, int64_t R, bool overwrite);
// clamped accumulate into a column vector: dst[i + r*d0] += sum_j delta[i + j*d0 + r*d0*d1]*partial[i + j*d0]
// (propagation.jl:90-96 with bound (d0,1)); scratch holds d0*R*chunks partial sums
template <typename T> cudaError_t launch_mul_acc_colvec(cudaStream_t, T *dst, const T *delta, const T *partial, int64_t d0, int64_t d1, int64_t R, T *scratch, int64_t scratch_elems, bool overwrite = false);
// generic clamped accumulate (any broadcast pattern) with atomics
// launch_mul_acc of the same-shape argument and launch_mul_acc_colvec of the column-vector argument of ONE instruction in one pass over delta
template <typename T> cudaError_t launch_mul_acc_colvec_fused(cudaStream_t, T *dst_col, const T *delta, const T *partial_col, int64_t d0, int64_t d1, int64_t R, T *scratch, int64_t scratch_elems, bool overwrite_col, T *dst_same, const T *partial_same, bool overwrite_same, bool sdelta = false, T scale = (T)1);
// launch_mul_acc with delta = scale * dscalar[r] for every element (the adjoint a `sum`/`mean` would have filled in)
template <typename T> cudaError_t launch_mul_acc_sdelta(cudaStream_t, T *dst, const T *dscalar, T scale, const T *partial, int64_t n, int64_t R, bool overwrite);
template <typename T> cudaError_t launch_mul_acc_generic(cudaStream_t, T *dst, int64_t dst_n, const int64_t *dst_stride, const T *delta, const T *partial, const int64_t *dims, int64_t n, int64_t R);
// out[k] = src[map[k]]      (pull_value! of IDX operands, getindex forward: tracked.jl:178-181,331-340)
template <typename T> cudaError_t launch_gather(cudaStream_t, T *out, const T *src, const int64_t *map, int64_t n);
// dst[map[k] + r*dst_n] += src[k + r*n]   (scatter-add: propagation.jl:45-50, tracked.jl:318-329)
template <typename T> cudaError_t launch_scatter_add(cudaStream_t, T *dst, int64_t dst_n, const int64_t *map, const T *src, int64_t n, int64_t R);
// der[i + r*m] = (i == i0 + r)      (seed!(output, i) for a block of R rows, api/utils.jl:50)
template <typename T> cudaError_t launch_seed_onehot(cudaStream_t, T *der, int64_t m, int64_t R, int64_t i0);
// result[r + j*ld] = xder[j + r*n]  (result_matrix[i, j] = input_deriv[j], api/utils.jl:52-54, for R rows at once)
template <typename T> cudaError_t launch_rows_out(cudaStream_t, T *result, int64_t ld, const T *xder, int64_t n, int64_t R);

// ---- forward-over-reverse (Hessian) tangent sweeps ---------------------------------------------------
// tv_out[i + k*n] = sum_p partial_p[i]*tv_p[i + k*n]
template <typename T> cudaError_t launch_tangent_forward(cudaStream_t, T *tv_out, int n_args, const T *const *partial, const T *const *tv_in, int64_t n, int64_t K);
// td_p[i+k*n] (+)= dt[i+k*n]*partial_p[i] + delta[i]*sum_q second_pq[i]*tv_q[i+k*n]   (dt may be NULL = known zero)
template <typename T> cudaError_t launch_tangent_reverse(cudaStream_t, T *td_p, const T *dt, const T *delta, const T *partial_p, int n_args, const T *const *second_pq, const T *const *tv_in, int64_t n, int64_t K, bool overwrite);
// all arguments are getindex views of the seeded input: td_x[map_p[i] + k*n_x] += delta[i]*second_pq[i] at k = map_q[i] - c0
// hessian! of f(x) = sum|mean(g.(views of x...)) in ONE kernel (engine.cu: SepHess): column c of the block is zero-filled and the few
// second partials that land in it are evaluated from x on the spot (inverse index: which (element, argument) pairs read x[c]);
// one thread applies a column's contributions in tape order, so the result does not depend on scheduling
struct SepHessParams {
    int n_args, n_ops;
    const int64_t *map[kMaxArgs];   // device; NULL = the argument is x itself
    const int64_t *inv_off;         // n_x + 1
    const int64_t *inv;             // packed element * n_args + argument, ascending per column
    const int32_t *ops;             // device, 4 x int32 per op
    const double *fconst;
    const void *x;
    int64_t n_x;
    void *H;
    int64_t ld, c0, K;
    double delta;                   // adjoint of every element of g's output: 1 (sum) or 1 / length (mean)
};
constexpr int kSepHessMaxEntries = 32;
template <typename T> cudaError_t launch_hessian_separable(cudaStream_t, const SepHessParams &);
// the same sweeps for closures whose arguments broadcast against the output (index maps from strides; atomics into clamped arguments)
template <typename T> cudaError_t launch_tangent_forward_g(cudaStream_t, T *tv_out, int n_args, const T *const *partial, const T *const *tv_in, const int64_t (*stride)[kMaxDims], const int64_t *sizes, const int64_t *dims, int64_t n, int64_t K);
template <typename T> cudaError_t launch_tangent_reverse_g(cudaStream_t, T *td_p, const T *dt, const T *delta, const T *partial_p, int n_args, int p, const T *const *second_pq, const T *const *tv_in, const int64_t (*stride)[kMaxDims], const int64_t *sizes, const int64_t *dims, int64_t n, int64_t K);
template <typename T> cudaError_t launch_tangent_reverse_sparse(cudaStream_t, T *td_x, int64_t n_x, int64_t K, int64_t c0, const T *delta, int n_args, const int64_t *const *maps, const T *const *second_pq /* [p*n_args+q] */, int64_t n);
// rows gather/scatter for getindex under tangents: out[k + c*n] = src[map[k] + c*src_n]
template <typename T> cudaError_t launch_gather_cols(cudaStream_t, T *out, const T *src, int64_t src_n, const int64_t *map, int64_t n, int64_t K);
// tangent of getindex applied to the seeded tape input: tv[k + c*n] = (map[k] == c0 + c)
template <typename T> cudaError_t launch_seed_gather(cudaStream_t, T *tv, const int64_t *map, int64_t n, int64_t K, int64_t c0);
// tv[i + k*n] = (i == c0 + k)
template <typename T> cudaError_t launch_seed_tangent(cudaStream_t, T *tv, int64_t n, int64_t K, int64_t c0);
// out[k] = sum_i x[i + k*n]   (column sums; tangent of sum)
template <typename T> cudaError_t launch_colsum(cudaStream_t, T *out, const T *x, int64_t n, int64_t K, T scale);
// 2-D strided copy: dst[i + k*ldd] = src[i + k*lds]
template <typename T> cudaError_t launch_copy2d(cudaStream_t, T *dst, int64_t ldd, const T *src, int64_t lds, int64_t n, int64_t K);

}  // namespace rdb

// ---- scalar-instruction blocks (scalar.cu) ------------------------------------------------------------
// A maximal run of consecutive ScalarInstructions (tape.jl:30-45; derivatives/scalars.jl:52-123) replayed as one unit.
// Host planner + device tables; wire layout in program.py (ScalarBlock).
#include <string>
#include <vector>
namespace rdb {

struct ScalarSeg { int lv0, lv1, grid; };     // a launch: levels [lv0, lv1) — one wide level on many CTAs, or a run of narrow levels on ONE CTA

struct ScalarBlock {
    int64_t n = 0, n_leaf = 0, n_edges = 0, n_exports = 0, n_rslots = 0;
    int n_ref = 0, n_flevels = 0, n_rlevels = 0;
    // many-seed reverse sweep (tape-order streaming plan, scalar.cu): schedule blocks of 32 positions, adjoint rows that must
    // live in global memory (read >= 2 blocks after they are produced), and planner statistics
    int64_t n_sblocks = 0, n_grow = 0, n_near = 0, n_staged = 0, n_direct = 0;
    size_t der_bytes = 0;
    // bundle ("VLIW") plans for deep graphs: forward, and reverse with R < 16 seeds (scalar.cu)
    int64_t n_fbundles = 0, n_rbundles = 0, n_lit = 0, n_brow = 0;
    bool fwd_bundles = false, rev_bundles = false;
    std::vector<int32_t> refbufs;               // tape buffer ids referenced (leaf origins and export targets)
    std::vector<char> ref_has_leaf, ref_has_export;
    std::vector<int64_t> ref_export_count;      // elements of the buffer this block exports into
    std::vector<ScalarSeg> fsegs, rsegs;        // forward / single-seed reverse launch plans
    std::vector<void *> h_ext_val, h_ext_der;   // per ref: base pointers of value(buffer) / deriv(buffer) (engine-owned)
    std::vector<int64_t> h_ext_size;
    int64_t R_alloc = 0;
    // device tables (owned; freed by scalar_block_free)
    void *d_tables = nullptr;                   // one allocation holding everything below except the pools
    int4 *d_finstr = nullptr;                   // level-sorted: (instr, a_code, b_code, expr)   a/b: >= 0 slot, < -1 literal, INT_MIN none
    int64_t *d_flevel_off = nullptr;
    int4 *d_ops = nullptr;                      // concatenated bytecode of the block's scalar functions
    int4 *d_exprs = nullptr;                    // (offset, n_ops, n_args, 0)
    int32_t *d_epos = nullptr;                  // 2n: where forward stores partial (i, a|b) in the reverse edge array, or -1
    int32_t *d_leaf_ord = nullptr; int64_t *d_leaf_elem = nullptr;
    int32_t *d_exp_slot = nullptr, *d_exp_ord = nullptr; int64_t *d_exp_elem = nullptr;
    int32_t *d_rslot = nullptr;                 // reverse-level-sorted slots: >= 0 instruction output, < 0 leaf ~l
    int64_t *d_reoff = nullptr;                 // n_rslots + 1 edge offsets
    int32_t *d_esrc = nullptr;                  // consumer instruction of every edge, reference accumulation order
    int64_t *d_rlevel_off = nullptr;
    int64_t *d_xoff = nullptr; int32_t *d_xord = nullptr; int64_t *d_xelem = nullptr;   // exports of instruction i (CSR over i)
    void **d_ext_val = nullptr, **d_ext_der = nullptr; int64_t *d_ext_size = nullptr;
    void *d_val = nullptr;                      // T[n_leaf + n] slot values
    void *d_epart = nullptr;                    // T[n_edges] partials in reverse edge order
    int4 *d_recA = nullptr, *d_recB = nullptr, *d_recC = nullptr, *d_recD = nullptr;   // per schedule position (scalar.cu, "streaming plan")
    int2 *d_pf = nullptr;                       // per schedule block: what to prefetch into the staging buffer
    int4 *d_vfA = nullptr, *d_vfB = nullptr, *d_vfC = nullptr, *d_vfD = nullptr, *d_vrA = nullptr, *d_vrB = nullptr;
    int32_t *d_bpos = nullptr;                  // 2n: where forward stores partial (i, a|b) in the bundle-ordered copy, or -1
    void *d_bpart = nullptr;                    // T[2 x 32 x bundles]: partials of edges 0/1 in reverse-bundle order   // bundle records
    int32_t *d_vsrc = nullptr;                  // bundle reverse: operand sources of every edge (same order as d_esrc)
    int32_t *d_tsrc = nullptr;                  // operand codes of every edge for the streaming kernel (same order as d_esrc)
    void *d_der = nullptr;                      // adjoints of instruction outputs: T[n x R] at [i*R + r] (R < 16), T[n_grow x R] (R >= 16)
};

// Parse + validate + plan.  buf_size[b] < 0 marks a buffer that cannot be referenced (CONST).  Returns "" or an error text.
std::string scalar_block_build(ScalarBlock &B, const int64_t *payload, int64_t len, const int32_t *expr, uint64_t n_expr,
                               uint64_t n_fconst, const std::vector<int64_t> &buf_size, size_t es, cudaStream_t s);
void scalar_block_free(ScalarBlock &B);
cudaError_t scalar_block_upload_ext(ScalarBlock &B, cudaStream_t s);
cudaError_t scalar_block_ensure_R(ScalarBlock &B, int64_t R, size_t es);
bool scalar_reverse_streams(const ScalarBlock &B, int64_t R, size_t es);
int64_t scalar_schedule_bundles_host(int64_t n_nodes, const int64_t *succ_off, const int32_t *succ, int32_t *bundle, int32_t *lane);   // R >= 16 seeds: streaming kernel (true) or one bundle warp per seed
// forward: pull leaf values, evaluate level by level, mirror exports; returns the number of kernel launches in *launches
template <typename T> cudaError_t scalar_block_forward(cudaStream_t, ScalarBlock &B, const double *fconst, int *launches);
// reverse for R seeds (deriv layout of tape buffers: [elem + r*size])
// leaves_zero: every leaf origin's deriv is logically zero on entry (pull_deriv! would load zeros)
template <typename T> cudaError_t scalar_block_reverse(cudaStream_t, ScalarBlock &B, int64_t R, bool leaves_zero, int *launches);

}  // namespace rdb
