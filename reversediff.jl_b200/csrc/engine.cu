// Host side of librdb200.so: tape-program loader/validator, buffer arena, forward/reverse executors, batched
// seed sweeps for jacobian!/hessian!, and the C ABI of include/rdb200.h.
//
// Reference semantics followed here (all paths relative to the ReverseDiff.jl tree):
//   pass loops              src/tape.jl:75-90, src/api/tape.jl:122-134
//   seeding / extraction    src/api/utils.jl:27-58,77-100 ; src/tracked.jl:202-213
//   `*`                     src/derivatives/linalg/arithmetic.jl:245-285
//   array +/-               src/derivatives/linalg/arithmetic.jl:45-61,127-154
//   sum / mean              src/derivatives/linalg/reductions.jl:15-27,73-85
//   elementwise closures    src/derivatives/broadcast.jl:163-204 ; src/derivatives/elementwise.jl:323-408,466-537
//   accumulators            src/derivatives/propagation.jl:33-108
//   getindex                src/tracked.jl:318-340
// Every reverse exec ends with unseed!(output) (e.g. arithmetic.jl:51,266; reductions.jl:19; broadcast.jl:174).
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "kernels.h"
#include "program.h"
#include "rdb200.h"

using namespace rdb;

// ------------------------------------------------------------------------------------------ error state
static thread_local char g_err[1024] = "";
static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
// for the other translation units of the library (comm.cu)
int rdb_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            cudaGetLastError(); /* do not leave a stale error for the next, unrelated call */             \
            return fail(RDB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                                \
    } while (0)
#define OK(call)                 \
    do {                         \
        int rc__ = (call);       \
        if (rc__ != RDB_OK) return rc__; \
    } while (0)

// ------------------------------------------------------------------------------------------ structures
struct rdb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int refs = 1;          // the handle itself + one per live tape: destruction order is the caller's business
    int gemm_f64_mode = 0, ozaki_slices = 0;   // for rdb_gemm on this context
    cudaStream_t copy_stream = nullptr;        // device->host result copies that overlap the tail of the reverse sweep
    cudaStream_t upload_stream = nullptr;      // host->device input copies in column panels (rdb_tape_set_input), see settle_uploads
    cudaStream_t side_stream = nullptr;        // operand slicing for the NEXT `*` GEMM while the current one runs (preslice_*)
    cudaEvent_t side_ev = nullptr;
    GemmState gemm;                            // workspace, sliced-operand cache, accuracy-certificate slots of the `*` kernels
};
cudaStream_t rdb_ctx_stream(rdb_ctx *c) { return c->stream; }
int rdb_ctx_device(rdb_ctx *c) { return c->device; }
// every C-ABI entry: the calling thread works on this context's device with this context's GEMM state
static void ctx_enter(rdb_ctx *c) {
    cudaSetDevice(c->device);
    g_gemm = &c->gemm;
    c->gemm.uniform.clear();     // "this deriv buffer is a constant fill" is knowledge of ONE sweep inside ONE call: another call, another
                                 // tape of the context, or a freed and reused address must never find it
}
static void ctx_release(rdb_ctx *c) {
    if (!c || --c->refs > 0) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->side_stream) cudaStreamSynchronize(c->side_stream);
    c->gemm.release();
    if (g_gemm == &c->gemm) g_gemm = nullptr;
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->upload_stream) cudaStreamDestroy(c->upload_stream);
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    if (c->side_ev) cudaEventDestroy(c->side_ev);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

struct Buf {
    int kind = 0, nd = 0;
    int64_t dims[kMaxDims] = {1, 1, 1, 1};
    int64_t size = 1;
    int alias_of = -1;
    void *val = nullptr;   // value(x)
    void *der = nullptr;   // deriv(x), size x R (batched seeds)
    void *tv = nullptr;    // value tangents  (size x K), hessian! only
    void *td = nullptr;    // deriv tangents  (size x K), hessian! only
    int n_consumers = 0;
    uint32_t version = 1;  // bumped whenever value(x) is rewritten: identifies sliced-operand cache entries
    // a large host input still arriving on the upload stream in column panels [up_c0[i], up_c0[i+1]): event i = panel i landed
    cudaEvent_t up_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t up_c0[5] = {0, 0, 0, 0, 0};
    int up_n = 0;
};

enum OperandMode { OPM_PLAIN = 0, OPM_FOLDED_T = 1, OPM_GATHER = 2 };

struct OperandPlan {
    OperandRec rec{};
    int mode = OPM_PLAIN;
    int64_t size = 1;           // elements of the operand as the instruction sees it
    int64_t rows = 1, cols = 1; // matrix view (vectors are cols == 1; row vectors rows == 1)
    int64_t *d_map = nullptr;   // device index map (OPM_GATHER, and getindex)
    void *dense_val = nullptr;  // densified value(a) for OPM_GATHER operands
    void *dense_der = nullptr;  // adjoint temporary for OPM_GATHER operands
};

struct InstrPlan {
    InstrRec rec{};
    OperandPlan in[kMaxArgs];
    int32_t *d_ops = nullptr;
    void *partial[kMaxArgs] = {nullptr, nullptr, nullptr, nullptr};
    void *second[kMaxArgs * (kMaxArgs + 1) / 2] = {nullptr};
    bool fused_into_prev = false;   // forward: this `sum`/`mean` was produced by the previous kernel
    std::unique_ptr<ScalarBlock> blk;   // OP_SCALAR_BLOCK: planner output + device tables (scalar.cu)
    std::vector<int32_t> h_ops;     // host copy of the scalar program (NVRTC specialisation)
    void *jit_fn[2] = {nullptr, nullptr};   // [reduce]
    bool jit_tried[2] = {false, false};
    int partial_alias[kMaxArgs] = {-1, -1, -1, -1, -1, -1, -1, -1};   // >= 0: this argument's partial is provably the SAME value as that earlier argument's: stored once
    int run_end = -2;                    // elementwise run fusion: last instruction of the run that starts here (-2 = not planned, own index = none)
    void *run_fn[2] = {nullptr, nullptr};  // [reduce]: the run's NVRTC kernel
    bool run_tried[2] = {false, false};
    bool rev_done[2] = {false, false};   // `*`: this operand's adjoint was already accumulated in this sweep (paired panels, mul_reverse)
    uint64_t dtag = 0;                   // `*`: identity of deriv(output) for the sliced-operand cache, fixed when first needed in a sweep
};

struct StepStat {
    std::string name;
    double ms = 0, bytes = 0, flops = 0;
    int count = 0;
};

static uint32_t g_next_tape_uid = 1;
struct rdb_tape {
    uint32_t uid = g_next_tape_uid++;
    rdb_ctx *ctx = nullptr;
    int dtype = RDB_F64;
    size_t es = 8;
    std::vector<Buf> bufs;
    std::vector<InstrPlan> instrs;
    std::vector<int> inputs;
    int output = -1;
    rdb_options opts{};
    std::vector<void *> allocs;
    double *d_fconst = nullptr;
    std::vector<double> h_fconst;
    void *scratch = nullptr;       // reduction scratch
    int64_t scratch_bytes = 0;
    int64_t R_alloc = 0;           // deriv buffers hold size x R_alloc
    int64_t K_alloc = 0;           // tangent buffers hold size x K_alloc
    void *stage = nullptr;         // device staging for host results
    int64_t stage_bytes = 0;
    int64_t launches = 0;
    bool second_order = false;     // partial/second arrays allocated for hessian!
    std::vector<StepStat> stats;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool forward_valid = false;
    // hessian! with a device result and no gradient/value request, launch-bound tapes: the whole sweep as one graph, keyed by its arguments
    struct HessGraph { cudaGraphExec_t exec = nullptr; int state = 0; int64_t cb = 0, ce = 0, ld = 0; void *result = nullptr; } hgraph;
    bool capturing = false;        // hessian_impl is being captured: no synchronisation, no host copies
    // accuracy certificate of the int8-slice `*` kernels (umma.cu guard): totals since load, and the fallback state
    int64_t guard_checked = 0, guard_flagged = 0, guard_fallbacks = 0, slice_pairs = 0;
    double guard_worst = 0;
    // rdb_gradient_async: the call is only enqueued; rdb_tape_wait completes it (and runs the deferred accuracy check of the int8-slice GEMMs)
    struct Pending { bool active = false; std::vector<void *> results; int on_device = 0; void *value_out = nullptr; bool guard = false; } pending;
    bool async_call = false;       // inside rdb_gradient_async: gradient_impl records done_ev instead of synchronising
    cudaEvent_t done_ev = nullptr;
    bool async = false;            // rdb_tape_set_async: calls with device results and no host outputs return without synchronising
    // hessian! of sum|mean(g.(views of x)) with a device result and nothing else requested: one fused kernel (kernels.cu
    // k_hessian_separable); state 0 = not examined yet, 1 = planned, -1 = the tape does not have that shape
    struct SepHess { int state = 0; int ew = -1; double delta = 1.0; const int64_t *map[kMaxArgs] = {nullptr}; int64_t *d_inv_off = nullptr, *d_inv = nullptr;
                     const void *lazy_x = nullptr; bool x_consumed = false; } sep;
    int hess_prefill = 0;          // hessian!: the result block's zero-fill runs on a side stream next to the small kernels (1), not (-1), unknown (0)
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    int guard_streak = 0;          // consecutive calls whose certificate failed; 3 make the tape DMMA-only for good
    bool dmma_now = false;         // this call is the DMMA replay of a call whose certificate failed
    uint64_t graph_gen = 0, hgraph_gen = 0;   // GemmState::alloc_gen when the graphs were captured
    int graph_guard_n = 0, hgraph_guard_n = 0; // certificate slots the captured kernels write
    int last_upload = -1;          // buffer of the most recent panelled upload (the one worth consuming panel by panel)
    // deriv zero-state, per ROOT buffer: dz[b] != 0 means deriv(b) is LOGICALLY zero while its storage may hold anything.
    // unseed! (every reverse exec ends with one, tracked.jl:207-213) only sets the flag; the next accumulation into the buffer
    // overwrites (beta = 0, plain store) instead of memset + read-modify-write; anything that cannot overwrite (scatter, atomics)
    // and every reader outside the sweep (extract_result!, rdb_get_deriv, hessian!) materialises the zeros first.
    std::vector<char> dz;
    // whole-sweep CUDA graph for launch-bound tapes (the 100x100 config is ~25 launches of a few microseconds each)
    cudaGraphExec_t gexec = nullptr;
    int graph_state = 0;           // 0 = first call runs eagerly (warms every lazy allocation), 1 = capture next, 2 = ready, -1 = off
    // early result extraction (gradient! with host results): deriv(input) is copied out as soon as its last accumulation of
    // the reverse sweep has been issued, overlapping the remaining instructions
    struct Early {
        bool enabled = false;
        std::vector<int> remaining;          // per buffer: accumulations still to come in this sweep
        std::vector<void *> dst;             // per buffer: host destination (nullptr = none)
        std::vector<char> issued;
        cudaEvent_t ev[kMaxArgs * 2] = {nullptr};
        int n_ev = 0;
    } early;
};

static cudaStream_t S(rdb_tape *t) { return t->ctx->stream; }

static int dev_alloc(rdb_tape *t, void **p, int64_t bytes) {
    if (bytes <= 0) bytes = 8;
    cudaError_t e = cudaMalloc(p, (size_t)bytes);
    if (e != cudaSuccess) return fail(RDB_ENOMEM, "cudaMalloc(%lld bytes) failed: %s", (long long)bytes, cudaGetErrorString(e));
    t->allocs.push_back(*p);
    return RDB_OK;
}
static void dev_free(rdb_tape *t, void *p) {
    if (!p) return;
    auto it = std::find(t->allocs.begin(), t->allocs.end(), p);
    if (it != t->allocs.end()) t->allocs.erase(it);
    cudaFree(p);
}

// per-step profiling ----------------------------------------------------------------------------------
struct StepTimer {
    rdb_tape *t;
    const char *name;
    double bytes, flops;
    bool on;
    StepTimer(rdb_tape *t_, const char *n, double b, double f) : t(t_), name(n), bytes(b), flops(f), on(t_->opts.profile != 0) {
        if (on) cudaEventRecord(t->ev0, S(t));
    }
    ~StepTimer() {
        if (!on) return;
        cudaEventRecord(t->ev1, S(t));
        cudaEventSynchronize(t->ev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, t->ev0, t->ev1);
        for (auto &s : t->stats)
            if (s.name == name) { s.ms += ms; s.bytes += bytes; s.flops += flops; s.count++; return; }
        StepStat s; s.name = name; s.ms = ms; s.bytes = bytes; s.flops = flops; s.count = 1;
        t->stats.push_back(s);
    }
};

// ------------------------------------------------------------------------------------------ typed ops
static thread_local int g_f64_mode = 0, g_f64_slices = 0;   // set from the tape's options around each API call
template <typename T>
static cudaError_t gemm_t(cudaStream_t s, bool ta, bool tb, int64_t M, int64_t N, int64_t K, T alpha, const T *A, int64_t lda,
                          const T *B, int64_t ldb, T beta, T *C, int64_t ldc);
static thread_local int64_t g_extra_launches = 0;           // kernels behind gemm calls beyond the one KL() counts
template <>
cudaError_t gemm_t<double>(cudaStream_t s, bool ta, bool tb, int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                           int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc) {
    cudaError_t e = gemm_f64(s, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, g_f64_mode, g_f64_slices);
    g_extra_launches += g_gemm_launches - 1;
    return e;
}
template <>
cudaError_t gemm_t<float>(cudaStream_t s, bool ta, bool tb, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                          int64_t lda, const float *B, int64_t ldb, float beta, float *C, int64_t ldc) {
    cudaError_t e = gemm_f32(s, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    g_extra_launches += g_gemm_launches - 1;
    return e;
}

#define KL(t, call)                                                                                        \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        (t)->launches++;                                                                                   \
        if (e__ != cudaSuccess)                                                                            \
            return fail(RDB_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// identity of value(buffer) for the sliced-operand cache (kernels.h: g_tag_a/g_tag_b); never 0
static uint64_t value_tag(rdb_tape *t, int b) {
    while (t->bufs[b].alias_of >= 0) b = t->bufs[b].alias_of;
    return ((uint64_t)(t->uid & 0xFFFF) << 48) | ((uint64_t)((b + 1) & 0xFFFF) << 32) | (uint64_t)t->bufs[b].version;
}
static void bump_version(rdb_tape *t, int b) {
    while (t->bufs[b].alias_of >= 0) b = t->bufs[b].alias_of;
    t->bufs[b].version++;
}
static uint64_t operand_tag(rdb_tape *t, const OperandPlan &o) { return o.mode == OPM_GATHER ? 0 : value_tag(t, o.rec.buffer); }

static bool is_elementwise(int op) {
    return op == OP_NBCAST || op == OP_MAP || op == OP_BCAST || op == OP_BCAST_ADD || op == OP_BCAST_SUB || op == OP_BCAST_MUL ||
           op == OP_BCAST_DIV || op == OP_BCAST_LDIV || op == OP_BCAST_POW || op == OP_TRACKER_BCAST;
}
static bool is_getindex(int op) { return op == OP_GETINDEX || op == OP_GETINDEX_GENERIC; }

// A matrix operand as the GEMM sees it: stored pointer, leading dimension and whether the stored matrix is the
// transpose of the operand (folded IDX transposes cost nothing: SURVEY.md Appendix A.1 "engine-minimal").
template <typename T>
struct MatRef {
    const T *ptr;
    int64_t ld;
    bool stored_t;
};

template <typename T>
static int operand_value(rdb_tape *t, OperandPlan &o, MatRef<T> &m) {
    Buf &b = t->bufs[o.rec.buffer];
    if (o.mode == OPM_PLAIN) { m.ptr = (const T *)b.val; m.ld = o.rows; m.stored_t = false; return RDB_OK; }
    if (o.mode == OPM_FOLDED_T) { m.ptr = (const T *)b.val; m.ld = o.cols; m.stored_t = true; return RDB_OK; }
    // pull_value! element by element + dense copy (tracked.jl:178-181,103) as ONE gather kernel
    KL(t, launch_gather<T>(S(t), (T *)o.dense_val, (const T *)b.val, o.d_map, o.size));
    m.ptr = (const T *)o.dense_val; m.ld = o.rows; m.stored_t = false;
    return RDB_OK;
}

// ------------------------------------------------------------------------------------------ forward
static void plan_col_panels(int64_t rows, int64_t cols, int64_t width[4]);
template <typename T> struct MatRef;
template <typename T> static void mul_operand_ref(rdb_tape *t, OperandPlan &o, MatRef<T> &m);
static bool preslice_on(rdb_tape *t);
static int preslice_begin(rdb_tape *t);
static int settle_uploads(rdb_tape *t, int keep);
static int droot(rdb_tape *t, int b);
template <typename T>
static int expr_params(rdb_tape *t, InstrPlan &I, bool reduce, int order, ExprParams &p) {
    Buf &out = t->bufs[I.rec.out];
    p = ExprParams{};
    p.n_args = I.rec.n_in; p.n_ops = (int)I.rec.expr_len; p.order = order; p.reduce = reduce ? 1 : 0;
    p.n = out.size;
    for (int d = 0; d < kMaxDims; ++d) p.dims[d] = out.dims[d];
    for (int a = 0; a < I.rec.n_in; ++a) {
        OperandPlan &o = I.in[a];
        MatRef<T> m;
        OK(operand_value<T>(t, o, m));
        if (m.stored_t) return fail(RDB_EUNSUPPORTED, "elementwise instruction on a transposed operand");
        p.arg[a].ptr = m.ptr;
        int64_t st = 1;
        bool same = true;
        for (int d = 0; d < kMaxDims; ++d) {
            int64_t od = d < o.rec.ndims ? o.rec.dims[d] : 1;
            p.arg[a].stride[d] = (od == 1 && out.dims[d] != 1) ? 0 : st;
            if (od != out.dims[d]) same = false;
            st *= od;
        }
        p.arg[a].same = same ? 1 : 0;
        p.partial[a] = I.partial_alias[a] >= 0 ? nullptr : I.partial[a];     // an aliased partial is written once, by the argument it aliases
    }
    for (int h = 0; h < kMaxArgs * (kMaxArgs + 1) / 2; ++h) p.second[h] = I.second[h];
    p.out = out.val; p.ops = I.d_ops; p.fconst = t->d_fconst; p.block_sums = t->scratch;
    return RDB_OK;
}

// ---- north star (a): a run of adjacent elementwise instructions of one shape replays as ONE forward kernel that still stores every
// instruction's value and partials.  An instruction joins the run that starts at `head` when its output has the head's shape and
// every argument produced inside the run is a plain same-shape TrackedArray operand (it then travels in a register); arguments
// from outside the run are loaded as usual (broadcast clamps included).  Needs the NVRTC path; without it the instructions launch
// one by one.  Results are bit-identical either way (same arithmetic, same order).
// Which arguments of a scalar program have IDENTICAL partials by construction?  The forward kernels evaluate one Dual{N}: register r
// carries partials D[r][p].  This walks the same recursion symbolically, using only identities that are exact in floating point for
// every input (x + 0 = x, 0 + x = x, f1 * 1 = f1); everything else stays an opaque term that names its register and its argument.
// `tanh(x + y)` gives D[.][x] = D[.][y] = "f1_3": the reference stores partial_x and partial_y separately (broadcast.jl:155,182-191),
// but they hold the same numbers, so one array serves both (25 % of the forward sweep's bytes at the MLP config).
static void partial_signatures(const int32_t *ops, int64_t n_ops, int n_args, std::vector<std::string> &sig) {
    std::vector<std::vector<std::string>> D((size_t)n_args + n_ops, std::vector<std::string>(n_args));
    for (int a = 0; a < n_args; ++a) for (int p = 0; p < n_args; ++p) D[a][p] = a == p ? "1" : "0";
    for (int64_t k = 0; k < n_ops; ++k) {
        const int code = ops[4 * k], a = ops[4 * k + 1], b = ops[4 * k + 2], imm = ops[4 * k + 3];
        const size_t r = (size_t)n_args + k;
        for (int p = 0; p < n_args; ++p) {
            std::string &d = D[r][p];
            const std::string tag = std::to_string(r) + "_" + std::to_string(p);
            switch (code) {
                case E_CONST: d = "0"; break;
                case E_ARG: d = D[imm][p]; break;
                case E_ADD: d = D[a][p] == "0" ? D[b][p] : (D[b][p] == "0" ? D[a][p] : "(" + D[a][p] + "+" + D[b][p] + ")"); break;
                case E_SUB: d = D[b][p] == "0" ? D[a][p] : "(" + D[a][p] + "-" + D[b][p] + ")"; break;
                case E_NEG: case E_POWI: case E_TANH: case E_EXP: case E_LOG: case E_SQRT: case E_SIN: case E_COS: case E_ABS2: case E_ABS:
                case E_INV: case E_SIGMOID: case E_LOG1P: case E_EXPM1: case E_ASIN: case E_ACOS: case E_ATAN: case E_SINH: case E_COSH: case E_TAN:
                    // f1(v_a) * D[a][p]: the factor depends on the register only
                    d = D[a][p] == "1" ? "f1_" + std::to_string(r) : "(f1_" + std::to_string(r) + "*" + D[a][p] + ")";
                    break;
                default: d = "op" + tag; break;               // products, quotients, powers, max/min, two-argument functions: opaque
            }
        }
    }
    sig = D.back();
}

static int plan_run(rdb_tape *t, size_t head) {
    InstrPlan &H = t->instrs[head];
    if (H.run_end != -2) return H.run_end;
    H.run_end = (int)head;
    static const bool on = [] { const char *e = getenv("RDB_FUSE_RUNS"); return !(e && e[0] == '0'); }();
    if (!on || !t->opts.fuse_elementwise) return H.run_end;
    Buf &ho = t->bufs[H.rec.out];
    auto root = [&](int b) { while (t->bufs[b].alias_of >= 0) b = t->bufs[b].alias_of; return b; };
    auto joinable = [&](size_t k) {
        InstrPlan &I = t->instrs[k];
        if (I.blk || !is_elementwise(I.rec.opcode) || I.rec.expr_len <= 0) return false;
        Buf &o = t->bufs[I.rec.out];
        if (o.size != ho.size) return false;
        for (int d = 0; d < kMaxDims; ++d) if (o.dims[d] != ho.dims[d]) return false;
        for (int a = 0; a < I.rec.n_in; ++a) {
            OperandPlan &op = I.in[a];
            bool inside = false;
            for (size_t j = head; j < k; ++j) inside |= root(t->instrs[j].rec.out) == root(op.rec.buffer);
            if (!inside) { if (op.rec.cls == CLS_TR) return false; continue; }       // (scalar operands keep the one-by-one path)
            if (op.rec.cls != CLS_TA || op.mode != OPM_PLAIN || t->bufs[op.rec.buffer].size != ho.size) return false;
            for (int d = 0; d < kMaxDims; ++d) if ((d < op.rec.ndims ? op.rec.dims[d] : 1) != ho.dims[d]) return false;
        }
        return true;
    };
    if (!joinable(head)) return H.run_end;
    size_t k = head + 1;
    while (k < t->instrs.size() && k - head < 8 && joinable(k)) ++k;
    H.run_end = (int)k - 1;
    return H.run_end;
}
// 1 = launched as one kernel, 0 = not available (caller launches the instructions one by one)
template <typename T>
static int expr_forward_run(rdb_tape *t, size_t head, size_t last, bool reduce) {
    InstrPlan &H = t->instrs[head];
    const int ri = reduce ? 1 : 0;
    if (H.run_tried[ri] && !H.run_fn[ri]) return 0;
    const int n_items = (int)(last - head + 1);
    std::vector<JitRunItem> items(n_items);
    auto root = [&](int b) { while (t->bufs[b].alias_of >= 0) b = t->bufs[b].alias_of; return b; };
    double bytes = 0;
    for (int k = 0; k < n_items; ++k) {
        InstrPlan &I = t->instrs[head + k];
        OK(expr_params<T>(t, I, false, 1, items[k].p));
        items[k].ops = I.h_ops.data();
        bytes += (double)t->bufs[I.rec.out].size * t->es;
        for (int a = 0; a < I.rec.n_in; ++a) {
            items[k].from_item[a] = -1;
            for (int j = 0; j < k; ++j) if (root(t->instrs[head + j].rec.out) == root(I.in[a].rec.buffer)) items[k].from_item[a] = j;
            if (items[k].from_item[a] < 0) bytes += (double)t->bufs[I.in[a].rec.buffer].size * t->es;
            if (I.partial[a] && I.partial_alias[a] < 0) bytes += (double)t->bufs[I.rec.out].size * t->es;
        }
    }
    if (!H.run_tried[ri]) {
        H.run_tried[ri] = true;
        H.run_fn[ri] = jit_run_kernel(items.data(), n_items, t->h_fconst.data(), sizeof(T) == 4, reduce);
        if (!H.run_fn[ri]) return 0;
    }
    StepTimer tm(t, reduce ? "expr_forward_run+sum" : "expr_forward_run", bytes, 0);
    KL(t, jit_run_launch(S(t), H.run_fn[ri], items.data(), n_items, t->scratch, reduce));
    return 1;
}

template <typename T>
static int expr_forward(rdb_tape *t, InstrPlan &I, bool reduce, int order) {
    Buf &out = t->bufs[I.rec.out];
    ExprParams p{};
    OK(expr_params<T>(t, I, reduce, order, p));
    int np = 0;
    for (int a = 0; a < I.rec.n_in; ++a) np += (I.partial[a] && I.partial_alias[a] < 0) ? 1 : 0;
    StepTimer tm(t, order == 2 ? "expr_forward2" : (reduce ? "expr_forward+sum" : "expr_forward"),
                 (double)out.size * t->es * (1 + np + 1), 0);
    if (order == 1 && out.size > 0) {
        const int ri = reduce ? 1 : 0;
        if (!I.jit_tried[ri]) {
            I.jit_tried[ri] = true;
            I.jit_fn[ri] = jit_expr_kernel(p, I.h_ops.data(), t->h_fconst.data(), sizeof(T) == 4, reduce);
        }
        if (I.jit_fn[ri]) {
            KL(t, jit_expr_launch(S(t), I.jit_fn[ri], p, reduce));
            return RDB_OK;
        }
    }
    KL(t, launch_expr_forward<T>(S(t), p));
    return RDB_OK;
}

template <typename T>
static int forward_pass(rdb_tape *t, int order = 1) {
    const size_t n = t->instrs.size();
    // An input that is still arriving in column panels (rdb_tape_set_input) is consumed panel by panel when its first reader
    // is a `*` that has it as the plain right operand: out[:, panel] = op(A) * b[:, panel] starts as soon as the panel has
    // landed, so the GEMM hides behind the upload.  Every other pending input is waited for here.
    int streamed = -1, streamed_at = -1;
    if (order == 1 && t->last_upload >= 0 && t->bufs[t->last_upload].up_n > 0) {
        const int cand = t->last_upload;
        for (size_t i = 0; i < n && streamed_at < 0; ++i) {
            InstrPlan &I = t->instrs[i];
            bool reads = false;
            if (I.blk) { for (int r = 0; r < I.blk->n_ref; ++r) reads |= droot(t, I.blk->refbufs[r]) == cand; }
            else for (int a = 0; a < I.rec.n_in; ++a) reads |= droot(t, I.in[a].rec.buffer) == cand;
            if (!reads) continue;
            streamed_at = (int)i;
            if (!I.blk && I.rec.opcode == OP_MUL && I.in[1].rec.buffer == cand && I.in[1].mode == OPM_PLAIN &&
                droot(t, I.in[0].rec.buffer) != cand && I.in[1].cols == t->bufs[cand].up_c0[t->bufs[cand].up_n])
                streamed = cand;
        }
    }
    OK(settle_uploads(t, streamed));
    for (size_t i = 0; i < n; ++i) {
        InstrPlan &I = t->instrs[i];
        Buf &out = t->bufs[I.rec.out];
        bump_version(t, I.rec.out);
        // fusion of an elementwise producer with the `sum`/`mean` that consumes it (adjacent instructions only)
        bool fuse_next = false;
        if (t->opts.fuse_elementwise && i + 1 < n) {
            InstrPlan &Nx = t->instrs[i + 1];
            if ((Nx.rec.opcode == OP_SUM || Nx.rec.opcode == OP_MEAN) && Nx.in[0].rec.buffer == I.rec.out &&
                Nx.in[0].mode == OPM_PLAIN &&
                (I.rec.opcode == OP_ADD || I.rec.opcode == OP_SUB || is_elementwise(I.rec.opcode)))
                fuse_next = true;
        }
        if (order == 1 && !I.blk && is_elementwise(I.rec.opcode)) {
            const size_t last = (size_t)plan_run(t, i);
            if (last > i) {
                InstrPlan &L = t->instrs[last];
                bool red = false;                                    // the `sum`/`mean` right behind the run rides along too
                if (t->opts.fuse_elementwise && last + 1 < n) {
                    InstrPlan &Nx = t->instrs[last + 1];
                    red = (Nx.rec.opcode == OP_SUM || Nx.rec.opcode == OP_MEAN) && Nx.in[0].rec.buffer == L.rec.out && Nx.in[0].mode == OPM_PLAIN;
                }
                const int rc = expr_forward_run<T>(t, i, last, red);
                if (rc < 0) return rc;
                if (rc == 1) {
                    for (size_t k = i + 1; k <= last; ++k) { bump_version(t, t->instrs[k].rec.out); t->instrs[k].fused_into_prev = false; }
                    if (last + 1 < n) t->instrs[last + 1].fused_into_prev = red;
                    i = last;
                    continue;
                }
            }
        }
        switch (I.rec.opcode) {
            case OP_MUL: {
                MatRef<T> A, B;
                OK(operand_value<T>(t, I.in[0], A));
                OK(operand_value<T>(t, I.in[1], B));
                int64_t M = I.in[0].rows, K = I.in[0].cols, N = I.in[1].cols;
                StepTimer tm(t, "gemm_forward", (double)(M * K + K * N + M * N) * t->es, 2.0 * M * N * K);
                if (streamed >= 0 && (int)i == streamed_at) {
                    Buf &bi = t->bufs[streamed];
                    if (preslice_on(t) && M >= 256 && N >= 256 && K >= 256) {
                        // while the panels arrive: slice what the next `*` reads of the inputs that are already complete
                        for (size_t j = i + 1; j < n; ++j) {
                            InstrPlan &J = t->instrs[j];
                            if (J.blk || J.rec.opcode != OP_MUL) continue;
                            if (J.in[0].rows >= 256 && J.in[1].cols >= 256 && J.in[0].cols >= 256) {
                                OK(preslice_begin(t));
                                for (int b = 0; b < 2; ++b) {
                                    OperandPlan &o = J.in[b];
                                    bool ready = o.mode != OPM_GATHER && droot(t, o.rec.buffer) != droot(t, streamed);
                                    for (size_t x = i; x < j && ready; ++x) if (droot(t, t->instrs[x].rec.out) == droot(t, o.rec.buffer)) ready = false;
                                    if (!ready) continue;
                                    MatRef<T> R2;
                                    mul_operand_ref<T>(t, o, R2);
                                    int nl2 = 0;
                                    if (b == 0) ozaki_preslice(t->ctx->side_stream, (const double *)R2.ptr, R2.ld, o.rows, o.cols, R2.stored_t, t->opts.ozaki_slices, operand_tag(t, o), &nl2);
                                    else ozaki_preslice(t->ctx->side_stream, (const double *)R2.ptr, R2.ld, o.cols, J.in[0].cols, !R2.stored_t, t->opts.ozaki_slices, operand_tag(t, o), &nl2);
                                    t->launches += nl2;
                                }
                            }
                            break;
                        }
                    }
                    for (int p = 0; p < bi.up_n; ++p) {
                        const int64_t c0 = bi.up_c0[p], nc = bi.up_c0[p + 1] - c0;
                        CU(cudaStreamWaitEvent(S(t), bi.up_ev[p], 0));
                        g_tag_a = operand_tag(t, I.in[0]);
                        KL(t, gemm_t<T>(S(t), A.stored_t, false, M, nc, K, (T)1, A.ptr, A.ld, B.ptr + c0 * B.ld, B.ld, (T)0, (T *)out.val + c0 * M, M));
                    }
                    bi.up_n = 0;
                    if (preslice_on(t) && M >= 256 && N >= 256 && K >= 256) {
                        // the panels were sliced one by one: slice the whole operand once more, off the critical path, so the
                        // reverse sweep finds it in the cache like after an unstreamed forward GEMM
                        OK(preslice_begin(t));
                        int nl2 = 0;
                        ozaki_preslice(t->ctx->side_stream, (const double *)B.ptr, B.ld, N, K, true, t->opts.ozaki_slices, operand_tag(t, I.in[1]), &nl2);
                        t->launches += nl2;
                    }
                    break;
                }
                if (order == 1 && preslice_on(t) && M >= 256 && N >= 256 && K >= 256) {
                    // while this GEMM runs: slice the operands of the next `*` whose operands exist already (tape inputs,
                    // constants, outputs of earlier instructions)
                    for (size_t j = i + 1; j < n; ++j) {
                        InstrPlan &J = t->instrs[j];
                        if (J.blk || J.rec.opcode != OP_MUL) continue;
                        bool ready = J.in[0].mode != OPM_GATHER && J.in[1].mode != OPM_GATHER && J.in[0].rows >= 256 && J.in[1].cols >= 256 &&
                                     J.in[0].cols >= 256;
                        for (size_t x = i; x < j && ready; ++x)
                            for (int b = 0; b < 2; ++b)
                                if (droot(t, t->instrs[x].rec.out) == droot(t, J.in[b].rec.buffer)) ready = false;
                        if (ready) {
                            OK(preslice_begin(t));
                            MatRef<T> A2, B2;
                            mul_operand_ref<T>(t, J.in[0], A2);
                            mul_operand_ref<T>(t, J.in[1], B2);
                            int nla = 0, nlb = 0;
                            ozaki_preslice(t->ctx->side_stream, (const double *)A2.ptr, A2.ld, J.in[0].rows, J.in[0].cols, A2.stored_t,
                                           t->opts.ozaki_slices, operand_tag(t, J.in[0]), &nla);
                            ozaki_preslice(t->ctx->side_stream, (const double *)B2.ptr, B2.ld, J.in[1].cols, J.in[0].cols, !B2.stored_t,
                                           t->opts.ozaki_slices, operand_tag(t, J.in[1]), &nlb);
                            t->launches += nla + nlb;
                        }
                        break;
                    }
                }
                g_tag_a = operand_tag(t, I.in[0]); g_tag_b = operand_tag(t, I.in[1]);
                KL(t, gemm_t<T>(S(t), A.stored_t, B.stored_t, M, N, K, (T)1, A.ptr, A.ld, B.ptr, B.ld, (T)0, (T *)out.val, M));
            } break;
            case OP_ADD:
            case OP_SUB: {
                MatRef<T> A, B;
                OK(operand_value<T>(t, I.in[0], A));
                OK(operand_value<T>(t, I.in[1], B));
                StepTimer tm(t, fuse_next ? "add+sum_forward" : "add_forward", 3.0 * out.size * t->es, 0);
                KL(t, launch_add<T>(S(t), (T *)out.val, A.ptr, B.ptr, I.rec.opcode == OP_ADD ? (T)1 : (T)-1, out.size,
                                    fuse_next ? (T *)t->scratch : nullptr));
            } break;
            case OP_NEG: {
                MatRef<T> A;
                OK(operand_value<T>(t, I.in[0], A));
                KL(t, launch_axpy<T>(S(t), (T *)out.val, A.ptr, (T)-1, out.size, true));
            } break;
            case OP_SUM:
            case OP_MEAN: {
                T scale = I.rec.opcode == OP_MEAN ? (T)1 / (T)I.in[0].size : (T)1;
                if (!I.fused_into_prev) {
                    MatRef<T> A;
                    OK(operand_value<T>(t, I.in[0], A));
                    StepTimer tm(t, "sum_forward", (double)I.in[0].size * t->es, 0);
                    KL(t, launch_block_sums<T>(S(t), A.ptr, I.in[0].size, (T *)t->scratch));
                }
                KL(t, launch_finish_sum<T>(S(t), (const T *)t->scratch, (T *)out.val, scale));
            } break;
            case OP_NBCAST: case OP_MAP: case OP_BCAST: case OP_BCAST_ADD: case OP_BCAST_SUB: case OP_BCAST_MUL:
            case OP_BCAST_DIV: case OP_BCAST_LDIV: case OP_BCAST_POW: case OP_TRACKER_BCAST:
                OK(expr_forward<T>(t, I, fuse_next, order));
                break;
            case OP_DOT: {
                MatRef<T> A, B;
                OK(operand_value<T>(t, I.in[0], A));
                OK(operand_value<T>(t, I.in[1], B));
                StepTimer tm(t, "dot_forward", 2.0 * I.in[0].size * t->es, 0);
                KL(t, launch_dot_block_sums<T>(S(t), A.ptr, B.ptr, I.in[0].size, (T *)t->scratch));
                KL(t, launch_finish_sum<T>(S(t), (const T *)t->scratch, (T *)out.val, (T)1));
            } break;
            case OP_SUMDIMS: {
                Buf &src = t->bufs[I.in[0].rec.buffer];
                const int64_t d0 = src.dims[0], d1 = src.size / src.dims[0];
                StepTimer tm(t, "sumdims_forward", (double)src.size * t->es, 0);
                if (out.size == 1) {
                    KL(t, launch_block_sums<T>(S(t), (const T *)src.val, src.size, (T *)t->scratch));
                    KL(t, launch_finish_sum<T>(S(t), (const T *)t->scratch, (T *)out.val, (T)1));
                } else if (out.dims[0] == 1 && out.size == d1) {            // sum(x, dims=1): column sums
                    KL(t, launch_colsum<T>(S(t), (T *)out.val, (const T *)src.val, d0, d1, (T)1));
                } else if (out.dims[0] == d0 && out.size == d0) {           // sum over every trailing dimension: row sums
                    CU(cudaMemsetAsync(out.val, 0, (size_t)out.size * t->es, S(t)));
                    KL(t, launch_mul_acc_colvec<T>(S(t), (T *)out.val, (const T *)src.val, (const T *)nullptr, d0, d1, 1, (T *)t->scratch, t->scratch_bytes / 8));
                } else {
                    return fail(RDB_EUNSUPPORTED, "sum over this combination of dimensions is not lowered");
                }
            } break;
            case OP_GETINDEX: case OP_GETINDEX_GENERIC: {
                Buf &src = t->bufs[I.in[0].rec.buffer];
                StepTimer tm(t, "getindex_forward", 2.0 * out.size * t->es + 8.0 * out.size, 0);
                KL(t, launch_gather<T>(S(t), (T *)out.val, (const T *)src.val, I.in[0].d_map, out.size));
            } break;
            case OP_SCALAR_BLOCK: {
                // scalar_forward_exec! for the whole run (scalars.jl:80-123), level-parallel
                ScalarBlock &B = *I.blk;
                if (order != 1) return fail(RDB_EUNSUPPORTED, "second-order sweep over scalar instructions");
                for (int k = 0; k < B.n_ref; ++k)
                    if (B.ref_has_export[k]) bump_version(t, B.refbufs[k]);
                StepTimer tm(t, "scalar_forward", (double)(B.n_leaf + 3 * B.n + B.n_edges) * t->es + 16.0 * B.n, 0);
                int nl = 0;
                KL(t, scalar_block_forward<T>(S(t), B, t->d_fconst, &nl));
                t->launches += nl - 1;
            } break;
            default:
                return fail(RDB_EUNSUPPORTED, "opcode %d has no forward kernel", I.rec.opcode);
        }
        if (i + 1 < n) t->instrs[i + 1].fused_into_prev = fuse_next;
    }
    t->forward_valid = true;
    return RDB_OK;
}

// ------------------------------------------------------------------------------------------ reverse
// called after every accumulation into deriv(buffer): when it was the last one of this sweep for a tape input whose result is
// wanted on the host, start the D2H copy on the copy stream right away
static int acc_done(rdb_tape *t, int b) {
    if (!t->early.enabled) return RDB_OK;
    while (t->bufs[b].alias_of >= 0) b = t->bufs[b].alias_of;
    if (!t->early.dst[b] || t->early.issued[b]) return RDB_OK;
    if (--t->early.remaining[b] > 0) return RDB_OK;
    if (t->early.n_ev >= kMaxArgs * 2) return RDB_OK;
    cudaEvent_t &ev = t->early.ev[t->early.n_ev];
    if (!ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    t->early.n_ev++;
    CU(cudaEventRecord(ev, S(t)));
    CU(cudaStreamWaitEvent(t->ctx->copy_stream, ev, 0));
    CU(cudaMemcpyAsync(t->early.dst[b], t->bufs[b].der, (size_t)t->bufs[b].size * t->es, cudaMemcpyDeviceToHost, t->ctx->copy_stream));
    t->early.issued[b] = 1;
    return RDB_OK;
}

template <typename T>
static int unseed(rdb_tape *t, Buf &b, int64_t R) {
    CU(cudaMemsetAsync(b.der, 0, (size_t)(b.size * R) * t->es, S(t)));
    return RDB_OK;
}
static int droot(rdb_tape *t, int b) {
    while (t->bufs[b].alias_of >= 0) b = t->bufs[b].alias_of;
    return b;
}
// logical unseed!: no memory traffic
static void dclear(rdb_tape *t, int b) { t->dz[droot(t, b)] = 1; }
// before a full-coverage accumulation: returns true if it must OVERWRITE (buffer was logically zero); marks the buffer live
static bool dtake(rdb_tape *t, int b) {
    const int r = droot(t, b);
    const bool ow = t->dz[r] != 0;
    t->dz[r] = 0;
    if (t->ctx && t->bufs[r].der) t->ctx->gemm.uniform_clear(t->bufs[r].der);   // about to be written: no longer known to be a fill
    return ow;
}
// deriv(b) was just OVERWRITTEN with one value in every element (adjoint of sum / mean, one seed): the int8 slicer may write its
// digit planes from element 0 (GemmState::uniform); any later accumulation goes through dtake / dmaterialize and withdraws this
static void dmark_uniform(rdb_tape *t, int b, int64_t R) {
    const int r = droot(t, b);
    if (R == 1 && t->dtype == RDB_F64 && t->ctx && t->bufs[r].der) t->ctx->gemm.uniform_set(t->bufs[r].der);
}
// before a partial / atomic accumulation, or any read from outside the sweep: make the storage really zero
static int dmaterialize(rdb_tape *t, int b, int64_t R, bool mark_live) {
    const int r = droot(t, b);
    if (t->dz[r] && t->bufs[r].der) {
        CU(cudaMemsetAsync(t->bufs[r].der, 0, (size_t)(t->bufs[r].size * R) * t->es, S(t)));
        if (mark_live) t->dz[r] = 0;
    }
    if (mark_live && t->ctx && t->bufs[r].der) t->ctx->gemm.uniform_clear(t->bufs[r].der);
    return RDB_OK;
}

// Column panels, in 64-column tiles, chosen so that every panel's 128x64 GEMM tile count fills whole waves of the 148 SMs
// (a 4096^2 GEMM is 13.8 waves; four equal panels would be 4 x 3.46 -> 16).  width[i] = columns of panel i (0 = unused).
static void plan_col_panels(int64_t rows, int64_t cols, int64_t width[4]) {
    const int64_t rt = (rows + 127) / 128, C = (cols + 63) / 64;
    auto waves = [&](int64_t x) { return (rt * x + 147) / 148; };
    int64_t best = -1;
    width[0] = width[1] = width[2] = width[3] = 0;
    for (int P = 3; P <= 4; ++P) {
        const int64_t mid = C / P;
        for (int64_t x1 = std::max<int64_t>(1, mid - 6); x1 <= mid + 6; ++x1)
            for (int64_t x2 = std::max<int64_t>(1, mid - 6); x2 <= mid + 6; ++x2)
                for (int64_t x3 = (P == 4 ? std::max<int64_t>(1, mid - 6) : 0); x3 <= (P == 4 ? mid + 6 : 0); ++x3) {
                    const int64_t last = C - x1 - x2 - x3;
                    if (last < 1 || last > mid + 6) continue;
                    // total waves, then the smallest last panel (its copy is the tail of the call)
                    const int64_t cost = (waves(x1) + waves(x2) + (x3 ? waves(x3) : 0) + waves(last)) * 4096 + last;
                    if (best < 0 || cost < best) { best = cost; width[0] = x1; width[1] = x2; width[2] = x3 ? x3 : last; width[3] = x3 ? last : 0; }
                }
    }
    if (best < 0) { const int64_t q = (C + 3) / 4; width[0] = width[1] = width[2] = q; width[3] = C > 3 * q ? C - 3 * q : 0; }
    int64_t c0 = 0;
    for (int i = 0; i < 4; ++i) { width[i] = std::max<int64_t>(0, std::min<int64_t>(width[i] * 64, cols - c0)); c0 += width[i]; }
    if (c0 < cols) width[3] += cols - c0;
}

// Inputs uploaded in panels (rdb_tape_set_input): make the main stream wait for them.  `keep` is left pending (its consumer
// waits panel by panel).
static int settle_uploads(rdb_tape *t, int keep) {
    if (t->sep.lazy_x) {                        // a device input that the one-kernel hessian! would have read in place: copy it now
        Buf &x = t->bufs[t->inputs[0]];
        CU(cudaMemcpyAsync(x.val, t->sep.lazy_x, (size_t)x.size * t->es, cudaMemcpyDeviceToDevice, S(t)));
        t->sep.lazy_x = nullptr;
    } else if (t->sep.x_consumed) {
        return fail(RDB_EINVAL, "the device input was read in place by an asynchronous hessian! call and never copied into the tape: "
                                "call rdb_tape_set_input again before this entry point");
    }
    for (size_t b = 0; b < t->bufs.size(); ++b) {
        Buf &B = t->bufs[b];
        if (B.up_n == 0 || (int)b == keep) continue;
        CU(cudaStreamWaitEvent(S(t), B.up_ev[B.up_n - 1], 0));     // panels arrive in order on one stream
        B.up_n = 0;
    }
    return RDB_OK;
}

// ---- `*` adjoints by column panels of the accumulated buffer (single seed) -------------------------------------------------
// mul_operand_ref: the GEMM view of a `*` operand's value (values are unchanged since the forward sweep: the densified copies of
// OPM_GATHER operands are reused)
template <typename T>
static void mul_operand_ref(rdb_tape *t, OperandPlan &o, MatRef<T> &m) {
    Buf &b = t->bufs[o.rec.buffer];
    if (o.mode == OPM_PLAIN) { m.ptr = (const T *)b.val; m.ld = o.rows; m.stored_t = false; }
    else if (o.mode == OPM_FOLDED_T) { m.ptr = (const T *)b.val; m.ld = o.cols; m.stored_t = true; }
    else { m.ptr = (const T *)o.dense_val; m.ld = o.rows; m.stored_t = false; }
}
static uint64_t g_delta_serial = 0;
static uint64_t new_delta_tag() { return (1ull << 63) | (++g_delta_serial & 0x7FFFFFFFFFFFFFFFull); }

// shape of the buffer an operand's adjoint lands in (the operand itself, or the origin of a folded transpose)
static void mul_root_shape(const OperandPlan &o, int64_t &rows, int64_t &cols) {
    if (o.mode == OPM_FOLDED_T) { rows = o.cols; cols = o.rows; } else { rows = o.rows; cols = o.cols; }
}

// columns [c0, c0 + nc) of deriv(root of operand `side`) (+)= this instruction's contribution; dtag identifies the content of
// deriv(output) for the sliced-operand cache (the same delta is an operand of every panel)
template <typename T>
static int mul_acc_cols(rdb_tape *t, InstrPlan &I, int side, int64_t c0, int64_t nc, T beta, uint64_t dtag) {
    OperandPlan &oa = I.in[0], &ob = I.in[1];
    const int64_t M = oa.rows, K = oa.cols, N = ob.cols;
    MatRef<T> A, B;
    mul_operand_ref<T>(t, oa, A);
    mul_operand_ref<T>(t, ob, B);
    const T *delta = (const T *)t->bufs[I.rec.out].der;
    // A panel reads row windows of operands the full GEMM would read whole: name the whole operand (tag + window) so the int8
    // path slices it once for all panels - and finds it in the cache when the forward sweep or the slicing ahead left it there.
    auto window = [](OperandWindow &w, const void *full, int64_t rows_full, int64_t row0, int64_t rows) {
        w = OperandWindow();
        if (rows != rows_full) { w.full = full; w.rows_full = rows_full; w.row0 = row0; }
    };
    if (side == 0) {
        T *dst = (T *)t->bufs[oa.rec.buffer].der;
        if (oa.mode == OPM_PLAIN) {
            // a.deriv(MxK)[:, c0:] = delta(MxN) * Bmat'(N x k-panel)
            const T *bp = B.stored_t ? B.ptr + c0 * B.ld : B.ptr + c0;
            g_tag_a = dtag;
            g_tag_b = operand_tag(t, ob);
            window(g_win_b, B.ptr, K, c0, nc);
            KL(t, gemm_t<T>(S(t), false, !B.stored_t, M, nc, N, (T)1, delta, M, bp, B.ld, beta, dst + c0 * M, M));
        } else {
            // origin.deriv(KxM)[:, c0:] = Bmat(KxN) * delta'(N x m-panel)
            g_tag_a = operand_tag(t, ob);
            g_tag_b = dtag;
            window(g_win_b, delta, M, c0, nc);
            KL(t, gemm_t<T>(S(t), B.stored_t, true, K, nc, N, (T)1, B.ptr, B.ld, delta + c0, M, beta, dst + c0 * K, K));
        }
    } else {
        T *dst = (T *)t->bufs[ob.rec.buffer].der;
        if (ob.mode == OPM_PLAIN) {
            // b.deriv(KxN)[:, c0:] = Amat'(KxM) * delta(M x n-panel)
            g_tag_a = operand_tag(t, oa);
            g_tag_b = dtag;
            window(g_win_b, delta, N, c0, nc);
            KL(t, gemm_t<T>(S(t), !A.stored_t, false, K, nc, M, (T)1, A.ptr, A.ld, delta + c0 * M, M, beta, dst + c0 * K, K));
        } else {
            // origin.deriv(NxK)[:, c0:] = delta'(NxM) * Amat(M x k-panel)
            const T *ap = A.stored_t ? A.ptr + c0 : A.ptr + c0 * A.ld;
            g_tag_a = dtag;
            g_tag_b = operand_tag(t, oa);
            window(g_win_b, A.ptr, K, c0, nc);
            KL(t, gemm_t<T>(S(t), true, A.stored_t, N, nc, M, (T)1, delta, M, ap, A.ld, beta, dst + c0 * N, N));
        }
    }
    return RDB_OK;
}

// The other accumulation into the same tape input that a pair of column-panelled GEMMs can be built with: the next `*`
// instruction (lower tape index) whose operand is rooted at `root`, provided running its adjoint NOW is the same computation as
// running it at its own turn: deriv(its output) must be final (no instruction in between, this one included, consumes that
// output) and reached.  Returns the instruction index and sets `side`, or -1.
static int find_mul_partner(rdb_tape *t, int k, int root, int64_t rows, int64_t cols, int &side) {
    for (int j = k - 1; j >= 0; --j) {
        InstrPlan &J = t->instrs[j];
        if (J.blk) {
            for (int r = 0; r < J.blk->n_ref; ++r) if (droot(t, J.blk->refbufs[r]) == root) return -1;
            continue;
        }
        for (int a = 0; a < J.rec.n_in; ++a) {
            OperandPlan &o = J.in[a];
            if (!o.rec.tracked || droot(t, o.rec.buffer) != root) continue;
            if (J.rec.opcode != OP_MUL || a > 1 || o.mode == OPM_GATHER || J.rev_done[a]) return -1;
            int64_t r2, c2;
            mul_root_shape(o, r2, c2);
            if (r2 != rows || c2 != cols) return -1;
            const int oj = droot(t, J.rec.out);
            if (t->dz[oj]) return -1;
            for (int i = j + 1; i <= k; ++i) {
                InstrPlan &X = t->instrs[i];
                if (X.blk) { for (int r = 0; r < X.blk->n_ref; ++r) if (droot(t, X.blk->refbufs[r]) == oj) return -1; continue; }
                for (int b = 0; b < X.rec.n_in; ++b) if (droot(t, X.in[b].rec.buffer) == oj) return -1;
                if (droot(t, X.rec.out) == oj) return -1;
            }
            side = a;
            return j;
        }
    }
    return -1;
}

// ---- slicing ahead (FP64 int8-slice GEMMs): the absmax/digit passes of the NEXT `*` GEMM's operands run on a side stream while
// the current GEMM computes; the GEMM then finds them in the sliced-operand cache (umma.cu, ozaki_preslice).
static bool preslice_on(rdb_tape *t) {
    static const bool on = [] { const char *e = getenv("RDB_PRESLICE"); return !(e && e[0] == '0'); }();
    return on && t->dtype == RDB_F64 && t->opts.gemm_f64_mode != 1 && !t->dmma_now && !t->capturing && !t->ctx->gemm.no_cache;
}
// the side stream may read everything the main stream has produced so far
static int preslice_begin(rdb_tape *t) {
    rdb_ctx *c = t->ctx;
    if (!c->side_stream) CU(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
    if (!c->side_ev) CU(cudaEventCreateWithFlags(&c->side_ev, cudaEventDisableTiming));
    CU(cudaEventRecord(c->side_ev, S(t)));
    CU(cudaStreamWaitEvent(c->side_stream, c->side_ev, 0));
    return RDB_OK;
}
// deriv(output) of a `*` as the adjoint GEMM of operand `side` reads it: side a -> rows m, k = n strided; side b -> rows n, k = m contiguous
static void preslice_delta(rdb_tape *t, InstrPlan &J, int side) {
    const int64_t M = J.in[0].rows, N = J.in[1].cols;
    if (!J.dtag) J.dtag = new_delta_tag();
    const double *d = (const double *)t->bufs[J.rec.out].der;
    int nl = 0;
    if (side == 0) ozaki_preslice(t->ctx->side_stream, d, M, M, N, false, t->opts.ozaki_slices, J.dtag, &nl);
    else ozaki_preslice(t->ctx->side_stream, d, M, N, M, true, t->opts.ozaki_slices, J.dtag, &nl);
    t->launches += nl;
}

// a.deriv += delta * value(b)'   and   b.deriv += value(a)' * delta     (arithmetic.jl:272-285), R seed columns
template <typename T>
static int mul_reverse(rdb_tape *t, int k, int64_t R) {
    InstrPlan &I = t->instrs[k];
    Buf &out = t->bufs[I.rec.out];
    OperandPlan &oa = I.in[0], &ob = I.in[1];
    const int64_t M = oa.rows, K = oa.cols, N = ob.cols;
    MatRef<T> A, B;
    mul_operand_ref<T>(t, oa, A);
    mul_operand_ref<T>(t, ob, B);
    const T *delta = (const T *)out.der;
    // gradient! with host results: when a GEMM is the LAST accumulation into a tape input, it is cut into column panels of the
    // result and each finished panel starts its D2H copy while the next one computes.  When TWO `*` accumulations are left and
    // the other one can run now (find_mul_partner), both are cut into the same panels and interleaved, so the input's adjoint
    // is final - and on its way to the host - panel by panel while the rest of the sweep still computes.
    auto &E = t->early;
    if (!I.dtag) I.dtag = new_delta_tag();
    const uint64_t dtag = I.dtag;
    if (R == 1 && preslice_on(t) && M >= 256 && N >= 256 && K >= 256) {
        // while this instruction's first adjoint GEMM runs: slice deriv(output) the way its second one reads it, and both views
        // of the next `*` instruction's deriv(output) if that adjoint is already final
        OK(preslice_begin(t));
        if (oa.rec.tracked && ob.rec.tracked && !I.rev_done[0] && !I.rev_done[1]) preslice_delta(t, I, 1);
        for (int j = k - 1; j >= 0; --j) {
            InstrPlan &J = t->instrs[j];
            if (J.blk || J.rec.opcode != OP_MUL) continue;
            const int oj = droot(t, J.rec.out);
            bool final_ = !t->dz[oj] && J.in[0].rows >= 256 && J.in[1].cols >= 256 && J.in[0].cols >= 256;
            for (int i = j + 1; i <= k && final_; ++i) {
                InstrPlan &X = t->instrs[i];
                if (X.blk) { final_ = false; break; }
                for (int b = 0; b < X.rec.n_in; ++b) if (droot(t, X.in[b].rec.buffer) == oj) final_ = false;
            }
            if (final_) {
                if (J.in[0].rec.tracked && !J.rev_done[0]) preslice_delta(t, J, 0);
                if (J.in[1].rec.tracked && !J.rev_done[1]) preslice_delta(t, J, 1);
            }
            break;                                                   // one instruction of lookahead
        }
    }
    auto panel_count = [&](int buffer, int64_t rows, int64_t cols, int remaining) -> int {
        const int rb = droot(t, buffer);
        if (E.enabled && R == 1 && E.dst[rb] && !E.issued[rb] && E.remaining[rb] == remaining && cols >= 2048 &&
            rows * cols * (int64_t)t->es >= (32 << 20) && E.n_ev + 4 <= kMaxArgs * 2)
            return 4;
        return 1;
    };
    auto panel_copy = [&](int buffer, int64_t first_elem, int64_t n_elems) -> int {
        const int rb = droot(t, buffer);
        cudaEvent_t &ev = E.ev[E.n_ev];
        if (!ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        E.n_ev++;
        CU(cudaEventRecord(ev, S(t)));
        CU(cudaStreamWaitEvent(t->ctx->copy_stream, ev, 0));
        CU(cudaMemcpyAsync((char *)E.dst[rb] + (size_t)first_elem * t->es, (char *)t->bufs[rb].der + (size_t)first_elem * t->es,
                           (size_t)n_elems * t->es, cudaMemcpyDeviceToHost, t->ctx->copy_stream));
        return RDB_OK;
    };
    // returns 1 if the accumulation was done by panels (alone or paired), 0 if the caller runs the plain GEMM
    auto by_panels = [&](int side, T beta) -> int {
        OperandPlan &o = I.in[side];
        if (o.mode == OPM_GATHER) return 0;
        int64_t rows, cols;
        mul_root_shape(o, rows, cols);
        const int root = droot(t, o.rec.buffer);
        int pj = -1, ps = 0;
        uint64_t ptag = 0;
        if (panel_count(o.rec.buffer, rows, cols, 2) > 1) {
            pj = find_mul_partner(t, k, root, rows, cols, ps);
            if (pj < 0) return 0;
            if (!t->instrs[pj].dtag) t->instrs[pj].dtag = new_delta_tag();
            ptag = t->instrs[pj].dtag;
        } else if (panel_count(o.rec.buffer, rows, cols, 1) == 1) return 0;
        int64_t width[4];
        plan_col_panels(rows, cols, width);
        int64_t c0 = 0;
        for (int pi = 0; pi < 4 && c0 < cols; ++pi) {
            if (width[pi] <= 0) continue;
            const int64_t nc = width[pi];
            OK(mul_acc_cols<T>(t, I, side, c0, nc, beta, dtag));
            if (pj >= 0) OK(mul_acc_cols<T>(t, t->instrs[pj], ps, c0, nc, (T)1, ptag));
            OK(panel_copy(o.rec.buffer, c0 * rows, nc * rows));
            c0 += nc;
        }
        if (c0 < cols) return fail(RDB_EINVAL, "internal: column panels do not cover the buffer");
        if (pj >= 0) t->instrs[pj].rev_done[ps] = true;
        E.issued[root] = 1; E.remaining[root] = 0;
        return 1;
    };
    if (oa.rec.tracked && I.rev_done[0]) I.rev_done[0] = false;          // accumulated earlier in this sweep as the partner of a pair
    else if (oa.rec.tracked) {
        Buf &ba = t->bufs[oa.rec.buffer];
        StepTimer tm(t, "gemm_reverse_a", (double)(M * N + K * N + 2 * M * K) * t->es * R, 2.0 * M * N * K * R);
        if (oa.mode == OPM_GATHER) OK(dmaterialize(t, oa.rec.buffer, R, true));
        const T beta_a = (oa.mode != OPM_GATHER && dtake(t, oa.rec.buffer)) ? (T)0 : (T)1;
        const int done = by_panels(0, beta_a);
        if (done < 0) return done;
        if (!done)
        for (int64_t r = 0; r < R; ++r) {
            const T *d = delta + r * M * N;
            if (oa.mode == OPM_PLAIN) {
                // a.deriv(MxK) += delta(MxN) * Bmat'(NxK)
                g_tag_b = operand_tag(t, ob);
                if (R == 1) g_tag_a = dtag;
                KL(t, gemm_t<T>(S(t), false, !B.stored_t, M, K, N, (T)1, d, M, B.ptr, B.ld, beta_a, (T *)ba.der + r * ba.size, M));
            } else if (oa.mode == OPM_FOLDED_T) {
                // scatter-add through the transpose map == origin.deriv(KxM) += Bmat(KxN) * delta'(NxM)
                g_tag_a = operand_tag(t, ob);
                if (R == 1) g_tag_b = dtag;
                KL(t, gemm_t<T>(S(t), B.stored_t, true, K, M, N, (T)1, B.ptr, B.ld, d, M, beta_a, (T *)ba.der + r * ba.size, K));
            } else {
                g_tag_b = operand_tag(t, ob);
                KL(t, gemm_t<T>(S(t), false, !B.stored_t, M, K, N, (T)1, d, M, B.ptr, B.ld, (T)0, (T *)oa.dense_der, M));
                KL(t, launch_scatter_add<T>(S(t), (T *)ba.der + r * ba.size, ba.size, oa.d_map, (const T *)oa.dense_der, oa.size, 1));
            }
        }
        OK(acc_done(t, oa.rec.buffer));
    }
    if (ob.rec.tracked && I.rev_done[1]) I.rev_done[1] = false;
    else if (ob.rec.tracked) {
        Buf &bb = t->bufs[ob.rec.buffer];
        StepTimer tm(t, "gemm_reverse_b", (double)(M * N + M * K + 2 * K * N) * t->es * R, 2.0 * M * N * K * R);
        if (ob.mode == OPM_GATHER) OK(dmaterialize(t, ob.rec.buffer, R, true));
        const T beta_b = (ob.mode != OPM_GATHER && dtake(t, ob.rec.buffer)) ? (T)0 : (T)1;
        const int done = by_panels(1, beta_b);
        if (done < 0) return done;
        if (done) {
        } else if (ob.mode == OPM_PLAIN) {
            // b.deriv(K x N*R) += Amat'(KxM) * delta(M x N*R): the R seed columns ride along as extra GEMM columns
            g_tag_a = operand_tag(t, oa);
            if (R == 1) g_tag_b = dtag;
            KL(t, gemm_t<T>(S(t), !A.stored_t, false, K, N * R, M, (T)1, A.ptr, A.ld, delta, M, beta_b, (T *)bb.der, K));
        } else {
            for (int64_t r = 0; r < R; ++r) {
                const T *d = delta + r * M * N;
                if (ob.mode == OPM_FOLDED_T) {
                    // origin.deriv(NxK) += delta'(NxM) * Amat(MxK)
                    g_tag_b = operand_tag(t, oa);
                    if (R == 1) g_tag_a = dtag;
                    KL(t, gemm_t<T>(S(t), true, A.stored_t, N, K, M, (T)1, d, M, A.ptr, A.ld, beta_b, (T *)bb.der + r * bb.size, N));
                } else {
                    KL(t, gemm_t<T>(S(t), !A.stored_t, false, K, N, M, (T)1, A.ptr, A.ld, d, M, (T)0, (T *)ob.dense_der, K));
                    KL(t, launch_scatter_add<T>(S(t), (T *)bb.der + r * bb.size, bb.size, ob.d_map, (const T *)ob.dense_der, ob.size, 1));
                }
            }
        }
        OK(acc_done(t, ob.rec.buffer));
    }
    return RDB_OK;
}

template <typename T>
static int expr_reverse(rdb_tape *t, InstrPlan &I, int64_t R, const T *sdelta = nullptr, T sscale = (T)1, bool probe = false) {
    // sdelta != nullptr: deriv(output)[e + r n] = sscale * sdelta[r] for every e - the `sum`/`mean` that consumes this instruction's
    // output would have written exactly that (reverse {sum, elementwise} in one step); only the patterns below take it.
    // probe: only report (1 / 0) whether the scalar-delta form is available for this instruction.
    Buf &out = t->bufs[I.rec.out];
    const T *delta = (const T *)out.der;
    if (sdelta || probe) {
        bool ok = t->opts.fuse_elementwise != 0;
        int n_same = 0, n_col = 0, n_tracked = 0;
        for (int a = 0; a < I.rec.n_in && ok; ++a) {
            OperandPlan &o = I.in[a];
            if (!o.rec.tracked) continue;
            ++n_tracked;
            if (o.mode != OPM_PLAIN) { ok = false; break; }
            const int64_t d0 = o.rec.ndims >= 1 ? o.rec.dims[0] : 1;
            bool same = true;
            for (int d = 0; d < kMaxDims; ++d) same &= (d < o.rec.ndims ? o.rec.dims[d] : 1) == out.dims[d];
            if (same) ++n_same;
            else if (o.size == d0 && d0 == out.dims[0] && d0 > 1 && out.nd >= 2) ++n_col;
            else ok = false;
        }
        // every tracked argument of the output's shape, or the {same shape, column vector} pair of two different buffers
        ok = ok && n_tracked > 0 && (n_col == 0 || (n_col == 1 && n_same == 1 && n_tracked == 2));
        if (ok && n_col == 1) {
            int ia = -1, ib = -1;
            for (int a = 0; a < I.rec.n_in; ++a) if (I.in[a].rec.tracked) { if (I.in[a].size == out.size) ia = a; else ib = a; }
            ok = droot(t, I.in[ia].rec.buffer) != droot(t, I.in[ib].rec.buffer);
        }
        if (probe) return ok ? 1 : 0;
        if (!ok) return fail(RDB_EINVAL, "internal: scalar-delta elementwise reverse on an unsupported pattern");
    }
    // {broadcast adjoint, bias column-reduce} in one pass: exactly one tracked argument of the output's shape and one tracked column
    // vector clamped along the columns (`tanh.(W*X .+ b)`): delta is read once for both accumulations (5 S E -> 4 S E).  Two different
    // buffers, so the order of accumulation into each of them is the reference's.
    if (t->opts.fuse_elementwise && out.nd >= 2) {
        int ia = -1, ib = -1, n_tracked = 0;
        for (int a = 0; a < I.rec.n_in; ++a) {
            OperandPlan &o = I.in[a];
            if (!o.rec.tracked) continue;
            ++n_tracked;
            if (o.mode != OPM_PLAIN) { ia = ib = -2; break; }
            const int64_t d0 = o.rec.ndims >= 1 ? o.rec.dims[0] : 1;
            bool same = true;
            for (int d = 0; d < kMaxDims; ++d) same &= (d < o.rec.ndims ? o.rec.dims[d] : 1) == out.dims[d];
            if (same && ia == -1) ia = a;
            else if (!same && o.size == d0 && d0 == out.dims[0] && d0 > 1 && ib == -1) ib = a;
            else { ia = ib = -2; break; }
        }
        if (ia >= 0 && ib >= 0 && n_tracked == 2 && droot(t, I.in[ia].rec.buffer) != droot(t, I.in[ib].rec.buffer)) {
            Buf &bs = t->bufs[I.in[ia].rec.buffer], &bc = t->bufs[I.in[ib].rec.buffer];
            const int64_t d0 = out.dims[0];
            const bool ow_same = dtake(t, I.in[ia].rec.buffer), ow_col = dtake(t, I.in[ib].rec.buffer);
            StepTimer tm(t, sdelta ? "sum+expr_reverse+colreduce" : "expr_reverse+colreduce", ((ow_same ? 3.0 : 4.0) * R + (sdelta ? 0.0 : 1.0)) * out.size * t->es, 0);
            KL(t, launch_mul_acc_colvec_fused<T>(S(t), (T *)bc.der, sdelta ? sdelta : delta, (const T *)I.partial[ib], d0, out.size / d0, R, (T *)t->scratch,
                                                 t->scratch_bytes / 8, ow_col, (T *)bs.der, (const T *)I.partial[ia], ow_same, sdelta != nullptr, sscale));
            t->launches++;
            OK(acc_done(t, I.in[ia].rec.buffer));
            OK(acc_done(t, I.in[ib].rec.buffer));
            return RDB_OK;
        }
    }
    for (int a = 0; a < I.rec.n_in; ++a) {
        OperandPlan &o = I.in[a];
        if (!o.rec.tracked) continue;
        Buf &b = t->bufs[o.rec.buffer];
        if (o.mode != OPM_PLAIN) return fail(RDB_EUNSUPPORTED, "elementwise instruction on an index-mapped operand");
        bool same = true;
        for (int d = 0; d < kMaxDims; ++d) {
            int64_t od = d < o.rec.ndims ? o.rec.dims[d] : 1;
            if (od != out.dims[d]) same = false;
        }
        if (same) {
            const bool ow = dtake(t, o.rec.buffer);
            StepTimer tm(t, sdelta ? "sum+expr_reverse" : "expr_reverse", ((ow ? 2.0 : 3.0) * R + (sdelta ? 0.0 : 1.0)) * out.size * t->es, 0);
            if (sdelta) KL(t, launch_mul_acc_sdelta<T>(S(t), (T *)b.der, sdelta, sscale, (const T *)I.partial[a], out.size, R, ow));
            else KL(t, launch_mul_acc<T>(S(t), (T *)b.der, delta, (const T *)I.partial[a], out.size, R, ow));
        } else if (o.size == 1) {
            // scalar (TrackedReal) argument: input_deriv += sum_i x[i]*partial[i]   (propagation.jl:98-108; unbroadcast(::Number) = sum)
            StepTimer tm(t, "expr_reverse_scalar", 2.0 * out.size * R * t->es, 0);
            const bool ow = dtake(t, o.rec.buffer);
            for (int64_t r = 0; r < R; ++r) {
                KL(t, launch_dot_block_sums<T>(S(t), delta + r * out.size, (const T *)I.partial[a], out.size, (T *)t->scratch));
                KL(t, launch_finish_sum<T>(S(t), (const T *)t->scratch, (T *)b.der + r, (T)1, !ow));
            }
        } else {
            int64_t d0 = o.rec.ndims >= 1 ? o.rec.dims[0] : 1;
            bool colvec = (o.size == d0) && (d0 == out.dims[0]) && d0 > 1;
            if (colvec) {
                StepTimer tm(t, "expr_reverse_colreduce", 2.0 * out.size * R * t->es, 0);
                KL(t, launch_mul_acc_colvec<T>(S(t), (T *)b.der, delta, (const T *)I.partial[a], d0, out.size / d0, R,
                                               (T *)t->scratch, t->scratch_bytes / 8, dtake(t, o.rec.buffer)));
                t->launches++;
            } else {
                int64_t dst_stride[kMaxDims], st = 1;
                for (int d = 0; d < kMaxDims; ++d) {
                    int64_t od = d < o.rec.ndims ? o.rec.dims[d] : 1;
                    dst_stride[d] = (od == 1 && out.dims[d] != 1) ? 0 : st;
                    st *= od;
                }
                StepTimer tm(t, "expr_reverse_generic", 3.0 * out.size * R * t->es, 0);
                OK(dmaterialize(t, o.rec.buffer, R, true));
                KL(t, launch_mul_acc_generic<T>(S(t), (T *)b.der, b.size, dst_stride, delta, (const T *)I.partial[a], out.dims,
                                                out.size, R));
            }
        }
        OK(acc_done(t, o.rec.buffer));
    }
    return RDB_OK;
}

// scalar_reverse_exec! for a whole run of scalar instructions (scalars.jl:52-74), R seeds at once
template <typename T>
static int scalar_reverse(rdb_tape *t, InstrPlan &I, int64_t R) {
    ScalarBlock &B = *I.blk;
    bool reached = false;
    for (int k = 0; k < B.n_ref; ++k)
        if (B.ref_has_export[k] && !t->dz[droot(t, B.refbufs[k])]) reached = true;
    if (!reached) return RDB_OK;                           // no adjoint reached any exported slot: every increment would add zero
    bool leaves_zero = true;                               // pull_deriv! of a logically zero origin loads zeros: skip the loads
    for (int k = 0; k < B.n_ref; ++k)
        if (B.ref_has_leaf[k] && !t->dz[droot(t, B.refbufs[k])]) leaves_zero = false;
    for (int k = 0; k < B.n_ref; ++k)                      // the kernels touch single elements: the storage must hold real zeros
        OK(dmaterialize(t, B.refbufs[k], R, true));
    {
        StepTimer tm(t, "scalar_reverse", (double)(B.n + B.n_edges + 2 * B.n_leaf) * R * t->es, 0);
        int nl = 0;
        KL(t, scalar_block_reverse<T>(S(t), B, R, leaves_zero, &nl));
        t->launches += nl - 1;
    }
    for (int k = 0; k < B.n_ref; ++k) {
        if (B.ref_has_export[k] && !B.ref_has_leaf[k] && B.ref_export_count[k] == t->bufs[B.refbufs[k]].size) dclear(t, B.refbufs[k]);
        if (B.ref_has_leaf[k]) OK(acc_done(t, B.refbufs[k]));
    }
    return RDB_OK;
}

template <typename T>
static int reverse_pass(rdb_tape *t, int64_t R) {
    for (auto &I : t->instrs) { I.rev_done[0] = I.rev_done[1] = false; I.dtag = 0; }
    for (size_t k = t->instrs.size(); k-- > 0;) {
        InstrPlan &I = t->instrs[k];
        if (I.rec.opcode == OP_SCALAR_BLOCK) { OK(scalar_reverse<T>(t, I, R)); continue; }
        Buf &out = t->bufs[I.rec.out];
        const T *delta = (const T *)out.der;
        if (t->dz[droot(t, I.rec.out)]) continue;        // no adjoint reached this output: every increment would add zero
        switch (I.rec.opcode) {
            case OP_MUL: OK(mul_reverse<T>(t, (int)k, R)); break;
            case OP_ADD:
            case OP_SUB:
                for (int a = 0; a < 2; ++a) {
                    if (!I.in[a].rec.tracked) continue;
                    Buf &b = t->bufs[I.in[a].rec.buffer];
                    T sign = (a == 1 && I.rec.opcode == OP_SUB) ? (T)-1 : (T)1;
                    const bool ow = dtake(t, I.in[a].rec.buffer);
                    StepTimer tm(t, "add_reverse", (ow ? 2.0 : 3.0) * out.size * R * t->es, 0);
                    KL(t, launch_axpy<T>(S(t), (T *)b.der, delta, sign, out.size * R, ow));
                    OK(acc_done(t, I.in[a].rec.buffer));
                }
                break;
            case OP_NEG:
                if (I.in[0].rec.tracked) {
                    Buf &b = t->bufs[I.in[0].rec.buffer];
                    KL(t, launch_axpy<T>(S(t), (T *)b.der, delta, (T)-1, out.size * R, dtake(t, I.in[0].rec.buffer)));
                    OK(acc_done(t, I.in[0].rec.buffer));
                }
                break;
            case OP_SUM:
            case OP_MEAN:
                if (I.in[0].rec.tracked) {
                    Buf &b = t->bufs[I.in[0].rec.buffer];
                    T scale = I.rec.opcode == OP_MEAN ? (T)1 / (T)I.in[0].size : (T)1;
                    // reverse {sum, +} in one step: the instruction right before is the array +/- that produced the summed buffer, and
                    // nothing has accumulated into that buffer's deriv yet.  x.deriv += scale * delta goes straight into the
                    // operands of the +/- (the same values in the same order as via deriv(a + b): 1 * (scale * delta)), deriv(a + b)
                    // stays logically zero - it would be unseeded right after anyway (arithmetic.jl:45-53) - and the +/- itself is then
                    // skipped by the "no adjoint reached this output" test.  cfg 2: 2 M E written instead of 5 M E moved.
                    if (t->opts.fuse_elementwise && k > 0 && t->dz[droot(t, I.in[0].rec.buffer)] && I.in[0].rec.cls == CLS_TA && I.in[0].mode == OPM_PLAIN) {
                        InstrPlan &A = t->instrs[k - 1];
                        const bool pm = !A.blk && (A.rec.opcode == OP_ADD || A.rec.opcode == OP_SUB) && A.rec.out == I.in[0].rec.buffer;
                        bool plain = pm;
                        for (int a = 0; a < 2 && plain; ++a)
                            if (A.in[a].rec.tracked && (A.in[a].rec.cls != CLS_TA || A.in[a].mode != OPM_PLAIN || t->bufs[A.in[a].rec.buffer].size != b.size)) plain = false;
                        if (plain) {
                            for (int a = 0; a < 2; ++a) {
                                if (!A.in[a].rec.tracked) continue;
                                Buf &ba = t->bufs[A.in[a].rec.buffer];
                                const T sg = (a == 1 && A.rec.opcode == OP_SUB) ? -scale : scale;
                                const bool ow = dtake(t, A.in[a].rec.buffer);
                                StepTimer tm(t, "sum+add_reverse", (ow ? 1.0 : 2.0) * ba.size * R * t->es, 0);
                                KL(t, launch_acc_scalar<T>(S(t), (T *)ba.der, ba.size, R, delta, sg, ow));
                                if (ow) dmark_uniform(t, A.in[a].rec.buffer, R);
                                OK(acc_done(t, A.in[a].rec.buffer));
                            }
                            break;
                        }
                    }
                    // reverse {sum, elementwise} in one step, same idea: the producer of the summed buffer is the elementwise instruction
                    // right before, nothing else has accumulated into its output's deriv: its reverse exec takes the scalar adjoint
                    // (scale * delta[r], the very value the fill would have stored) and deriv(output) stays logically zero
                    if (t->opts.fuse_elementwise && k > 0 && t->dz[droot(t, I.in[0].rec.buffer)] && I.in[0].rec.cls == CLS_TA && I.in[0].mode == OPM_PLAIN) {
                        InstrPlan &A = t->instrs[k - 1];
                        if (!A.blk && is_elementwise(A.rec.opcode) && A.rec.out == I.in[0].rec.buffer && expr_reverse<T>(t, A, R, nullptr, (T)1, true) == 1) {
                            OK(expr_reverse<T>(t, A, R, delta, scale, false));
                            break;
                        }
                    }
                    const bool ow = dtake(t, I.in[0].rec.buffer);
                    StepTimer tm(t, "sum_reverse", (ow ? 1.0 : 2.0) * b.size * R * t->es, 0);
                    KL(t, launch_acc_scalar<T>(S(t), (T *)b.der, b.size, R, delta, scale, ow));
                    if (ow) dmark_uniform(t, I.in[0].rec.buffer, R);
                    OK(acc_done(t, I.in[0].rec.buffer));
                }
                break;
            case OP_NBCAST: case OP_MAP: case OP_BCAST: case OP_BCAST_ADD: case OP_BCAST_SUB: case OP_BCAST_MUL:
            case OP_BCAST_DIV: case OP_BCAST_LDIV: case OP_BCAST_POW: case OP_TRACKER_BCAST:
                OK(expr_reverse<T>(t, I, R));
                break;
            case OP_DOT:
                for (int a = 0; a < 2; ++a) {
                    if (!I.in[a].rec.tracked) continue;
                    Buf &b = t->bufs[I.in[a].rec.buffer];
                    Buf &other = t->bufs[I.in[1 - a].rec.buffer];
                    StepTimer tm(t, "dot_reverse", 3.0 * b.size * R * t->es, 0);
                    KL(t, launch_scal_acc<T>(S(t), (T *)b.der, (const T *)other.val, delta, b.size, R, dtake(t, I.in[a].rec.buffer)));
                    OK(acc_done(t, I.in[a].rec.buffer));
                }
                break;
            case OP_SUMDIMS:
                if (I.in[0].rec.tracked) {
                    Buf &b = t->bufs[I.in[0].rec.buffer];
                    int64_t dstride[kMaxDims], st = 1;
                    for (int d = 0; d < kMaxDims; ++d) { dstride[d] = (out.dims[d] == 1 && b.dims[d] != 1) ? 0 : st; st *= out.dims[d]; }
                    StepTimer tm(t, "sumdims_reverse", 2.0 * b.size * R * t->es, 0);
                    KL(t, launch_bcast_acc<T>(S(t), (T *)b.der, b.dims, dstride, delta, b.size, R, out.size, dtake(t, I.in[0].rec.buffer)));
                    OK(acc_done(t, I.in[0].rec.buffer));
                }
                break;
            case OP_GETINDEX: case OP_GETINDEX_GENERIC:
                if (I.in[0].rec.tracked) {
                    Buf &b = t->bufs[I.in[0].rec.buffer];
                    StepTimer tm(t, "getindex_reverse", 3.0 * out.size * R * t->es, 0);
                    OK(dmaterialize(t, I.in[0].rec.buffer, R, true));
                    KL(t, launch_scatter_add<T>(S(t), (T *)b.der, b.size, I.in[0].d_map, delta, out.size, R));
                    OK(acc_done(t, I.in[0].rec.buffer));
                }
                break;
            default:
                return fail(RDB_EUNSUPPORTED, "opcode %d has no reverse kernel", I.rec.opcode);
        }
        dclear(t, I.rec.out);                              // unseed!(output)
    }
    return RDB_OK;
}

// ------------------------------------------------------------------------------------------ allocation
static int ensure_batch(rdb_tape *t, int64_t R) {
    if (R <= t->R_alloc) return RDB_OK;
    for (auto &b : t->bufs) {
        if (b.kind == BUF_CONST || b.alias_of >= 0) continue;
        dev_free(t, b.der);
        b.der = nullptr;
        OK(dev_alloc(t, &b.der, b.size * R * (int64_t)t->es));
        CU(cudaMemsetAsync(b.der, 0, (size_t)(b.size * R) * t->es, S(t)));
    }
    for (auto &b : t->bufs)
        if (b.alias_of >= 0) b.der = t->bufs[b.alias_of].der;
    for (auto &I : t->instrs) {
        if (!I.blk) continue;
        ScalarBlock &B = *I.blk;
        if (scalar_block_ensure_R(B, R, t->es) != cudaSuccess) {
            cudaGetLastError();
            return fail(RDB_ENOMEM, "scalar block: cannot allocate %lld x %lld adjoints", (long long)B.n, (long long)R);
        }
        for (int k = 0; k < B.n_ref; ++k) {
            Buf &b = t->bufs[B.refbufs[k]];
            B.h_ext_val[k] = b.val; B.h_ext_der[k] = b.der; B.h_ext_size[k] = b.size;
        }
        CU(scalar_block_upload_ext(B, S(t)));
    }
    // column-reduction scratch grows with R
    int64_t need = t->scratch_bytes;
    for (auto &I : t->instrs)
        if (is_elementwise(I.rec.opcode))
            for (int a = 0; a < I.rec.n_in; ++a)
                if (I.in[a].rec.tracked && I.in[a].size != t->bufs[I.rec.out].size)
                    need = std::max<int64_t>(need, I.in[a].size * R * 64 * 8);
    if (need > t->scratch_bytes) {
        dev_free(t, t->scratch);
        OK(dev_alloc(t, &t->scratch, need));
        t->scratch_bytes = need;
    }
    t->R_alloc = R;
    return RDB_OK;
}

static int ensure_stage(rdb_tape *t, int64_t bytes) {
    if (bytes <= t->stage_bytes) return RDB_OK;
    dev_free(t, t->stage);
    t->stage = nullptr;
    OK(dev_alloc(t, &t->stage, bytes));
    t->stage_bytes = bytes;
    return RDB_OK;
}

static int64_t auto_seed_block(rdb_tape *t, int64_t rows, bool tangents) {
    if (t->opts.seed_block > 0) return std::min<int64_t>(rows, t->opts.seed_block);
    int64_t per_seed = 0;
    for (auto &b : t->bufs)
        if (b.kind != BUF_CONST && b.alias_of < 0) per_seed += b.size * (int64_t)t->es * (tangents ? 2 : 1);
    for (auto &I : t->instrs)
        if (I.blk) per_seed += (I.blk->n_grow + 1) * (int64_t)t->es;
    int64_t budget = 24LL << 30;   // 24 GB of the 180 GB for seed-batched adjoints
    int64_t R = std::max<int64_t>(1, budget / std::max<int64_t>(per_seed, 1));
    // scalar-instruction tapes: the streaming reverse kernel costs the same for any number of seeds up to ~19k (one warp per 32
    // seeds, 4 x 148 schedulers), so take them all at once; array tapes: GEMM batches of 8192 columns are already at peak
    bool scalar_only = !t->instrs.empty();
    for (auto &I : t->instrs) if (!I.blk) scalar_only = false;
    R = std::min<int64_t>(R, scalar_only ? 32768 : 8192);
    return std::min(R, rows);
}

// ------------------------------------------------------------------------------------------ loader
static bool is_transpose_map(const int64_t *m, int64_t r, int64_t c) {
    // operand is r x c, origin is c x r: operand (i,j) at k = i + j*r  ->  origin index j + i*c
    for (int64_t j = 0; j < c; ++j)
        for (int64_t i = 0; i < r; ++i)
            if (m[i + j * r] != j + i * c) return false;
    return true;
}
static bool is_identity_map(const int64_t *m, int64_t n) {
    for (int64_t i = 0; i < n; ++i)
        if (m[i] != i) return false;
    return true;
}

static int load_program(rdb_tape *t, const uint8_t *raw, size_t nbytes) {
    if (nbytes < sizeof(Header)) return fail(RDB_EINVAL, "tape program too short");
    Header h;
    memcpy(&h, raw, sizeof(h));
    if (h.magic != kMagic) return fail(RDB_EINVAL, "not a tape program (bad magic)");
    if (h.version != kVersion) return fail(RDB_EINVAL, "tape program version %u, engine expects %u", h.version, kVersion);
    if (h.total_bytes != nbytes) return fail(RDB_EINVAL, "tape program is truncated (%llu != %zu)", (unsigned long long)h.total_bytes, nbytes);
    if (h.dtype != RDB_F64 && h.dtype != RDB_F32) return fail(RDB_EINVAL, "unknown dtype %u", h.dtype);
    auto in_range = [&](uint64_t off, uint64_t bytes) { return off <= nbytes && bytes <= nbytes - off; };
    if (!in_range(h.off_buffers, (uint64_t)h.n_buffers * sizeof(BufferRec)) || !in_range(h.off_inputs, (uint64_t)h.n_inputs * 4) ||
        !in_range(h.off_instr, (uint64_t)h.n_instr * sizeof(InstrRec)) || !in_range(h.off_idx, h.n_idx * 8) ||
        !in_range(h.off_expr, h.n_expr * 4) || !in_range(h.off_fconst, h.n_fconst * 8) || !in_range(h.off_const, h.n_const))
        return fail(RDB_EINVAL, "tape program section out of range");
    t->dtype = (int)h.dtype;
    t->es = h.dtype == RDB_F64 ? 8 : 4;
    const BufferRec *brec = (const BufferRec *)(raw + h.off_buffers);
    const int32_t *inputs = (const int32_t *)(raw + h.off_inputs);
    const InstrRec *irec = (const InstrRec *)(raw + h.off_instr);
    const int64_t *idx = (const int64_t *)(raw + h.off_idx);
    const int32_t *expr = (const int32_t *)(raw + h.off_expr);
    const double *fconst = (const double *)(raw + h.off_fconst);
    const uint8_t *consts = raw + h.off_const;

    // buffers ------------------------------------------------------------------------------------------
    t->bufs.resize(h.n_buffers);
    for (uint32_t i = 0; i < h.n_buffers; ++i) {
        Buf &b = t->bufs[i];
        const BufferRec &r = brec[i];
        if (r.ndims < 0 || r.ndims > kMaxDims) return fail(RDB_EINVAL, "buffer %u: bad ndims", i);
        b.kind = r.kind; b.nd = r.ndims; b.alias_of = r.alias_of; b.size = 1;
        for (int d = 0; d < kMaxDims; ++d) {
            b.dims[d] = d < r.ndims ? r.dims[d] : 1;
            if (b.dims[d] < 0) return fail(RDB_EINVAL, "buffer %u: negative dim", i);
            b.size *= b.dims[d];
        }
        if (b.kind != BUF_TA && b.kind != BUF_TR && b.kind != BUF_CONST) return fail(RDB_EINVAL, "buffer %u: bad kind", i);
        if (b.alias_of >= 0) {
            if ((uint32_t)b.alias_of >= i || t->bufs[b.alias_of].size != b.size) return fail(RDB_EINVAL, "buffer %u: bad alias", i);
            b.val = t->bufs[b.alias_of].val;   // reshape shares value AND deriv (tracked.jl:402-408)
            continue;
        }
        OK(dev_alloc(t, &b.val, b.size * (int64_t)t->es));
        if (b.kind == BUF_CONST) {
            if (r.const_offset < 0 || !in_range(h.off_const + r.const_offset, b.size * t->es)) return fail(RDB_EINVAL, "buffer %u: const data out of range", i);
            CU(cudaMemcpyAsync(b.val, consts + r.const_offset, (size_t)b.size * t->es, cudaMemcpyHostToDevice, S(t)));
        } else {
            CU(cudaMemsetAsync(b.val, 0, (size_t)b.size * t->es, S(t)));
        }
    }
    for (uint32_t i = 0; i < h.n_inputs; ++i) {
        if (inputs[i] < 0 || (uint32_t)inputs[i] >= h.n_buffers || t->bufs[inputs[i]].kind != BUF_TA) return fail(RDB_EINVAL, "bad input buffer id");
        t->inputs.push_back(inputs[i]);
    }
    if (h.output_buffer < 0 || (uint32_t)h.output_buffer >= h.n_buffers) return fail(RDB_EINVAL, "bad output buffer id");
    t->output = h.output_buffer;

    t->h_fconst.assign(fconst, fconst + h.n_fconst);
    if (h.n_fconst) {
        OK(dev_alloc(t, (void **)&t->d_fconst, h.n_fconst * 8));
        CU(cudaMemcpyAsync(t->d_fconst, fconst, h.n_fconst * 8, cudaMemcpyHostToDevice, S(t)));
    }

    // instructions ---------------------------------------------------------------------------------------
    // a record with more than kWireArgs operands is followed by OP_MORE_OPERANDS records holding the rest
    std::vector<uint32_t> heads;
    for (uint32_t r = 0; r < h.n_instr;) {
        const InstrRec &R = irec[r];
        if (R.opcode == OP_MORE_OPERANDS) return fail(RDB_EINVAL, "record %u: operand continuation without an instruction", r);
        if (R.n_in < 1 || R.n_in > kMaxArgs) return fail(RDB_EINVAL, "record %u: bad input count", r);
        heads.push_back(r);
        uint32_t extra = R.n_in > kWireArgs ? (uint32_t)((R.n_in - 1) / kWireArgs) : 0;
        for (uint32_t e = 1; e <= extra; ++e) {
            const int want = std::min(kWireArgs, R.n_in - (int)e * kWireArgs);
            if (r + e >= h.n_instr || irec[r + e].opcode != OP_MORE_OPERANDS || irec[r + e].n_in != want)
                return fail(RDB_EINVAL, "record %u: missing or malformed operand continuation", r);
        }
        r += 1 + extra;
    }
    t->instrs.resize(heads.size());
    for (uint32_t i = 0; i < (uint32_t)heads.size(); ++i) {
        InstrPlan &I = t->instrs[i];
        const uint32_t r0 = heads[i];
        I.rec = irec[r0];
        const int op = I.rec.opcode;
        if (I.rec.out < 0 || (uint32_t)I.rec.out >= h.n_buffers || t->bufs[I.rec.out].kind == BUF_CONST) return fail(RDB_EINVAL, "instruction %u: bad output", i);
        switch (op) {
            case OP_MUL: case OP_ADD: case OP_SUB: case OP_DOT: if (I.rec.n_in != 2) return fail(RDB_EINVAL, "instruction %u: needs 2 inputs", i); break;
            case OP_NEG: case OP_SUM: case OP_MEAN: case OP_GETINDEX: case OP_GETINDEX_GENERIC: case OP_SUMDIMS:
                if (I.rec.n_in != 1) return fail(RDB_EINVAL, "instruction %u: needs 1 input", i);
                break;
            case OP_NBCAST: case OP_MAP: case OP_BCAST: case OP_BCAST_ADD: case OP_BCAST_SUB: case OP_BCAST_MUL:
            case OP_BCAST_DIV: case OP_BCAST_LDIV: case OP_BCAST_POW: case OP_TRACKER_BCAST: break;
            case OP_SCALAR_BLOCK: if (I.rec.n_in != 1) return fail(RDB_EINVAL, "instruction %u: scalar block carries one payload operand", i); break;
            default:
                // @grad / ChainRules closures, det, inv, cat, ... cannot be lowered
                return fail(RDB_EUNSUPPORTED, "instruction %u: opcode %d cannot be replayed on the device (no CPU fallback)", i, op);
        }
        Buf &out = t->bufs[I.rec.out];
        for (int a = 0; a < I.rec.n_in; ++a) {
            OperandPlan &o = I.in[a];
            o.rec = irec[r0 + a / kWireArgs].in[a % kWireArgs];
            if (o.rec.buffer < 0 || (uint32_t)o.rec.buffer >= h.n_buffers) return fail(RDB_EINVAL, "instruction %u: bad operand buffer", i);
            if (o.rec.ndims < 0 || o.rec.ndims > kMaxDims) return fail(RDB_EINVAL, "instruction %u: bad operand ndims", i);
            Buf &b = t->bufs[o.rec.buffer];
            if (o.rec.tracked && b.kind == BUF_CONST) return fail(RDB_EINVAL, "instruction %u: tracked CONST operand", i);
            b.n_consumers++;
            o.size = 1;
            for (int d = 0; d < o.rec.ndims; ++d) o.size *= o.rec.dims[d];
            o.rows = o.rec.ndims >= 1 ? o.rec.dims[0] : 1;
            o.cols = o.rec.ndims >= 2 ? o.rec.dims[1] : 1;
            const bool indexed = (o.rec.cls == CLS_IDX) || is_getindex(op);
            if (!indexed) {
                if (o.size != b.size) return fail(RDB_EINVAL, "instruction %u: operand/buffer size mismatch", i);
                continue;
            }
            if (o.rec.idx_kind == IDX_TRANSPOSE) {
                if (b.nd != 2 || b.dims[0] != o.cols || b.dims[1] != o.rows) return fail(RDB_EINVAL, "instruction %u: transpose tag on a non-matching origin", i);
                if (op != OP_MUL) return fail(RDB_EUNSUPPORTED, "instruction %u: transposed operand outside `*`", i);
                o.mode = OPM_FOLDED_T;
                continue;
            }
            if (o.rec.idx_kind != IDX_EXPLICIT) return fail(RDB_EINVAL, "instruction %u: indexed operand without a map", i);
            const int64_t expect = is_getindex(op) ? out.size : o.size;
            if (o.rec.idx_offset < 0 || o.rec.idx_len != expect || (uint64_t)(o.rec.idx_offset + o.rec.idx_len) > h.n_idx)
                return fail(RDB_EINVAL, "instruction %u: index map out of range", i);
            const int64_t *m = idx + o.rec.idx_offset;
            for (int64_t k = 0; k < o.rec.idx_len; ++k)
                if (m[k] < 0 || m[k] >= b.size) return fail(RDB_EINVAL, "instruction %u: index map entry out of bounds", i);
            if (op == OP_MUL && o.rec.ndims == 2 && b.size == o.size &&
                ((b.nd == 2 && b.dims[0] == o.cols && b.dims[1] == o.rows) || (b.nd == 1 && o.rows == 1)) &&
                is_transpose_map(m, o.rows, o.cols)) {
                o.mode = OPM_FOLDED_T;     // pure transpose: folded into the GEMM operand flag
                continue;
            }
            if (op == OP_MUL && b.size == o.size && is_identity_map(m, o.size)) { o.mode = OPM_PLAIN; continue; }
            o.mode = OPM_GATHER;
            OK(dev_alloc(t, (void **)&o.d_map, o.rec.idx_len * 8));
            CU(cudaMemcpyAsync(o.d_map, m, (size_t)o.rec.idx_len * 8, cudaMemcpyHostToDevice, S(t)));
            if (op == OP_MUL) {
                OK(dev_alloc(t, &o.dense_val, o.size * (int64_t)t->es));
                if (o.rec.tracked) OK(dev_alloc(t, &o.dense_der, o.size * (int64_t)t->es));
            } else if (!is_getindex(op)) {
                return fail(RDB_EUNSUPPORTED, "instruction %u: index-mapped operand outside `*`/getindex", i);
            }
        }
        // shape checks + per-instruction caches
        if (op == OP_MUL) {
            if (I.in[0].cols != I.in[1].rows || out.size != I.in[0].rows * I.in[1].cols) return fail(RDB_EINVAL, "instruction %u: `*` shape mismatch", i);
        } else if (op == OP_ADD || op == OP_SUB || op == OP_NEG) {
            for (int a = 0; a < I.rec.n_in; ++a)
                if (I.in[a].size != out.size || I.in[a].mode != OPM_PLAIN) return fail(RDB_EINVAL, "instruction %u: array +/- shape mismatch", i);
        } else if (op == OP_SUM || op == OP_MEAN) {
            if (out.size != 1 || I.in[0].mode != OPM_PLAIN) return fail(RDB_EINVAL, "instruction %u: sum/mean shape mismatch", i);
        } else if (is_elementwise(op)) {
            if (I.rec.expr_len < 1 || I.rec.n_expr_args != I.rec.n_in || (uint64_t)(I.rec.expr_offset + 4 * I.rec.expr_len) > h.n_expr)
                return fail(RDB_EINVAL, "instruction %u: bad scalar expression", i);
            if (I.rec.n_in + I.rec.expr_len > kMaxExprRegs) return fail(RDB_EUNSUPPORTED, "instruction %u: scalar expression too large", i);
            const int32_t *ops = expr + I.rec.expr_offset;
            for (int64_t k = 0; k < I.rec.expr_len; ++k) {
                int code = ops[4 * k], ra = ops[4 * k + 1], rb = ops[4 * k + 2], imm = ops[4 * k + 3];
                int reg = I.rec.n_in + (int)k;
                if (code < E_CONST || code > E_LAST) return fail(RDB_EUNSUPPORTED, "instruction %u: scalar function %d is not on the whitelist", i, code);
                if (code == E_CONST) { if (imm < 0 || (uint64_t)imm >= h.n_fconst) return fail(RDB_EINVAL, "instruction %u: bad literal", i); }
                else if (code == E_ARG) { if (imm < 0 || imm >= I.rec.n_in) return fail(RDB_EINVAL, "instruction %u: bad arg ref", i); }
                else {
                    if (ra < 0 || ra >= reg) return fail(RDB_EINVAL, "instruction %u: bad SSA ref", i);
                    bool binary = code == E_ADD || code == E_SUB || code == E_MUL || code == E_DIV || code == E_POW || code == E_MAX || code == E_MIN;
                    if (binary && (rb < 0 || rb >= reg)) return fail(RDB_EINVAL, "instruction %u: bad SSA ref", i);
                }
            }
            I.h_ops.assign(ops, ops + 4 * I.rec.expr_len);
            OK(dev_alloc(t, (void **)&I.d_ops, I.rec.expr_len * 16));
            CU(cudaMemcpyAsync(I.d_ops, ops, (size_t)I.rec.expr_len * 16, cudaMemcpyHostToDevice, S(t)));
            for (int a = 0; a < I.rec.n_in; ++a) {
                if (I.in[a].mode != OPM_PLAIN) return fail(RDB_EUNSUPPORTED, "instruction %u: elementwise op on an index-mapped operand", i);
                for (int d = 0; d < kMaxDims; ++d) {
                    int64_t od = d < I.in[a].rec.ndims ? I.in[a].rec.dims[d] : 1;
                    if (od != out.dims[d] && od != 1) return fail(RDB_EINVAL, "instruction %u: operand does not broadcast to the output", i);
                }
            }
            static const bool dedup = [] { const char *e = getenv("RDB_PARTIAL_DEDUP"); return !(e && e[0] == '0'); }();
            std::vector<std::string> sig;
            if (dedup && t->opts.fuse_elementwise) partial_signatures(ops, I.rec.expr_len, I.rec.n_in, sig);
            for (int a = 0; a < I.rec.n_in; ++a) {
                // results[I].partial[p] (broadcast.jl:155; elementwise.jl:239): only tracked arguments are ever read back
                if (!I.in[a].rec.tracked) continue;
                for (int p = 0; p < a && !sig.empty() && I.partial_alias[a] < 0; ++p)
                    if (I.in[p].rec.tracked && I.partial_alias[p] < 0 && sig[p] == sig[a] && sig[a] != "0" && sig[a].compare(0, 2, "op") != 0) I.partial_alias[a] = p;
                if (I.partial_alias[a] >= 0) I.partial[a] = I.partial[I.partial_alias[a]];
                else OK(dev_alloc(t, &I.partial[a], out.size * (int64_t)t->es));
            }
        } else if (is_getindex(op)) {
            if (I.in[0].mode != OPM_GATHER) return fail(RDB_EINVAL, "instruction %u: getindex without a map", i);
        } else if (op == OP_DOT) {
            if (out.size != 1 || I.in[0].size != I.in[1].size || I.in[0].mode != OPM_PLAIN || I.in[1].mode != OPM_PLAIN)
                return fail(RDB_EINVAL, "instruction %u: dot shape mismatch", i);
        } else if (op == OP_SUMDIMS) {
            Buf &src = t->bufs[I.in[0].rec.buffer];
            if (I.in[0].mode != OPM_PLAIN) return fail(RDB_EINVAL, "instruction %u: sum! on an index-mapped operand", i);
            for (int d = 0; d < kMaxDims; ++d)
                if (out.dims[d] != src.dims[d] && out.dims[d] != 1) return fail(RDB_EINVAL, "instruction %u: sum! shape mismatch", i);
        } else if (op == OP_SCALAR_BLOCK) {
            const OperandRec &o = I.rec.in[0];
            if (o.idx_kind != IDX_EXPLICIT || o.idx_offset < 0 || o.idx_len < 0 || (uint64_t)(o.idx_offset + o.idx_len) > h.n_idx)
                return fail(RDB_EINVAL, "instruction %u: scalar block payload out of range", i);
            std::vector<int64_t> sizes(h.n_buffers);
            for (uint32_t b = 0; b < h.n_buffers; ++b) sizes[b] = t->bufs[b].kind == BUF_CONST ? -1 : t->bufs[b].size;
            I.blk.reset(new ScalarBlock());
            std::string err = scalar_block_build(*I.blk, idx + o.idx_offset, o.idx_len, expr, h.n_expr, h.n_fconst, sizes, t->es, S(t));
            if (!err.empty())
                return fail(err.find("whitelist") != std::string::npos ? RDB_EUNSUPPORTED : RDB_EINVAL, "instruction %u: %s", i, err.c_str());
        }
    }
    // reduction scratch: kReduceBlocks doubles for sums + column-reduction partials
    t->scratch_bytes = std::max<int64_t>(kReduceBlocks * 8, 1 << 20);
    for (auto &I : t->instrs)
        if (is_elementwise(I.rec.opcode))
            for (int a = 0; a < I.rec.n_in; ++a)
                if (I.in[a].rec.tracked && I.in[a].size != t->bufs[I.rec.out].size)
                    t->scratch_bytes = std::max<int64_t>(t->scratch_bytes, I.in[a].size * 4096 * 8);
    OK(dev_alloc(t, &t->scratch, t->scratch_bytes));
    t->dz.assign(t->bufs.size(), 1);
    OK(ensure_batch(t, 1));
    CU(cudaStreamSynchronize(S(t)));
    return RDB_OK;
}

// ------------------------------------------------------------------------------------------ gradient
template <typename T>
static int seed_scalar_and_reverse(rdb_tape *t) {
    Buf &out = t->bufs[t->output];
    if (out.size != 1) return fail(RDB_EINVAL, "gradient! needs a scalar tape output");
    for (size_t b = 0; b < t->bufs.size(); ++b) t->dz[droot(t, (int)b)] = 1;   // unseed!(input); intermediates are zero by invariant
    if (t->ctx) t->ctx->gemm.uniform.clear();                                    // no deriv buffer is known to be a fill any more
    KL(t, launch_fill<T>(S(t), (T *)out.der, 1, (T)1));                     // seed!(output)
    t->dz[droot(t, t->output)] = 0;
    return reverse_pass<T>(t, 1);
}

struct NoCacheScope {
    GemmState &G;
    bool prev;
    NoCacheScope(GemmState &g, bool on) : G(g), prev(g.no_cache) { G.no_cache = on; }
    ~NoCacheScope() { G.no_cache = prev; }
};
static bool graph_eligible(rdb_tape *t) {
    if (!t->opts.use_graph || t->opts.profile || S(t) == nullptr || t->dmma_now) return false;     // the legacy default stream cannot be captured
    int64_t elems = 0;
    for (auto &b : t->bufs) elems += b.size;
    return elems <= (1 << 20);                                                       // launch-bound tapes only
}

// forward sweep + seeded reverse sweep as ONE graph launch; returns 1 if the graph ran, 0 if the caller must run eagerly
template <typename T>
static int graph_sweep(rdb_tape *t) {
    GemmState &G = t->ctx->gemm;
    if (t->graph_state == 2 && t->graph_gen != G.alloc_gen) {
        // the context's GEMM workspace moved since the capture (another tape grew it): the graph holds dead pointers
        cudaGraphExecDestroy(t->gexec);
        t->gexec = nullptr;
        t->graph_state = 1;
    }
    if (t->graph_state == 2) {
        if (G.guard_used != 0) CU(guard_fold(S(t), G));   // the graph writes certificate slots [0, n)
        CU(cudaGraphLaunch(t->gexec, S(t)));
        G.guard_used = t->graph_guard_n;
        t->launches += 1;
        return 1;
    }
    if (t->graph_state == 0) { t->graph_state = 1; return 0; }
    if (t->graph_state != 1) return 0;
    t->graph_state = -1;
    CU(guard_fold(S(t), G));                               // the captured kernels take certificate slots 0, 1, ...
    const uint64_t gen0 = G.alloc_gen;
    if (cudaStreamBeginCapture(S(t), cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return 0; }
    t->capturing = true;                                   // no side-stream work inside the capture
    int rc = forward_pass<T>(t);
    if (rc == RDB_OK) rc = seed_scalar_and_reverse<T>(t);
    t->capturing = false;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(S(t), &g);
    t->graph_guard_n = G.guard_used;
    G.guard_used = 0;                                      // nothing ran yet
    if (rc != RDB_OK || e != cudaSuccess || !g) { cudaGetLastError(); if (g) cudaGraphDestroy(g); return 0; }
    if (G.alloc_gen != gen0) { cudaGraphDestroy(g); t->graph_state = 1; return 0; }   // the workspace grew inside the capture: run eagerly, capture next time
    e = cudaGraphInstantiate(&t->gexec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { cudaGetLastError(); t->gexec = nullptr; return 0; }
    t->graph_state = 2;
    t->graph_gen = G.alloc_gen;
    CU(cudaGraphLaunch(t->gexec, S(t)));
    G.guard_used = t->graph_guard_n;
    t->launches += 1;
    return 1;
}

template <typename T>
static int gradient_impl(rdb_tape *t, void *const *result_ptrs, int on_device, void *value_out) {
    // graph tapes slice every operand into the context workspace on every replay (GemmState::no_cache), also in the eager warm-up
    // call, so that the workspace already has its final size when the capture starts
    NoCacheScope nc(t->ctx->gemm, graph_eligible(t));
    if (graph_eligible(t)) {
        int ran = graph_sweep<T>(t);
        if (ran < 0) return ran;
        if (ran == 1) {
            for (size_t i = 0; i < t->inputs.size(); ++i) {
                if (!result_ptrs || !result_ptrs[i]) continue;
                Buf &b = t->bufs[t->inputs[i]];
                OK(dmaterialize(t, t->inputs[i], 1, false));
                CU(cudaMemcpyAsync(result_ptrs[i], b.der, (size_t)b.size * t->es, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, S(t)));
            }
            if (value_out) CU(cudaMemcpyAsync(value_out, t->bufs[t->output].val, t->es, cudaMemcpyDeviceToHost, S(t)));
            if (t->async_call) { CU(cudaEventRecord(t->done_ev, S(t))); return RDB_OK; }
            CU(cudaStreamSynchronize(S(t)));
            return RDB_OK;
        }
    }
    OK(forward_pass<T>(t));
    if (value_out) CU(cudaMemcpyAsync(value_out, t->bufs[t->output].val, t->es, cudaMemcpyDeviceToHost, S(t)));
    // host results: plan the early copies (extract_result!, api/utils.jl:77-100, overlapped with the rest of the sweep)
    auto &E = t->early;
    E.enabled = false;
    if (!on_device && result_ptrs) {
        if (!t->ctx->copy_stream) CU(cudaStreamCreateWithFlags(&t->ctx->copy_stream, cudaStreamNonBlocking));
        E.remaining.assign(t->bufs.size(), 0);
        E.dst.assign(t->bufs.size(), nullptr);
        E.issued.assign(t->bufs.size(), 0);
        E.n_ev = 0;
        auto root = [&](int b) { while (t->bufs[b].alias_of >= 0) b = t->bufs[b].alias_of; return b; };
        for (auto &I : t->instrs) {
            for (int a = 0; a < I.rec.n_in; ++a)
                if (I.in[a].rec.tracked) E.remaining[root(I.in[a].rec.buffer)]++;
            if (I.blk)
                for (int k = 0; k < I.blk->n_ref; ++k)
                    if (I.blk->ref_has_leaf[k]) E.remaining[root(I.blk->refbufs[k])]++;
        }
        bool distinct = true;
        for (size_t i = 0; i < t->inputs.size(); ++i) {
            int r = root(t->inputs[i]);
            if (result_ptrs[i] && E.dst[r]) distinct = false;      // two result slots alias one buffer: keep the plain path
            if (result_ptrs[i]) E.dst[r] = result_ptrs[i];
        }
        // 0: never; 1 (default): always; 2: synchronous calls only (a pipelined caller overlaps whole evaluations instead)
        static const int early_mode = [] { const char *e = getenv("RDB_EARLY_D2H"); return e ? atoi(e) : 1; }();
        E.enabled = distinct && early_mode != 0 && !(early_mode == 2 && t->async_call);
    }
    int rc = seed_scalar_and_reverse<T>(t);
    E.enabled = false;
    OK(rc);
    for (size_t i = 0; i < t->inputs.size(); ++i) {                          // extract_result! for whatever was not copied early
        if (!result_ptrs || !result_ptrs[i]) continue;
        int r = t->inputs[i];
        while (t->bufs[r].alias_of >= 0) r = t->bufs[r].alias_of;
        if (!on_device && !E.issued.empty() && E.issued[r] && E.dst[r] == result_ptrs[i]) continue;
        Buf &b = t->bufs[t->inputs[i]];
        OK(dmaterialize(t, t->inputs[i], 1, false));
        CU(cudaMemcpyAsync(result_ptrs[i], b.der, (size_t)b.size * t->es, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, S(t)));
    }
    if (t->async_call) {                                     // enqueue only: the early copies join the main stream, one event closes the call
        if (t->ctx->copy_stream) {
            if (!t->join_ev) CU(cudaEventCreateWithFlags(&t->join_ev, cudaEventDisableTiming));
            CU(cudaEventRecord(t->join_ev, t->ctx->copy_stream));
            CU(cudaStreamWaitEvent(S(t), t->join_ev, 0));
        }
        CU(cudaEventRecord(t->done_ev, S(t)));
        return RDB_OK;
    }
    CU(cudaStreamSynchronize(S(t)));
    if (t->ctx->copy_stream) CU(cudaStreamSynchronize(t->ctx->copy_stream));
    return RDB_OK;
}

// ------------------------------------------------------------------------------------------ jacobian
template <typename T>
static int jacobian_impl(rdb_tape *t, int slot, int64_t rb, int64_t re, void *result, int64_t ld, int on_device, void *value_out) {
    if (slot < 0 || (size_t)slot >= t->inputs.size()) return fail(RDB_EINVAL, "bad input slot");
    Buf &out = t->bufs[t->output];
    Buf &x = t->bufs[t->inputs[slot]];
    const int64_t m = out.size, n = x.size;
    if (rb < 0 || re > m || rb > re) return fail(RDB_EINVAL, "row range [%lld,%lld) outside [0,%lld)", (long long)rb, (long long)re, (long long)m);
    const int64_t rows = re - rb;
    if (rows > 0 && ld < rows) return fail(RDB_EINVAL, "leading dimension smaller than the row block");
    OK(forward_pass<T>(t));
    if (value_out) CU(cudaMemcpyAsync(value_out, out.val, (size_t)m * t->es, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, S(t)));
    if (rows == 0) { CU(cudaStreamSynchronize(S(t))); return RDB_OK; }
    const int64_t R = auto_seed_block(t, rows, false);
    OK(ensure_batch(t, R));
    T *dres = (T *)result;
    int64_t dld = ld;
    if (!on_device) {
        OK(ensure_stage(t, rows * n * (int64_t)t->es));
        dres = (T *)t->stage;
        dld = rows;
    }
    for (int64_t i0 = rb; i0 < re; i0 += R) {
        const int64_t Rb = std::min(R, re - i0);
        for (size_t b = 0; b < t->bufs.size(); ++b) t->dz[droot(t, (int)b)] = 1;       // unseed!(input)
        if (t->ctx) t->ctx->gemm.uniform.clear();                                    // no deriv buffer is known to be a fill any more
        KL(t, launch_seed_onehot<T>(S(t), (T *)out.der, m, Rb, i0));                    // seed!(output, i) for Rb rows
        t->dz[droot(t, t->output)] = 0;
        OK(reverse_pass<T>(t, Rb));
        OK(dmaterialize(t, t->inputs[slot], Rb, false));
        StepTimer tm(t, "rows_out", 2.0 * n * Rb * t->es, 0);
        KL(t, launch_rows_out<T>(S(t), dres + (i0 - rb), dld, (const T *)x.der, n, Rb)); // result_matrix[i, :] = input_deriv
    }
    if (!on_device)
        CU(cudaMemcpy2DAsync(result, (size_t)ld * t->es, t->stage, (size_t)rows * t->es, (size_t)rows * t->es, (size_t)n,
                             cudaMemcpyDeviceToHost, S(t)));
    CU(cudaStreamSynchronize(S(t)));
    return RDB_OK;
}

// ------------------------------------------------------------------------------------------ hessian
static int ensure_tangents(rdb_tape *t, int64_t K) {
    if (K <= t->K_alloc) return RDB_OK;
    for (auto &b : t->bufs) {
        if (b.kind == BUF_CONST || b.alias_of >= 0) continue;
        dev_free(t, b.tv); dev_free(t, b.td);
        b.tv = b.td = nullptr;
        OK(dev_alloc(t, &b.tv, b.size * K * (int64_t)t->es));
        OK(dev_alloc(t, &b.td, b.size * K * (int64_t)t->es));
        CU(cudaMemsetAsync(b.td, 0, (size_t)(b.size * K) * t->es, S(t)));
    }
    for (auto &b : t->bufs)
        if (b.alias_of >= 0) { b.tv = t->bufs[b.alias_of].tv; b.td = t->bufs[b.alias_of].td; }
    t->K_alloc = K;
    return RDB_OK;
}

static int ensure_second_order(rdb_tape *t) {
    if (t->second_order) return RDB_OK;
    for (auto &I : t->instrs) {
        const int op = I.rec.opcode;
        if (!(op == OP_MUL || op == OP_ADD || op == OP_SUB || op == OP_NEG || op == OP_SUM || op == OP_MEAN || op == OP_DOT || op == OP_SUMDIMS || is_getindex(op) || is_elementwise(op)))
            return fail(RDB_EUNSUPPORTED, "hessian!: opcode %d has no forward-over-reverse rule yet", op);
        if (op == OP_DOT) {
            for (int a = 0; a < 2; ++a)
                if (I.in[a].mode != OPM_PLAIN) return fail(RDB_EUNSUPPORTED, "hessian!: index-mapped `dot` operand");
            continue;
        }
        if (op == OP_MUL) {
            if (I.in[0].rec.tracked && I.in[1].rec.tracked && (I.in[0].mode != OPM_PLAIN || I.in[1].mode != OPM_PLAIN))
                return fail(RDB_EUNSUPPORTED, "hessian!: `*` of two tracked operands needs both of them untransposed");
            for (int a = 0; a < 2; ++a)
                if (I.in[a].mode == OPM_GATHER) return fail(RDB_EUNSUPPORTED, "hessian!: gathered `*` operand");
            continue;
        }
        if (!is_elementwise(op)) continue;
        Buf &out = t->bufs[I.rec.out];
        const int N = I.rec.n_in;
        if (N > 4) return fail(RDB_EUNSUPPORTED, "hessian!: elementwise closures of more than four arguments have no second-order kernel yet");
        for (int a = 0; a < N; ++a) {
            if (!I.partial[a]) OK(dev_alloc(t, &I.partial[a], out.size * (int64_t)t->es));
        }
        for (int h = 0; h < N * (N + 1) / 2; ++h) OK(dev_alloc(t, &I.second[h], out.size * (int64_t)t->es));
    }
    t->second_order = true;
    return RDB_OK;
}

// per-argument strides of an elementwise instruction against its output (0 along clamped dimensions, propagation.jl:25-27)
static void bcast_strides(const InstrPlan &I, const Buf &out, int64_t (*strides)[kMaxDims], int64_t *sizes) {
    for (int a = 0; a < I.rec.n_in; ++a) {
        const OperandPlan &o = I.in[a];
        int64_t st = 1;
        for (int d = 0; d < kMaxDims; ++d) {
            const int64_t od = d < o.rec.ndims ? o.rec.dims[d] : 1;
            strides[a][d] = (od == 1 && out.dims[d] != 1) ? 0 : st;
            st *= od;
        }
        sizes[a] = o.size;
    }
}

static int tri_index(int p, int q, int N) {
    if (p > q) std::swap(p, q);
    return p * N - p * (p - 1) / 2 + (q - p);
}

// forward-over-reverse: value tangents forward, (adjoint, adjoint tangent) backward, K directions at a time.
//
// Traffic is everything here (the result alone is n*n*E bytes), so the sweep is planned:
//  * value tangents are computed only for buffers whose tangent some nonlinear instruction will read (liveness);
//  * the tangent of `getindex(input)` is written directly as a one-hot selector of the index map — the n x K seed block of
//    the input itself is only materialised if a non-getindex instruction consumes it;
//  * adjoint tangents carry a "known zero" flag per buffer and sweep: a zero tangent is never read, the first accumulation
//    into a buffer overwrites instead of memset + read-modify-write (the reference's unseed!/increment pattern with the zero
//    state tracked on the host);
//  * the adjoint tangent of the input IS the result block H[:, c0:c0+K] (column-major, ld = n): it is accumulated in place.
template <typename T>
static int hessian_impl(rdb_tape *t, int64_t cb, int64_t ce, void *result, int64_t ld, int on_device, void *grad_out, void *value_out) {
    if (t->inputs.size() != 1) return fail(RDB_EINVAL, "Taking the Hessian of a function with multiple arguments is not yet supported");
    Buf &out = t->bufs[t->output];
    const int xid = t->inputs[0];
    Buf &x = t->bufs[xid];
    if (out.size != 1) return fail(RDB_EINVAL, "hessian! needs a scalar tape output");
    const int64_t n = x.size;
    if (cb < 0 || ce > n || cb > ce) return fail(RDB_EINVAL, "column range outside [0,n)");
    const int64_t cols = ce - cb;
    if (cols > 0 && ld < n) return fail(RDB_EINVAL, "leading dimension smaller than n");
    for (auto &I : t->instrs)
        if (I.blk) return fail(RDB_EUNSUPPORTED, "hessian!: the forward-over-reverse sweep does not cover scalar instructions; "
                                                 "record the reference's reverse-over-reverse tape instead (HessianTape(..., mode=\"reverse\"))");
    OK(ensure_second_order(t));
    // this sweep keeps the reference's physical unseed!/increment pattern for the (size-n) first-order adjoints
    for (size_t b = 0; b < t->bufs.size(); ++b) OK(dmaterialize(t, (int)b, 1, false));
    OK(forward_pass<T>(t, 2));
    if (value_out) CU(cudaMemcpyAsync(value_out, out.val, t->es, cudaMemcpyDeviceToHost, S(t)));
    if (cols == 0 && !grad_out) { CU(cudaStreamSynchronize(S(t))); return RDB_OK; }
    const int64_t K = std::max<int64_t>(1, auto_seed_block(t, std::max<int64_t>(cols, 1), true));
    OK(ensure_tangents(t, K));
    T *dres = (T *)result;
    int64_t dld = ld;
    if (!on_device && cols > 0) {
        OK(ensure_stage(t, n * cols * (int64_t)t->es));
        dres = (T *)t->stage;
        dld = n;
    }
    const size_t nb = t->bufs.size();
    auto root = [&](int b) { while (t->bufs[b].alias_of >= 0) b = t->bufs[b].alias_of; return b; };

    // ---- static plan (backward over the tape, mirrors the sweep): which adjoint tangents can be non-zero, which elementwise
    // instructions can take the sparse path (all arguments are getindex views of the input, incoming adjoint tangent zero),
    // and which value tangents somebody will actually read
    const size_t ni = t->instrs.size();
    std::vector<int> producer(nb, -1);
    for (size_t ii = 0; ii < ni; ++ii) producer[root(t->instrs[ii].rec.out)] = (int)ii;
    std::vector<char> td_nz(nb, 0), sparse_ok(ni, 0), need_tv(nb, 0);
    for (size_t ii = ni; ii-- > 0;) {
        InstrPlan &I = t->instrs[ii];
        const bool dtz = !td_nz[root(I.rec.out)];
        if (is_elementwise(I.rec.opcode)) {
            bool views = dtz;
            for (int a = 0; a < I.rec.n_in && views; ++a) {
                int pr = producer[root(I.in[a].rec.buffer)];
                views = I.in[a].rec.tracked && pr >= 0 && is_getindex(t->instrs[pr].rec.opcode) &&
                        root(t->instrs[pr].in[0].rec.buffer) == root(xid) && I.in[a].size == t->bufs[I.rec.out].size;
            }
            sparse_ok[ii] = views;
            if (views) td_nz[root(xid)] = 1;
            else for (int a = 0; a < I.rec.n_in; ++a) if (I.in[a].rec.tracked) td_nz[root(I.in[a].rec.buffer)] = 1;
        } else if ((I.rec.opcode == OP_MUL || I.rec.opcode == OP_DOT) && I.in[0].rec.tracked && I.in[1].rec.tracked) {
            // bilinear: the cross terms delta * tv(b)' and tv(a)' * delta exist even when no adjoint tangent flows in
            for (int a = 0; a < 2; ++a) td_nz[root(I.in[a].rec.buffer)] = 1;
        } else if (!dtz) {
            for (int a = 0; a < I.rec.n_in; ++a) if (I.in[a].rec.tracked) td_nz[root(I.in[a].rec.buffer)] = 1;
        }
    }
    for (size_t ii = ni; ii-- > 0;) {
        InstrPlan &I = t->instrs[ii];
        const bool bilinear = (I.rec.opcode == OP_MUL || I.rec.opcode == OP_DOT) && I.in[0].rec.tracked && I.in[1].rec.tracked;
        const bool nonlinear = (is_elementwise(I.rec.opcode) && !sparse_ok[ii]) || bilinear;
        for (int a = 0; a < I.rec.n_in; ++a) {
            if (!I.in[a].rec.tracked) continue;
            if (nonlinear || (!is_elementwise(I.rec.opcode) && need_tv[root(I.rec.out)])) need_tv[root(I.in[a].rec.buffer)] = 1;
        }
    }
    // the input's own n x K seed block is needed only by consumers other than getindex
    bool need_x_tv = false;
    for (size_t ii = 0; ii < ni; ++ii) {
        InstrPlan &I = t->instrs[ii];
        for (int a = 0; a < I.rec.n_in; ++a)
            if (root(I.in[a].rec.buffer) == root(xid) && !is_getindex(I.rec.opcode) &&
                ((is_elementwise(I.rec.opcode) && !sparse_ok[ii]) || (!is_elementwise(I.rec.opcode) && need_tv[root(I.rec.out)]) ||
                 ((I.rec.opcode == OP_MUL || I.rec.opcode == OP_DOT) && I.in[0].rec.tracked && I.in[1].rec.tracked)))
                need_x_tv = true;
    }

    // first-order adjoints: with no gradient requested, only the adjoints that some second-order term multiplies are needed -
    // deriv(output of a nonlinear instruction) and everything between it and the tape output.  need_der[b]: the sub-DAG that
    // produces b contains a nonlinear instruction.
    const bool lean = !grad_out && on_device && cols > 0;
    std::vector<char> need_der(nb, 1);
    if (lean) {
        std::fill(need_der.begin(), need_der.end(), 0);
        for (size_t ii = 0; ii < ni; ++ii) {
            InstrPlan &I = t->instrs[ii];
            bool nl = is_elementwise(I.rec.opcode) || ((I.rec.opcode == OP_MUL || I.rec.opcode == OP_DOT) && I.in[0].rec.tracked && I.in[1].rec.tracked);
            for (int a = 0; a < I.rec.n_in && !nl; ++a) nl = need_der[root(I.in[a].rec.buffer)] != 0;
            if (nl) need_der[root(I.rec.out)] = 1;
        }
        need_der[root(t->output)] = 1;                       // seeded below, so it must be unseeded again
    }
    if (!t->capturing) {                                     // (nothing is created inside a stream capture)
        rdb_ctx *c = t->ctx;
        if (!c->side_stream) CU(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
        if (!t->fork_ev) CU(cudaEventCreateWithFlags(&t->fork_ev, cudaEventDisableTiming));
        if (!t->join_ev) CU(cudaEventCreateWithFlags(&t->join_ev, cudaEventDisableTiming));
    }
    auto want_der = [&](int b) { return need_der[root(b)] != 0; };

    void *const x_td_own = x.td;
    bool first = true;
    for (int64_t c0 = cb; c0 < ce || first; c0 += K) {
        const int64_t Kb = cols > 0 ? std::min(K, ce - c0) : 1;
        // accumulate the input's adjoint tangent straight into the result block when its leading dimension is n
        const bool in_place = cols > 0 && dld == n;
        if (in_place) x.td = dres + (c0 - cb) * dld;
        for (auto &b : t->bufs) if (b.alias_of >= 0) b.td = t->bufs[b.alias_of].td;
        // The zero-fill of the block (n x Kb: the bulk of the call's HBM traffic when the second-order terms are sparse) does not
        // depend on the forward sweep or on the small first-order kernels: fork it onto the side stream, join before the first
        // scatter into the block.  Whether the first writer wants zeros is learnt on the first (eager) sweep.
        bool prefilled = false;
        if (t->hess_prefill == 1 && cols > 0) {
            rdb_ctx *c = t->ctx;
            CU(cudaEventRecord(t->fork_ev, S(t)));
            CU(cudaStreamWaitEvent(c->side_stream, t->fork_ev, 0));
            CU(cudaMemsetAsync(x.td, 0, (size_t)(n * Kb) * t->es, c->side_stream));
            CU(cudaEventRecord(t->join_ev, c->side_stream));
            prefilled = true;
        }
        // make the block's storage really zero before a scatter / atomic accumulation: join the forked fill, or fill now
        auto zero_x_td = [&]() -> int {
            if (prefilled) { CU(cudaStreamWaitEvent(S(t), t->join_ev, 0)); prefilled = false; }
            else {
                CU(cudaMemsetAsync(x.td, 0, (size_t)(n * Kb) * t->es, S(t)));
                if (t->hess_prefill == 0) t->hess_prefill = 1;
            }
            return RDB_OK;
        };
        // a full-coverage first write into the block: nothing to fill, and no point forking a fill next time
        auto overwrite_x_td = [&]() -> int {
            if (prefilled) { CU(cudaStreamWaitEvent(S(t), t->join_ev, 0)); prefilled = false; }
            if (t->hess_prefill == 0) t->hess_prefill = -1;
            return RDB_OK;
        };

        // ---- tangent-forward sweep
        if (need_x_tv) KL(t, launch_seed_tangent<T>(S(t), (T *)x.tv, n, Kb, c0));
        for (auto &I : t->instrs) {
            Buf &o = t->bufs[I.rec.out];
            if (!need_tv[root(I.rec.out)]) continue;
            switch (I.rec.opcode) {
                case OP_GETINDEX: case OP_GETINDEX_GENERIC: {
                    Buf &src = t->bufs[I.in[0].rec.buffer];
                    StepTimer tm(t, "tangent_getindex", 1.0 * o.size * Kb * t->es, 0);
                    if (root(I.in[0].rec.buffer) == root(xid))
                        KL(t, launch_seed_gather<T>(S(t), (T *)o.tv, I.in[0].d_map, o.size, Kb, c0));
                    else
                        KL(t, launch_gather_cols<T>(S(t), (T *)o.tv, (const T *)src.tv, src.size, I.in[0].d_map, o.size, Kb));
                } break;
                case OP_ADD: case OP_SUB: {
                    T sign = I.rec.opcode == OP_ADD ? (T)1 : (T)-1;
                    Buf &a = t->bufs[I.in[0].rec.buffer], &b = t->bufs[I.in[1].rec.buffer];
                    if (I.in[0].rec.tracked) KL(t, launch_axpy<T>(S(t), (T *)o.tv, (const T *)a.tv, (T)1, o.size * Kb, true));
                    else CU(cudaMemsetAsync(o.tv, 0, (size_t)(o.size * Kb) * t->es, S(t)));
                    if (I.in[1].rec.tracked) KL(t, launch_axpy<T>(S(t), (T *)o.tv, (const T *)b.tv, sign, o.size * Kb, false));
                } break;
                case OP_NEG:
                    KL(t, launch_axpy<T>(S(t), (T *)o.tv, (const T *)t->bufs[I.in[0].rec.buffer].tv, (T)-1, o.size * Kb, true));
                    break;
                case OP_SUM: case OP_MEAN: {
                    T scale = I.rec.opcode == OP_MEAN ? (T)1 / (T)I.in[0].size : (T)1;
                    KL(t, launch_colsum<T>(S(t), (T *)o.tv, (const T *)t->bufs[I.in[0].rec.buffer].tv, I.in[0].size, Kb, scale));
                } break;
                case OP_SUMDIMS: {
                    // linear: the same reduction of every direction's tangent (reductions.jl:39-46); the three shapes the forward sweep lowers
                    Buf &src = t->bufs[I.in[0].rec.buffer];
                    const int64_t d0 = src.dims[0], d1 = src.size / src.dims[0];
                    if (o.size == 1) KL(t, launch_colsum<T>(S(t), (T *)o.tv, (const T *)src.tv, src.size, Kb, (T)1));
                    else if (o.dims[0] == 1 && o.size == d1) KL(t, launch_colsum<T>(S(t), (T *)o.tv, (const T *)src.tv, d0, d1 * Kb, (T)1));
                    else if (o.dims[0] == d0 && o.size == d0) {
                        CU(cudaMemsetAsync(o.tv, 0, (size_t)(o.size * Kb) * t->es, S(t)));
                        for (int64_t k = 0; k < Kb; ++k)
                            KL(t, launch_mul_acc_colvec<T>(S(t), (T *)o.tv + k * o.size, (const T *)src.tv + k * src.size, (const T *)nullptr, d0, d1, 1, (T *)t->scratch, t->scratch_bytes / 8));
                    } else return fail(RDB_EUNSUPPORTED, "sum over this combination of dimensions is not lowered");
                } break;
                case OP_DOT: {
                    // d(x . y) = dx . y + x . dy (reductions.jl:91-104): per direction k a dot product, all directions as one
                    // (1 x n) * (n x Kb) product per tracked operand
                    Buf &a = t->bufs[I.in[0].rec.buffer], &b = t->bufs[I.in[1].rec.buffer];
                    const int64_t nn = a.size;
                    bool first_term = true;
                    if (I.in[0].rec.tracked) { KL(t, gemm_t<T>(S(t), true, false, 1, Kb, nn, (T)1, (const T *)b.val, nn, (const T *)a.tv, nn, (T)0, (T *)o.tv, 1)); first_term = false; }
                    if (I.in[1].rec.tracked) KL(t, gemm_t<T>(S(t), true, false, 1, Kb, nn, (T)1, (const T *)a.val, nn, (const T *)b.tv, nn, first_term ? (T)0 : (T)1, (T *)o.tv, 1));
                } break;
                case OP_MUL: {
                    // exactly one tracked operand (checked): tangent is linear in it
                    OperandPlan &oa = I.in[0], &ob = I.in[1];
                    int64_t M = oa.rows, Kd = oa.cols, N = ob.cols;
                    if (oa.rec.tracked && ob.rec.tracked) {
                        // d(a b) = da b + a db: a * [tv(b) for every direction] as one GEMM, then tv(a)_k * b per direction
                        Buf &a = t->bufs[oa.rec.buffer], &b = t->bufs[ob.rec.buffer];
                        KL(t, gemm_t<T>(S(t), false, false, M, N * Kb, Kd, (T)1, (const T *)a.val, M, (const T *)b.tv, Kd, (T)0, (T *)o.tv, M));
                        for (int64_t k = 0; k < Kb; ++k)
                            KL(t, gemm_t<T>(S(t), false, false, M, N, Kd, (T)1, (const T *)a.tv + k * a.size, M, (const T *)b.val, Kd, (T)1, (T *)o.tv + k * o.size, M));
                    } else if (ob.rec.tracked) {
                        Buf &a = t->bufs[oa.rec.buffer], &b = t->bufs[ob.rec.buffer];
                        const bool st = oa.mode == OPM_FOLDED_T, sb = ob.mode == OPM_FOLDED_T;
                        if (!sb) KL(t, gemm_t<T>(S(t), st, false, M, N * Kb, Kd, (T)1, (const T *)a.val, st ? oa.cols : oa.rows, (const T *)b.tv, Kd, (T)0, (T *)o.tv, M));
                        else                                  // b is the transposed view of a tracked N x Kd buffer: its tangents are stored untransposed, one product per direction
                            for (int64_t k = 0; k < Kb; ++k)
                                KL(t, gemm_t<T>(S(t), st, true, M, N, Kd, (T)1, (const T *)a.val, st ? oa.cols : oa.rows, (const T *)b.tv + k * b.size, ob.cols, (T)0, (T *)o.tv + k * o.size, M));
                    } else {
                        Buf &a = t->bufs[oa.rec.buffer], &b = t->bufs[ob.rec.buffer];
                        const bool sa = oa.mode == OPM_FOLDED_T, st = ob.mode == OPM_FOLDED_T;
                        for (int64_t k = 0; k < Kb; ++k)
                            KL(t, gemm_t<T>(S(t), sa, st, M, N, Kd, (T)1, (const T *)a.tv + k * a.size, sa ? oa.cols : oa.rows, (const T *)b.val, st ? ob.cols : ob.rows, (T)0, (T *)o.tv + k * o.size, M));
                    }
                } break;
                default: {   // elementwise closures
                    const T *part[kMaxArgs], *tvs[kMaxArgs];
                    for (int a = 0; a < I.rec.n_in; ++a) {
                        part[a] = (const T *)I.partial[a];
                        tvs[a] = I.in[a].rec.tracked ? (const T *)t->bufs[I.in[a].rec.buffer].tv : nullptr;   // untracked: no tangent
                    }
                    StepTimer tm(t, "tangent_forward", (double)(I.rec.n_in + 1) * o.size * Kb * t->es, 0);
                    bool same_shape = true;
                    for (int a = 0; a < I.rec.n_in; ++a) same_shape &= I.in[a].size == o.size;
                    if (same_shape) KL(t, launch_tangent_forward<T>(S(t), (T *)o.tv, I.rec.n_in, part, tvs, o.size, Kb));
                    else {
                        int64_t strides[kMaxArgs][kMaxDims], sizes[kMaxArgs];
                        bcast_strides(I, o, strides, sizes);
                        KL(t, launch_tangent_forward_g<T>(S(t), (T *)o.tv, I.rec.n_in, part, tvs, strides, sizes, o.dims, o.size, Kb));
                    }
                }
            }
        }
        // ---- reverse sweep carrying (adjoint, adjoint tangent); first-order adjoints are recomputed per chunk (size n, cheap)
        std::vector<char> td_zero(nb, 1);               // per ROOT buffer: adjoint tangent is (logically) zero
        for (int b : t->inputs) if (want_der(b)) OK(unseed<T>(t, t->bufs[b], 1));
        KL(t, launch_fill<T>(S(t), (T *)out.der, 1, (T)1));
        for (size_t ii = t->instrs.size(); ii-- > 0;) {
            InstrPlan &I = t->instrs[ii];
            Buf &o = t->bufs[I.rec.out];
            const T *delta = (const T *)o.der;
            const bool dt_zero = td_zero[root(I.rec.out)];
            const T *dt = dt_zero ? nullptr : (const T *)o.td;
            switch (I.rec.opcode) {
                case OP_GETINDEX: case OP_GETINDEX_GENERIC: {
                    const int rb = root(I.in[0].rec.buffer);
                    Buf &b = t->bufs[I.in[0].rec.buffer];
                    if (want_der(I.in[0].rec.buffer)) KL(t, launch_scatter_add<T>(S(t), (T *)b.der, b.size, I.in[0].d_map, delta, o.size, 1));
                    if (!dt_zero) {
                        if (td_zero[rb]) {
                            if (rb == root(xid)) OK(zero_x_td()); else CU(cudaMemsetAsync(b.td, 0, (size_t)(b.size * Kb) * t->es, S(t)));
                            td_zero[rb] = 0;
                        }
                        StepTimer tm(t, "tangent_getindex_reverse", 3.0 * o.size * Kb * t->es, 0);
                        KL(t, launch_scatter_add<T>(S(t), (T *)b.td, b.size, I.in[0].d_map, dt, o.size, Kb));
                    }
                } break;
                case OP_ADD: case OP_SUB:
                    for (int a = 0; a < 2; ++a) {
                        if (!I.in[a].rec.tracked) continue;
                        const int rb = root(I.in[a].rec.buffer);
                        Buf &b = t->bufs[I.in[a].rec.buffer];
                        T sign = (a == 1 && I.rec.opcode == OP_SUB) ? (T)-1 : (T)1;
                        if (want_der(I.in[a].rec.buffer)) KL(t, launch_axpy<T>(S(t), (T *)b.der, delta, sign, o.size, false));
                        if (!dt_zero) {
                            if (td_zero[rb] && rb == root(xid)) OK(overwrite_x_td());
                            KL(t, launch_axpy<T>(S(t), (T *)b.td, dt, sign, o.size * Kb, td_zero[rb] != 0)); td_zero[rb] = 0;
                        }
                    }
                    break;
                case OP_NEG: {
                    const int rb = root(I.in[0].rec.buffer);
                    Buf &b = t->bufs[I.in[0].rec.buffer];
                    if (want_der(I.in[0].rec.buffer)) KL(t, launch_axpy<T>(S(t), (T *)b.der, delta, (T)-1, o.size, false));
                    if (!dt_zero) {
                        if (td_zero[rb] && rb == root(xid)) OK(overwrite_x_td());
                        KL(t, launch_axpy<T>(S(t), (T *)b.td, dt, (T)-1, o.size * Kb, td_zero[rb] != 0)); td_zero[rb] = 0;
                    }
                } break;
                case OP_SUM: case OP_MEAN: {
                    const int rb = root(I.in[0].rec.buffer);
                    Buf &b = t->bufs[I.in[0].rec.buffer];
                    T scale = I.rec.opcode == OP_MEAN ? (T)1 / (T)I.in[0].size : (T)1;
                    if (want_der(I.in[0].rec.buffer)) KL(t, launch_acc_scalar<T>(S(t), (T *)b.der, b.size, 1, delta, scale, false));
                    if (!dt_zero) {
                        if (td_zero[rb] && rb == root(xid)) OK(overwrite_x_td());
                        KL(t, launch_acc_scalar<T>(S(t), (T *)b.td, b.size, Kb, dt, scale, td_zero[rb] != 0)); td_zero[rb] = 0;
                    }
                } break;
                case OP_SUMDIMS: {
                    if (!I.in[0].rec.tracked) break;
                    const int rb = root(I.in[0].rec.buffer);
                    Buf &b = t->bufs[I.in[0].rec.buffer];
                    int64_t dstride[kMaxDims], st = 1;
                    for (int d = 0; d < kMaxDims; ++d) { dstride[d] = (o.dims[d] == 1 && b.dims[d] != 1) ? 0 : st; st *= o.dims[d]; }
                    if (want_der(I.in[0].rec.buffer)) KL(t, launch_bcast_acc<T>(S(t), (T *)b.der, b.dims, dstride, delta, b.size, 1, o.size, false));
                    if (!dt_zero) {
                        if (td_zero[rb] && rb == root(xid)) OK(overwrite_x_td());
                        KL(t, launch_bcast_acc<T>(S(t), (T *)b.td, b.dims, dstride, dt, b.size, Kb, o.size, td_zero[rb] != 0)); td_zero[rb] = 0;
                    }
                } break;
                case OP_DOT:
                    // out = x . y, delta and dt_k scalars.  First order: deriv(x) += delta y.  Second order, per direction k:
                    // td(x)_k += dt_k y + delta tv(y)_k (the second term only when y is tracked too), and the same with x and y exchanged
                    for (int a = 0; a < 2; ++a) {
                        if (!I.in[a].rec.tracked) continue;
                        const int rb = root(I.in[a].rec.buffer);
                        Buf &b = t->bufs[I.in[a].rec.buffer], &other = t->bufs[I.in[1 - a].rec.buffer];
                        if (want_der(I.in[a].rec.buffer)) KL(t, launch_scal_acc<T>(S(t), (T *)b.der, (const T *)other.val, delta, b.size, 1, false));
                        const bool cross = I.in[1 - a].rec.tracked;
                        if (dt_zero && !cross) continue;
                        if (td_zero[rb] && rb == root(xid)) OK(overwrite_x_td());
                        if (!dt_zero) { KL(t, launch_scal_acc<T>(S(t), (T *)b.td, (const T *)other.val, dt, b.size, Kb, td_zero[rb] != 0)); td_zero[rb] = 0; }
                        if (cross) { KL(t, launch_scal_acc<T>(S(t), (T *)b.td, (const T *)other.tv, delta, b.size * Kb, 1, td_zero[rb] != 0)); td_zero[rb] = 0; }
                    }
                    break;
                case OP_MUL: {
                    OperandPlan &oa = I.in[0], &ob = I.in[1];
                    int64_t M = oa.rows, Kd = oa.cols, N = ob.cols;
                    if (oa.rec.tracked && ob.rec.tracked) {
                        // out = a b with both tracked (untransposed).  First order: deriv(a) += delta b', deriv(b) += a' delta.
                        // Second order, per direction k:   td(a)_k += dt_k b' + delta tv(b)_k'      td(b)_k += a' dt_k + tv(a)_k' delta
                        Buf &a = t->bufs[oa.rec.buffer], &b = t->bufs[ob.rec.buffer];
                        const int ra = root(oa.rec.buffer), rb = root(ob.rec.buffer);
                        if (want_der(oa.rec.buffer)) KL(t, gemm_t<T>(S(t), false, true, M, Kd, N, (T)1, delta, M, (const T *)b.val, Kd, (T)1, (T *)a.der, M));
                        if (want_der(ob.rec.buffer)) KL(t, gemm_t<T>(S(t), true, false, Kd, N, M, (T)1, (const T *)a.val, M, delta, M, (T)1, (T *)b.der, Kd));
                        // td(b): one GEMM over all directions for a' dt, then the cross term per direction
                        if (td_zero[rb] && rb == root(xid)) OK(overwrite_x_td());
                        bool bz = td_zero[rb] != 0;
                        if (!dt_zero) { KL(t, gemm_t<T>(S(t), true, false, Kd, N * Kb, M, (T)1, (const T *)a.val, M, dt, M, bz ? (T)0 : (T)1, (T *)b.td, Kd)); bz = false; }
                        for (int64_t k = 0; k < Kb; ++k)
                            KL(t, gemm_t<T>(S(t), true, false, Kd, N, M, (T)1, (const T *)a.tv + k * a.size, M, delta, M, bz ? (T)0 : (T)1, (T *)b.td + k * b.size, Kd));
                        td_zero[rb] = 0;
                        // td(a) (a and b may be the same buffer: x * x accumulates both halves into one td)
                        if (td_zero[ra] && ra == root(xid)) OK(overwrite_x_td());
                        const bool az = td_zero[ra] != 0;
                        for (int64_t k = 0; k < Kb; ++k) {
                            bool z = az;
                            if (!dt_zero) { KL(t, gemm_t<T>(S(t), false, true, M, Kd, N, (T)1, dt + k * o.size, M, (const T *)b.val, Kd, z ? (T)0 : (T)1, (T *)a.td + k * a.size, M)); z = false; }
                            KL(t, gemm_t<T>(S(t), false, true, M, Kd, N, (T)1, delta, M, (const T *)b.tv + k * b.size, Kd, z ? (T)0 : (T)1, (T *)a.td + k * a.size, M));
                        }
                        td_zero[ra] = 0;
                    } else if (ob.rec.tracked && ob.mode == OPM_FOLDED_T) {
                        // b = B' with B (N x Kd) tracked: the adjoints are computed directly transposed (arithmetic.jl:272-285 on the
                        // Adjoint): deriv(B) += delta' op(a), td(B)_k += dt_k' op(a)
                        const int rb = root(ob.rec.buffer);
                        Buf &a = t->bufs[oa.rec.buffer], &b = t->bufs[ob.rec.buffer];
                        const bool st = oa.mode == OPM_FOLDED_T;
                        if (want_der(ob.rec.buffer)) KL(t, gemm_t<T>(S(t), true, st, N, Kd, M, (T)1, delta, M, (const T *)a.val, st ? oa.cols : oa.rows, (T)1, (T *)b.der, N));
                        if (!dt_zero) {
                            if (td_zero[rb] && rb == root(xid)) OK(overwrite_x_td());
                            for (int64_t k = 0; k < Kb; ++k)
                                KL(t, gemm_t<T>(S(t), true, st, N, Kd, M, (T)1, dt + k * o.size, M, (const T *)a.val, st ? oa.cols : oa.rows, td_zero[rb] ? (T)0 : (T)1, (T *)b.td + k * b.size, N));
                            td_zero[rb] = 0;
                        }
                    } else if (ob.rec.tracked) {
                        const int rb = root(ob.rec.buffer);
                        Buf &a = t->bufs[oa.rec.buffer], &b = t->bufs[ob.rec.buffer];
                        bool st = oa.mode == OPM_FOLDED_T;
                        if (want_der(ob.rec.buffer)) KL(t, gemm_t<T>(S(t), !st, false, Kd, N, M, (T)1, (const T *)a.val, st ? oa.cols : oa.rows, delta, M, (T)1, (T *)b.der, Kd));
                        if (!dt_zero) {
                            if (td_zero[rb] && rb == root(xid)) OK(overwrite_x_td());
                            KL(t, gemm_t<T>(S(t), !st, false, Kd, N * Kb, M, (T)1, (const T *)a.val, st ? oa.cols : oa.rows, dt, M, td_zero[rb] ? (T)0 : (T)1, (T *)b.td, Kd));
                            td_zero[rb] = 0;
                        }
                    } else if (oa.mode == OPM_FOLDED_T) {
                        // a = A' with A (Kd x M) tracked: deriv(A) += op(b) delta', td(A)_k += op(b) dt_k'
                        const int rb = root(oa.rec.buffer);
                        Buf &a = t->bufs[oa.rec.buffer], &b = t->bufs[ob.rec.buffer];
                        const bool st = ob.mode == OPM_FOLDED_T;
                        if (want_der(oa.rec.buffer)) KL(t, gemm_t<T>(S(t), st, true, Kd, M, N, (T)1, (const T *)b.val, st ? ob.cols : ob.rows, delta, M, (T)1, (T *)a.der, Kd));
                        if (!dt_zero) {
                            if (td_zero[rb] && rb == root(xid)) OK(overwrite_x_td());
                            for (int64_t k = 0; k < Kb; ++k)
                                KL(t, gemm_t<T>(S(t), st, true, Kd, M, N, (T)1, (const T *)b.val, st ? ob.cols : ob.rows, dt + k * o.size, M, td_zero[rb] ? (T)0 : (T)1, (T *)a.td + k * a.size, Kd));
                            td_zero[rb] = 0;
                        }
                    } else {
                        const int rb = root(oa.rec.buffer);
                        Buf &a = t->bufs[oa.rec.buffer], &b = t->bufs[ob.rec.buffer];
                        bool st = ob.mode == OPM_FOLDED_T;
                        if (want_der(oa.rec.buffer)) KL(t, gemm_t<T>(S(t), false, !st, M, Kd, N, (T)1, delta, M, (const T *)b.val, st ? ob.cols : ob.rows, (T)1, (T *)a.der, M));
                        if (!dt_zero) {
                            if (td_zero[rb] && rb == root(xid)) OK(overwrite_x_td());
                            for (int64_t k = 0; k < Kb; ++k)
                                KL(t, gemm_t<T>(S(t), false, !st, M, Kd, N, (T)1, dt + k * o.size, M, (const T *)b.val, st ? ob.cols : ob.rows, td_zero[rb] ? (T)0 : (T)1, (T *)a.td + k * a.size, M));
                            td_zero[rb] = 0;
                        }
                    }
                } break;
                default: {
                    const int N = I.rec.n_in;
                    if (sparse_ok[ii]) {
                        // every argument is a getindex view of the seeded input and no adjoint tangent flows in
                        const int rx = root(xid);
                        if (td_zero[rx]) { OK(zero_x_td()); td_zero[rx] = 0; }
                        const int64_t *maps[kMaxArgs];
                        const T *sec[kMaxArgs * kMaxArgs];
                        for (int p = 0; p < N; ++p) {
                            maps[p] = t->instrs[producer[root(I.in[p].rec.buffer)]].in[0].d_map;
                            for (int q = 0; q < N; ++q) sec[p * N + q] = (const T *)I.second[tri_index(p, q, N)];
                        }
                        StepTimer tm(t, "tangent_reverse_sparse", (double)n * Kb * t->es, 0);
                        KL(t, launch_tangent_reverse_sparse<T>(S(t), (T *)x.td, n, Kb, c0, delta, N, maps, sec, o.size));
                        for (int p = 0; p < N; ++p)
                            if (want_der(I.in[p].rec.buffer))
                                KL(t, launch_mul_acc<T>(S(t), (T *)t->bufs[I.in[p].rec.buffer].der, delta, (const T *)I.partial[p], o.size, 1, false));
                        break;
                    }
                    const T *tvs[kMaxArgs];
                    for (int a = 0; a < N; ++a) tvs[a] = I.in[a].rec.tracked ? (const T *)t->bufs[I.in[a].rec.buffer].tv : nullptr;
                    for (int p = 0; p < N; ++p) {
                        if (!I.in[p].rec.tracked) continue;                    // nothing flows into an untracked argument
                        const int rb = root(I.in[p].rec.buffer);
                        Buf &b = t->bufs[I.in[p].rec.buffer];
                        const T *sec[kMaxArgs];
                        for (int q = 0; q < N; ++q) sec[q] = (const T *)I.second[tri_index(p, q, N)];
                        StepTimer tm(t, "tangent_reverse", (double)(N + (dt ? 2 : 1) + (td_zero[rb] ? 0 : 1)) * o.size * Kb * t->es, 0);
                        bool same_shape = true;
                        for (int a = 0; a < N; ++a) same_shape &= I.in[a].size == o.size;
                        if (same_shape) {
                            if (td_zero[rb] && rb == root(xid)) OK(overwrite_x_td());
                            KL(t, launch_tangent_reverse<T>(S(t), (T *)b.td, dt, delta, (const T *)I.partial[p], N, sec, tvs, o.size, Kb, td_zero[rb] != 0));
                        } else {
                            // broadcasting closure: the accumulating kernel needs real zeros (or the running sum) in td
                            if (td_zero[rb]) {
                                if (rb == root(xid)) OK(zero_x_td()); else CU(cudaMemsetAsync(b.td, 0, (size_t)(b.size * Kb) * t->es, S(t)));
                            }
                            int64_t strides[kMaxArgs][kMaxDims], sizes[kMaxArgs];
                            bcast_strides(I, o, strides, sizes);
                            KL(t, launch_tangent_reverse_g<T>(S(t), (T *)b.td, dt, delta, (const T *)I.partial[p], N, p, sec, tvs, strides, sizes, o.dims, o.size, Kb));
                        }
                        td_zero[rb] = 0;
                        if (want_der(I.in[p].rec.buffer)) {
                            if (I.in[p].size == o.size) KL(t, launch_mul_acc<T>(S(t), (T *)b.der, delta, (const T *)I.partial[p], o.size, 1, false));
                            else {                         // clamped argument: deriv[g_p(e)] += delta[e] * partial[e]  (propagation.jl:117-188)
                                int64_t strides[kMaxArgs][kMaxDims], sizes[kMaxArgs];
                                bcast_strides(I, o, strides, sizes);
                                KL(t, launch_mul_acc_generic<T>(S(t), (T *)b.der, b.size, strides[p], delta, (const T *)I.partial[p], o.dims, o.size, 1));
                            }
                        }
                    }
                }
            }
            if (want_der(I.rec.out)) OK(unseed<T>(t, o, 1));
        }
        if (cols > 0) {
            if (td_zero[root(xid)]) {          // f does not depend on x through any second-order path: exact zeros
                if (prefilled) { CU(cudaStreamWaitEvent(S(t), t->join_ev, 0)); prefilled = false; }
                else CU(cudaMemsetAsync(x.td, 0, (size_t)(n * Kb) * t->es, S(t)));
            }
            if (prefilled) { CU(cudaStreamWaitEvent(S(t), t->join_ev, 0)); prefilled = false; }   // never leave the fork unjoined
            if (!in_place) {
                StepTimer tm(t, "hessian_cols_out", 2.0 * n * Kb * t->es, 0);
                KL(t, launch_copy2d<T>(S(t), dres + (c0 - cb) * dld, dld, (const T *)x.td, n, n, Kb));   // H[:, c0:c0+Kb] = tangent of deriv(x)
            }
        }
        first = false;
        if (cols == 0) break;
    }
    x.td = x_td_own;
    for (auto &b : t->bufs) if (b.alias_of >= 0) b.td = t->bufs[b.alias_of].td;
    for (size_t b = 0; b < t->bufs.size(); ++b) t->dz[droot(t, (int)b)] = 1;   // intermediates were physically unseeded
    if (t->ctx) t->ctx->gemm.uniform.clear();                                    // no deriv buffer is known to be a fill any more
    t->dz[droot(t, xid)] = 0;                                                   // deriv(x) holds the gradient (zeros in a lean sweep)
    if (grad_out) CU(cudaMemcpyAsync(grad_out, x.der, (size_t)n * t->es, cudaMemcpyDefault, S(t)));
    if (!on_device && cols > 0)
        CU(cudaMemcpy2DAsync(result, (size_t)ld * t->es, t->stage, (size_t)n * t->es, (size_t)n * t->es, (size_t)cols, cudaMemcpyDeviceToHost, S(t)));
    if (!t->capturing && !(t->async && on_device && !grad_out && !value_out)) CU(cudaStreamSynchronize(S(t)));
    return RDB_OK;
}

// hessian! entry: launch-bound tapes with a device result replay the whole sweep (forward with second partials, first- and
// second-order reverse, result block written in place) as ONE captured graph.  First call with a given (columns, result, ld)
// runs eagerly and warms every lazy allocation, the second captures, later ones only launch.
// ---- hessian! fast path: the tape is  v_1 = x[idx_1], ..., v_k = x[idx_k];  y = g.(v...) (or of x itself);  out = sum|mean(y).
// Then H = sum_i delta * d2g(x[maps(i)]) scattered at (map_p(i), map_q(i)): every column holds a handful of entries, and the call is
// the zero-fill of the block plus those entries - one kernel, no forward sweep, no tangents (src/api/hessians.jl:86-90 asks for
// the matrix only).  Values of the intermediates are NOT refreshed (forward_valid stays false).
template <typename T>
static int sep_hess_plan(rdb_tape *t) {
    auto &P = t->sep;
    P.state = -1;
    static const bool on = [] { const char *e = getenv("RDB_HESS_FUSED"); return !(e && e[0] == '0'); }();
    const size_t ni = t->instrs.size();
    if (!on || t->inputs.size() != 1 || ni < 2) return RDB_OK;
    auto root = [&](int b) { while (t->bufs[b].alias_of >= 0) b = t->bufs[b].alias_of; return b; };
    const int rx = root(t->inputs[0]);
    const int64_t n = t->bufs[rx].size;
    InstrPlan &R = t->instrs[ni - 1], &E = t->instrs[ni - 2];
    if ((R.rec.opcode != OP_SUM && R.rec.opcode != OP_MEAN) || root(R.rec.out) != root(t->output) || root(R.in[0].rec.buffer) != root(E.rec.out)) return RDB_OK;
    if (!is_elementwise(E.rec.opcode) || E.rec.opcode == OP_TRACKER_BCAST || E.rec.expr_len <= 0 || !E.d_ops) return RDB_OK;
    const int N = E.rec.n_in;
    if (N > 4) return RDB_OK;
    const int64_t m = t->bufs[E.rec.out].size;
    std::vector<int> producer(t->bufs.size(), -1);
    for (size_t ii = 0; ii + 2 < ni; ++ii) {
        InstrPlan &G = t->instrs[ii];
        if (!is_getindex(G.rec.opcode) || root(G.in[0].rec.buffer) != rx || !G.in[0].d_map) return RDB_OK;
        producer[root(G.rec.out)] = (int)ii;
    }
    std::vector<std::vector<int64_t>> maps(N);
    for (int a = 0; a < N; ++a) {
        OperandPlan &op = E.in[a];
        if (!op.rec.tracked || op.rec.cls != CLS_TA || t->bufs[op.rec.buffer].size != m) return RDB_OK;
        const int rb = root(op.rec.buffer);
        if (rb == rx) { P.map[a] = nullptr; continue; }            // the argument is x itself (m == n checked above)
        const int pr = producer[rb];
        if (pr < 0 || t->instrs[pr].in[0].rec.idx_len != m) return RDB_OK;
        P.map[a] = t->instrs[pr].in[0].d_map;
        maps[a].resize(m);
        CU(cudaMemcpyAsync(maps[a].data(), P.map[a], (size_t)m * 8, cudaMemcpyDeviceToHost, S(t)));
    }
    CU(cudaStreamSynchronize(S(t)));
    std::vector<int64_t> off(n + 1, 0);
    auto col_of = [&](int a, int64_t i) { return P.map[a] ? maps[a][i] : i; };
    for (int a = 0; a < N; ++a) for (int64_t i = 0; i < m; ++i) ++off[col_of(a, i) + 1];
    for (int64_t j = 0; j < n; ++j) { if (off[j + 1] > kSepHessMaxEntries) return RDB_OK; off[j + 1] += off[j]; }
    std::vector<int64_t> inv(off[n]), fill(off.begin(), off.end() - 1);
    for (int64_t i = 0; i < m; ++i) for (int a = 0; a < N; ++a) inv[fill[col_of(a, i)]++] = i * N + a;     // ascending (element, argument)
    OK(dev_alloc(t, (void **)&P.d_inv_off, (n + 1) * 8));
    OK(dev_alloc(t, (void **)&P.d_inv, std::max<int64_t>(1, off[n]) * 8));
    CU(cudaMemcpyAsync(P.d_inv_off, off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, S(t)));
    if (off[n]) CU(cudaMemcpyAsync(P.d_inv, inv.data(), (size_t)off[n] * 8, cudaMemcpyHostToDevice, S(t)));
    CU(cudaStreamSynchronize(S(t)));
    P.ew = (int)ni - 2;
    P.delta = R.rec.opcode == OP_MEAN ? (double)((T)1 / (T)m) : 1.0;
    P.state = 1;
    return RDB_OK;
}
template <typename T>
static int sep_hess_run(rdb_tape *t, int64_t cb, int64_t ce, void *result, int64_t ld) {
    auto &P = t->sep;
    InstrPlan &E = t->instrs[P.ew];
    Buf &x = t->bufs[t->inputs[0]];
    const void *xin = x.val;
    if (P.lazy_x) { xin = P.lazy_x; P.lazy_x = nullptr; P.x_consumed = true; }     // read in place; the tape's copy is not refreshed
    else OK(settle_uploads(t, -1));
    SepHessParams K{};
    K.n_args = E.rec.n_in; K.n_ops = (int)E.rec.expr_len;
    for (int a = 0; a < K.n_args; ++a) K.map[a] = P.map[a];
    K.inv_off = P.d_inv_off; K.inv = P.d_inv; K.ops = E.d_ops; K.fconst = t->d_fconst;
    K.x = xin; K.n_x = x.size; K.H = result; K.ld = ld; K.c0 = cb; K.K = ce - cb; K.delta = P.delta;
    {
        StepTimer tm(t, "hessian_separable", (double)x.size * K.K * t->es, 0);
        KL(t, launch_hessian_separable<T>(S(t), K));
    }
    t->launches += 1;                                       // (KL counted one: the call is a fill kernel + its dependent patch kernel)
    t->forward_valid = false;
    if (!t->async) CU(cudaStreamSynchronize(S(t)));
    return RDB_OK;
}

template <typename T>
static int hessian_entry(rdb_tape *t, int64_t cb, int64_t ce, void *result, int64_t ld, int on_device, void *grad_out, void *value_out) {
    if (on_device && !grad_out && !value_out && ce > cb && t->inputs.size() == 1 && cb >= 0 && ce <= t->bufs[t->inputs[0]].size &&
        ld >= t->bufs[t->inputs[0]].size && t->bufs[t->output].size == 1) {
        if (t->sep.state == 0) OK(sep_hess_plan<T>(t));
        if (t->sep.state == 1) return sep_hess_run<T>(t, cb, ce, result, ld);
    }
    auto &G = t->hgraph;
    GemmState &GS = t->ctx->gemm;
    const bool want = graph_eligible(t) && on_device && !grad_out && !value_out && ce > cb && t->inputs.size() == 1;
    NoCacheScope nc(GS, want);
    if (G.state == 2 && t->hgraph_gen != GS.alloc_gen) {     // the context's GEMM workspace moved since the capture
        cudaGraphExecDestroy(G.exec); G.exec = nullptr; G.state = 1;
    }
    const bool same = want && G.cb == cb && G.ce == ce && G.ld == ld && G.result == result;
    if (!want || !same) {
        if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
        G.state = want ? 1 : 0; G.cb = cb; G.ce = ce; G.ld = ld; G.result = result;
        return hessian_impl<T>(t, cb, ce, result, ld, on_device, grad_out, value_out);
    }
    if (G.state == 1) {
        OK(settle_uploads(t, -1));
        G.state = -1;
        CU(guard_fold(S(t), GS));
        const uint64_t gen0 = GS.alloc_gen;
        if (cudaStreamBeginCapture(S(t), cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            return hessian_impl<T>(t, cb, ce, result, ld, on_device, grad_out, value_out);
        }
        t->capturing = true;
        const int rc = hessian_impl<T>(t, cb, ce, result, ld, on_device, grad_out, value_out);
        t->capturing = false;
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(S(t), &g);
        t->hgraph_guard_n = GS.guard_used;
        GS.guard_used = 0;
        if (rc != RDB_OK || e != cudaSuccess || !g) {
            cudaGetLastError();
            if (g) cudaGraphDestroy(g);
            return hessian_impl<T>(t, cb, ce, result, ld, on_device, grad_out, value_out);      // eager from now on (state -1)
        }
        if (GS.alloc_gen != gen0) {                      // the workspace grew inside the capture: eager now, capture next time
            cudaGraphDestroy(g);
            G.state = 1;
            return hessian_impl<T>(t, cb, ce, result, ld, on_device, grad_out, value_out);
        }
        const cudaError_t ei = cudaGraphInstantiate(&G.exec, g, 0);
        cudaGraphDestroy(g);
        if (ei != cudaSuccess) { cudaGetLastError(); G.exec = nullptr; return hessian_impl<T>(t, cb, ce, result, ld, on_device, grad_out, value_out); }
        G.state = 2;
        t->hgraph_gen = GS.alloc_gen;
    }
    if (G.state != 2) return hessian_impl<T>(t, cb, ce, result, ld, on_device, grad_out, value_out);
    OK(settle_uploads(t, -1));
    if (GS.guard_used != 0) CU(guard_fold(S(t), GS));
    CU(cudaGraphLaunch(G.exec, S(t)));
    GS.guard_used = t->hgraph_guard_n;
    t->launches += 1;
    // host-side state the sweep leaves behind (tail of hessian_impl)
    for (size_t b = 0; b < t->bufs.size(); ++b) t->dz[droot(t, (int)b)] = 1;
    if (t->ctx) t->ctx->gemm.uniform.clear();                                    // no deriv buffer is known to be a fill any more
    t->dz[droot(t, t->inputs[0])] = 0;
    t->forward_valid = true;
    if (!t->async) CU(cudaStreamSynchronize(S(t)));
    return RDB_OK;
}

// ================================================================================================ C ABI
static int fold_launches(rdb_tape *t, int rc) {
    t->launches += g_extra_launches;
    g_extra_launches = 0;
    // slicing that ran ahead on the side stream and was not consumed must not outlive the call (the next call may refill the
    // same cache entries from the main stream)
    if (t->ctx && t->ctx->side_stream && !t->async && !t->async_call) cudaStreamSynchronize(t->ctx->side_stream);
    return rc;
}
// Run one API call on a tape.  FP64 `*` instructions of eligible shapes go through exact int8 slices on tcgen05, whose error is
// norm-wise (2^-56 of the row/column maxima), not component-wise like DGEMM's.  Every such launch therefore leaves a rigorous
// a-posteriori certificate (umma.cu: max error bound <= 2^-40 * max |product|); when any launch of the call fails it, the call
// is replayed with the DMMA kernels (plain FP64 FMA arithmetic, what the reference's BLAS does) before it returns.  Three failing
// calls in a row make the tape DMMA-only.  gemm_f64_mode: 0 = this, 1 = DMMA only, 2 = int8 slices without the replay.
template <typename F>
static int run_call(rdb_tape *t, F &&run) {
    rdb_ctx *c = t->ctx;
    ctx_enter(c);
    g_f64_mode = t->opts.gemm_f64_mode;
    g_f64_slices = t->opts.ozaki_slices;
    g_extra_launches = 0;
    int rc = fold_launches(t, run());
    if (t->async_call) { t->pending.guard = rc == RDB_OK && t->dtype == RDB_F64 && t->opts.gemm_f64_mode != 1; return rc; }   // rdb_tape_wait checks
    if (rc != RDB_OK || t->dtype != RDB_F64 || t->opts.gemm_f64_mode == 1 || !c->gemm.guard) return rc;
    if (c->gemm.guard_used == 0 && !c->gemm.guard_pending) return rc;   // no int8-slice launch in this call
    GuardSummary gs;
    CU(guard_read(S(t), c->gemm, &gs));
    t->guard_checked += (int64_t)gs.checked;
    t->slice_pairs += (int64_t)gs.pairs;
    t->guard_flagged += (int64_t)gs.flagged;
    if (gs.worst_ratio > t->guard_worst) t->guard_worst = gs.worst_ratio;
    if (gs.flagged == 0) { t->guard_streak = 0; return rc; }
    if (t->opts.gemm_f64_mode != 0) return rc;
    t->guard_fallbacks++;
    t->dmma_now = true;
    g_f64_mode = 1;
    g_extra_launches = 0;
    rc = fold_launches(t, run());
    t->dmma_now = false;
    if (++t->guard_streak >= 3) t->opts.gemm_f64_mode = 1;
    return rc;
}
#define DISPATCH(t, fn, ...) \
    run_call((t), [&]() -> int { return (t)->dtype == RDB_F64 ? fn<double>(__VA_ARGS__) : fn<float>(__VA_ARGS__); })

extern "C" {

const char *rdb_version(void) { return "rdb200 0.1 (sm_100a)"; }
const char *rdb_last_error(void) { return g_err; }

rdb_ctx *rdb_ctx_create(int device_id) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        fail(RDB_ENODEVICE, "no CUDA device (%s): the B200 engine has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
        return nullptr;
    }
    if (device_id < 0 || device_id >= n) { fail(RDB_EINVAL, "device %d out of range (%d devices)", device_id, n); return nullptr; }
    if (cudaSetDevice(device_id) != cudaSuccess) { fail(RDB_ECUDA, "cudaSetDevice(%d) failed", device_id); return nullptr; }
    auto *c = new rdb_ctx();
    c->device = device_id;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { fail(RDB_ECUDA, "cudaStreamCreate failed"); delete c; return nullptr; }
    return c;
}
void rdb_ctx_destroy(rdb_ctx *c) { ctx_release(c); }
int rdb_ctx_set_stream(rdb_ctx *c, void *stream) {
    if (!c) return fail(RDB_EINVAL, "null context");
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)stream;
    c->own_stream = false;
    return RDB_OK;
}
int rdb_ctx_set_gemm_f64_mode(rdb_ctx *c, int mode, int slices) {
    if (!c) return fail(RDB_EINVAL, "null context");
    if (mode < 0 || mode > 2 || slices < 0 || slices > 8) return fail(RDB_EINVAL, "bad gemm_f64_mode/slices");
    c->gemm_f64_mode = mode;
    c->ozaki_slices = slices;
    return RDB_OK;
}
int rdb_ctx_synchronize(rdb_ctx *c) {
    if (!c) return fail(RDB_EINVAL, "null context");
    ctx_enter(c);
    CU(cudaStreamSynchronize(c->stream));
    return RDB_OK;
}
int rdb_ctx_gemm_guard(rdb_ctx *c, int64_t *checked, int64_t *flagged, double *worst_ratio) {
    if (!c) return fail(RDB_EINVAL, "null context");
    ctx_enter(c);
    GuardSummary gs;
    CU(guard_read(c->stream, c->gemm, &gs));
    if (checked) *checked = (int64_t)gs.checked;
    if (flagged) *flagged = (int64_t)gs.flagged;
    if (worst_ratio) *worst_ratio = gs.worst_ratio;
    return RDB_OK;
}
int rdb_tape_set_async(rdb_tape *t, int on) {
    if (!t) return fail(RDB_EINVAL, "null tape");
    t->async = on != 0;
    return RDB_OK;
}
int64_t rdb_tape_slice_pairs(rdb_tape *t) { return t ? t->slice_pairs : -1; }
int rdb_tape_guard_stats(rdb_tape *t, int64_t *checked, int64_t *flagged, int64_t *fallbacks, double *worst_ratio) {
    if (!t) return fail(RDB_EINVAL, "null tape");
    if (checked) *checked = t->guard_checked;
    if (flagged) *flagged = t->guard_flagged;
    if (fallbacks) *fallbacks = t->guard_fallbacks;
    if (worst_ratio) *worst_ratio = t->guard_worst;
    return RDB_OK;
}

rdb_tape *rdb_tape_load(rdb_ctx *c, const void *program, size_t nbytes, const rdb_options *opts) {
    if (!c || !program) { fail(RDB_EINVAL, "null argument"); return nullptr; }
    ctx_enter(c);
    auto *t = new rdb_tape();
    t->ctx = c;
    c->refs++;
    t->opts.fuse_elementwise = 1; t->opts.use_graph = 1;
    if (opts) t->opts = *opts;
    cudaEventCreate(&t->ev0);
    cudaEventCreate(&t->ev1);
    if (load_program(t, (const uint8_t *)program, nbytes) != RDB_OK) { rdb_tape_destroy(t); return nullptr; }
    return t;
}
void rdb_tape_destroy(rdb_tape *t) {
    if (!t) return;
    if (t->ctx) { ctx_enter(t->ctx); cudaStreamSynchronize(t->ctx->stream); }
    for (void *p : t->allocs) cudaFree(p);
    for (auto &I : t->instrs) if (I.blk) scalar_block_free(*I.blk);
    if (t->ev0) cudaEventDestroy(t->ev0);
    if (t->ev1) cudaEventDestroy(t->ev1);
    if (t->gexec) cudaGraphExecDestroy(t->gexec);
    if (t->hgraph.exec) cudaGraphExecDestroy(t->hgraph.exec);
    if (t->fork_ev) cudaEventDestroy(t->fork_ev);
    if (t->join_ev) cudaEventDestroy(t->join_ev);
    if (t->done_ev) cudaEventDestroy(t->done_ev);
    for (auto &e : t->early.ev) if (e) cudaEventDestroy(e);
    if (t->ctx && t->ctx->upload_stream) cudaStreamSynchronize(t->ctx->upload_stream);
    for (auto &b : t->bufs) for (auto &e : b.up_ev) if (e) cudaEventDestroy(e);
    ctx_release(t->ctx);
    delete t;
}

int rdb_tape_set_input(rdb_tape *t, int slot, const void *data, int is_device) {
    if (!t || !data) return fail(RDB_EINVAL, "null argument");
    if (slot < 0 || (size_t)slot >= t->inputs.size()) return fail(RDB_EINVAL, "input slot %d out of range", slot);
    ctx_enter(t->ctx);
    Buf &b = t->bufs[t->inputs[slot]];
    static const bool stream_inputs = [] { const char *e = getenv("RDB_STREAM_INPUTS"); return !(e && e[0] == '0'); }();
    const bool matrix = b.nd == 2 && b.dims[0] * b.dims[1] == b.size && b.alias_of < 0;
    if (!is_device && stream_inputs && matrix && b.dims[1] >= 2048 && b.size * (int64_t)t->es >= (32 << 20)) {
        // value!(input, x) as column panels on the upload stream; the forward sweep waits per panel or for all (forward_pass)
        rdb_ctx *c = t->ctx;
        if (!c->upload_stream) CU(cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking));
        int64_t width[4];
        plan_col_panels(b.dims[0], b.dims[1], width);
        b.up_n = 0;
        int64_t c0 = 0;
        for (int p = 0; p < 4; ++p) {
            if (width[p] <= 0) continue;
            const size_t off = (size_t)c0 * b.dims[0] * t->es, nb = (size_t)width[p] * b.dims[0] * t->es;
            CU(cudaMemcpyAsync((char *)b.val + off, (const char *)data + off, nb, cudaMemcpyHostToDevice, c->upload_stream));
            if (!b.up_ev[b.up_n]) CU(cudaEventCreateWithFlags(&b.up_ev[b.up_n], cudaEventDisableTiming));
            CU(cudaEventRecord(b.up_ev[b.up_n], c->upload_stream));
            b.up_c0[b.up_n] = c0;
            c0 += width[p];
            b.up_c0[++b.up_n] = c0;
        }
        t->last_upload = t->inputs[slot];
    } else if (is_device && t->async && t->sep.state == 1 && slot == 0) {
        // asynchronous mode + the one-kernel hessian!: the kernel reads the caller's array in place (the header already asks for
        // device inputs to stay alive until the stream is synchronised); any other entry point copies it first (settle_uploads)
        t->sep.lazy_x = data;
        t->sep.x_consumed = false;
    } else {
        if (b.up_n > 0) { CU(cudaStreamWaitEvent(S(t), b.up_ev[b.up_n - 1], 0)); b.up_n = 0; }
        CU(cudaMemcpyAsync(b.val, data, (size_t)b.size * t->es, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, S(t)));
        t->sep.lazy_x = nullptr; t->sep.x_consumed = false;
    }
    bump_version(t, t->inputs[slot]);
    t->forward_valid = false;
    return RDB_OK;
}
int rdb_forward(rdb_tape *t) {
    if (!t) return fail(RDB_EINVAL, "null tape");
    OK(DISPATCH(t, forward_pass, t, 1));
    CU(cudaStreamSynchronize(S(t)));
    return RDB_OK;
}
int rdb_reverse_scalar_seed(rdb_tape *t) {
    if (!t) return fail(RDB_EINVAL, "null tape");
    ctx_enter(t->ctx);
    OK(settle_uploads(t, -1));
    OK(DISPATCH(t, seed_scalar_and_reverse, t));
    CU(cudaStreamSynchronize(S(t)));
    return RDB_OK;
}
int rdb_gradient(rdb_tape *t, void *const *result_ptrs, int on_device, void *value_out) {
    if (!t) return fail(RDB_EINVAL, "null tape");
    return DISPATCH(t, gradient_impl, t, result_ptrs, on_device, value_out);
}
int rdb_gradient_async(rdb_tape *t, void *const *result_ptrs, int on_device, void *value_out) {
    if (!t) return fail(RDB_EINVAL, "null tape");
    if (t->pending.active) return fail(RDB_EINVAL, "this tape has an asynchronous call in flight: rdb_tape_wait first");
    ctx_enter(t->ctx);
    if (!t->done_ev) CU(cudaEventCreateWithFlags(&t->done_ev, cudaEventDisableTiming));
    t->pending.results.assign(t->inputs.size(), nullptr);
    if (result_ptrs) for (size_t i = 0; i < t->inputs.size(); ++i) t->pending.results[i] = result_ptrs[i];
    t->pending.on_device = on_device; t->pending.value_out = value_out; t->pending.guard = false;
    t->async_call = true;
    const int rc = DISPATCH(t, gradient_impl, t, result_ptrs, on_device, value_out);
    t->async_call = false;
    t->pending.active = rc == RDB_OK;
    return rc;
}
int rdb_tape_wait(rdb_tape *t) {
    if (!t) return fail(RDB_EINVAL, "null tape");
    if (!t->pending.active) return RDB_OK;
    rdb_ctx *c = t->ctx;
    ctx_enter(c);
    t->pending.active = false;
    CU(cudaEventSynchronize(t->done_ev));
    if (c->side_stream) CU(cudaStreamSynchronize(c->side_stream));      // slicing that ran ahead and was not consumed
    if (!t->pending.guard || !c->gemm.guard || (c->gemm.guard_used == 0 && !c->gemm.guard_pending)) return RDB_OK;
    // the accuracy certificate of the call's int8-slice GEMMs, deferred from run_call
    GuardSummary gs;
    CU(guard_read(S(t), c->gemm, &gs));
    t->guard_checked += (int64_t)gs.checked;
    t->slice_pairs += (int64_t)gs.pairs;
    t->guard_flagged += (int64_t)gs.flagged;
    if (gs.worst_ratio > t->guard_worst) t->guard_worst = gs.worst_ratio;
    if (gs.flagged == 0) { t->guard_streak = 0; return RDB_OK; }
    if (t->opts.gemm_f64_mode != 0) return RDB_OK;
    // replay with the DMMA kernels, synchronously; the inputs are still in the tape's buffers
    t->guard_fallbacks++;
    const int mode0 = t->opts.gemm_f64_mode;
    t->opts.gemm_f64_mode = 1;
    t->dmma_now = true;
    void *const *rp = t->pending.results.data();
    const int on_device = t->pending.on_device;
    void *value_out = t->pending.value_out;
    const int rc = DISPATCH(t, gradient_impl, t, rp, on_device, value_out);
    t->dmma_now = false;
    t->opts.gemm_f64_mode = (++t->guard_streak >= 3) ? 1 : mode0;
    return rc;
}
int rdb_jacobian(rdb_tape *t, int slot, int64_t rb, int64_t re, void *result, int64_t ld, int on_device, void *value_out) {
    if (!t || (!result && re > rb)) return fail(RDB_EINVAL, "null argument");
    return DISPATCH(t, jacobian_impl, t, slot, rb, re, result, ld, on_device, value_out);
}
int rdb_hessian(rdb_tape *t, int64_t cb, int64_t ce, void *result, int64_t ld, int on_device, void *grad_out, void *value_out) {
    if (!t || (!result && ce > cb)) return fail(RDB_EINVAL, "null argument");
    return DISPATCH(t, hessian_entry, t, cb, ce, result, ld, on_device, grad_out, value_out);
}
int64_t rdb_buffer_size(rdb_tape *t, int b) {
    if (!t || b < 0 || (size_t)b >= t->bufs.size()) return -1;
    return t->bufs[b].size;
}
int rdb_get_value(rdb_tape *t, int b, void *dst) {
    if (!t || !dst || b < 0 || (size_t)b >= t->bufs.size()) return fail(RDB_EINVAL, "bad buffer id");
    ctx_enter(t->ctx);
    OK(settle_uploads(t, -1));
    CU(cudaMemcpyAsync(dst, t->bufs[b].val, (size_t)t->bufs[b].size * t->es, cudaMemcpyDeviceToHost, S(t)));
    CU(cudaStreamSynchronize(S(t)));
    return RDB_OK;
}
int rdb_get_deriv(rdb_tape *t, int b, void *dst) {
    if (!t || !dst || b < 0 || (size_t)b >= t->bufs.size()) return fail(RDB_EINVAL, "bad buffer id");
    if (!t->bufs[b].der) return fail(RDB_EINVAL, "buffer %d has no deriv (CONST)", b);
    ctx_enter(t->ctx);
    OK(dmaterialize(t, b, std::max<int64_t>(t->R_alloc, 1), false));
    CU(cudaMemcpyAsync(dst, t->bufs[b].der, (size_t)t->bufs[b].size * t->es, cudaMemcpyDeviceToHost, S(t)));
    CU(cudaStreamSynchronize(S(t)));
    return RDB_OK;
}
int rdb_tape_stats(rdb_tape *t, char *json, size_t cap) {
    if (!t || !json || cap < 3) return fail(RDB_EINVAL, "bad argument");
    std::string s = "[";
    char buf[512];
    for (size_t i = 0; i < t->stats.size(); ++i) {
        auto &st = t->stats[i];
        snprintf(buf, sizeof(buf), "%s{\"step\":\"%s\",\"count\":%d,\"ms\":%.6f,\"bytes\":%.0f,\"flops\":%.0f}", i ? "," : "", st.name.c_str(), st.count, st.ms, st.bytes, st.flops);
        s += buf;
    }
    s += "]";
    if (s.size() + 1 > cap) return fail(RDB_EINVAL, "stats buffer too small");
    memcpy(json, s.c_str(), s.size() + 1);
    t->stats.clear();
    return RDB_OK;
}
int64_t rdb_tape_launch_count(rdb_tape *t) { return t ? t->launches : -1; }


int rdb_gemm(rdb_ctx *c, int dtype, int ta, int tb, int64_t m, int64_t n, int64_t k, double alpha, const void *A, int64_t lda,
             const void *B, int64_t ldb, double beta, void *C, int64_t ldc) {
    if (!c) return fail(RDB_EINVAL, "null context");
    ctx_enter(c);
    cudaError_t e = dtype == RDB_F64
                        ? gemm_f64(c->stream, ta != 0, tb != 0, m, n, k, alpha, (const double *)A, lda, (const double *)B, ldb, beta, (double *)C, ldc,
                                   c->gemm_f64_mode, c->ozaki_slices)
                        : gemm_f32(c->stream, ta != 0, tb != 0, m, n, k, (float)alpha, (const float *)A, lda, (const float *)B, ldb, (float)beta, (float *)C, ldc);
    if (e != cudaSuccess) return fail(RDB_ECUDA, "gemm launch failed: %s", cudaGetErrorString(e));
    return RDB_OK;
}

int rdb_add_sum(rdb_ctx *c, int dtype, int64_t n, const void *a, const void *b, double sign, void *out, void *sum_out) {
    if (!c) return fail(RDB_EINVAL, "null context");
    ctx_enter(c);
    static void *scratch = nullptr;
    if (!scratch && cudaMalloc(&scratch, kReduceBlocks * 8) != cudaSuccess) return fail(RDB_ENOMEM, "scratch alloc failed");
    cudaError_t e;
    if (dtype == RDB_F64) {
        e = launch_add<double>(c->stream, (double *)out, (const double *)a, (const double *)b, sign, n, (double *)scratch);
        if (e == cudaSuccess) e = launch_finish_sum<double>(c->stream, (const double *)scratch, (double *)sum_out, 1.0);
    } else {
        e = launch_add<float>(c->stream, (float *)out, (const float *)a, (const float *)b, (float)sign, n, (float *)scratch);
        if (e == cudaSuccess) e = launch_finish_sum<float>(c->stream, (const float *)scratch, (float *)sum_out, 1.f);
    }
    if (e != cudaSuccess) return fail(RDB_ECUDA, "add_sum launch failed: %s", cudaGetErrorString(e));
    return RDB_OK;
}

int64_t rdb_schedule_bundles(int64_t n_nodes, const int64_t *succ_off, const int32_t *succ, int32_t *bundle, int32_t *lane) {
    return scalar_schedule_bundles_host(n_nodes, succ_off, succ, bundle, lane);
}

int rdb_plan_col_panels(int64_t rows, int64_t cols, int64_t widths[4]) {
    if (!widths || rows <= 0 || cols <= 0) return 0;
    plan_col_panels(rows, cols, widths);
    int n = 0;
    for (int i = 0; i < 4; ++i) n += widths[i] > 0;
    return n;
}

}  // extern "C"
