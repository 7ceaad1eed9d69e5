// Scalar-instruction blocks: a maximal run of consecutive ScalarInstructions (reference src/tape.jl:30-45) replayed as one unit.
//
// Reference semantics (src/derivatives/scalars.jl):
//   forward  :80-123  per instruction, in tape order: pull_value!(inputs); ONE ForwardDiff.derivative!/gradient! evaluation of the
//                     scalar function; value!(output); cache[] = partial(s)
//   reverse  :52-74   per instruction, in REVERSE tape order: increment_deriv!(a, deriv(output)*a_partial) [then b]; unseed!(output)
// A TrackedReal obtained by scalar getindex carries no instruction: its value/deriv are pulled from / pushed to origin[index]
// (src/tracked.jl:178-197,348-351).
//
// B200 mapping.  The CPU runs these loops one instruction at a time; a tape of 1e5-1e6 scalar instructions has no other
// parallelism for ONE seed than its dataflow, and R independent seeds for jacobian!/hessian!.  So the planner (host, at load)
//   * levelises the forward dataflow: instructions of one level are evaluated in parallel (values do not depend on the order);
//   * turns the reverse loop into its PULL form: the adjoint of slot s is  init(s) + sum over the instructions c that consume s,
//     in DESCENDING tape order (operand a before b), of deriv(out_c)*partial_c — exactly the sequence of increments the
//     reference applies to s.deriv, so every sum is built in the reference's order, but each adjoint is written ONCE, by one
//     thread, with no read-modify-write, no atomics and no unseed! traffic (a slot is final before anyone reads it);
//   * executes both graphs with the kernel that fits their shape (all of them produce the same sums in the same order):
//       - shallow, wide graphs: one launch per dataflow level (k_scalar_forward, k_scalar_reverse_slots);
//       - deep graphs (long `acc += term` chains), forward and reverse with few seeds: the dataflow list-scheduled into bundles
//         of <= 32 independent slots executed by ONE warp (k_scalar_forward_bundles, k_scalar_reverse_bundles);
//       - 16 or more seeds: the streaming kernel (lanes = seeds, one warp walks the whole schedule for its 32 seeds without any
//         synchronisation, k_scalar_reverse_stream) or one bundle warp per seed, whichever a fitted cost model prefers
//         (scalar_reverse_streams).
//     In every kernel the adjoints that are still "near" live in a shared-memory ring and only the ones with a far reader are
//     ever stored to global memory; records, partials and far operands are staged ahead of their use by cp.async.
//   * partials are stored by the forward sweep directly where the reverse kernels read them (reverse-edge order, plus a
//     bundle-ordered copy for the bundle kernel).
// Compiled with -fmad=false (parity of deriv*partial + accumulate with the uncontracted CPU arithmetic).
#include <cuda_runtime.h>
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include "kernels.h"
#include "expr_eval.cuh"

namespace rdb {

namespace {
constexpr int kFwdThreads = 512;
constexpr int kWideLevel = 8192;      // a level at least this wide gets a grid of its own
constexpr int kNone = INT_MIN;

struct FwdP {
    const int4 *finstr; const int64_t *level_off; const int4 *ops; const int4 *exprs; const int32_t *epos; const int32_t *bpos; void *bpart;
    const double *fconst; void *val; void *epart; int64_t n_leaf; int lv0, lv1;
};

template <typename T>
__global__ void __launch_bounds__(kFwdThreads) k_scalar_forward(const FwdP P) {
    T *val = reinterpret_cast<T *>(P.val);
    T *epart = reinterpret_cast<T *>(P.epart);
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    for (int lv = P.lv0; lv < P.lv1; ++lv) {
        const int64_t j0 = P.level_off[lv], j1 = P.level_off[lv + 1];
        for (int64_t j = j0 + tid; j < j1; j += nt) {
            const int4 I = P.finstr[j];
            const int4 E = P.exprs[I.w];
            T x[2];
            x[0] = I.y >= 0 ? val[I.y] : (I.y == kNone ? (T)0 : (T)P.fconst[-2 - I.y]);     // pull_value!(a)
            x[1] = I.z >= 0 ? val[I.z] : (I.z == kNone ? (T)0 : (T)P.fconst[-2 - I.z]);     // pull_value!(b)
            T v, d[2], h[1];
            eval_expr<T, 2, 1>(P.ops + E.x, E.y, P.fconst, x, v, d, h, E.z);
            val[P.n_leaf + I.x] = v;                                                        // value!(output, ...)
            const int ea = P.epos[2 * (int64_t)I.x], eb = P.epos[2 * (int64_t)I.x + 1];
            if (ea >= 0) epart[ea] = d[0];                                                  // cache[] = partials
            if (eb >= 0) epart[eb] = d[1];
            if (P.bpos) {                                                                   // ... and in bundle order (k_scalar_reverse_bundles)
                T *bpart = reinterpret_cast<T *>(P.bpart);
                const int ba = P.bpos[2 * (int64_t)I.x], bb = P.bpos[2 * (int64_t)I.x + 1];
                if (ba >= 0) bpart[ba] = d[0];
                if (bb >= 0) bpart[bb] = d[1];
            }
        }
        if (lv + 1 < P.lv1) __syncthreads();      // multi-level launches run on ONE CTA
    }
}

template <typename T>
__global__ void k_scalar_pull(T *val, const int32_t *leaf_ord, const int64_t *leaf_elem, void *const *ext_val, int64_t n_leaf,
                              int64_t n_instr, const double *fconst, int64_t n_lit) {
    int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (l < n_leaf) val[l] = reinterpret_cast<const T *>(ext_val[leaf_ord[l]])[leaf_elem[l]];
    else if (l < n_leaf + n_lit) val[n_instr + l] = (T)fconst[l - n_leaf];      // literal pool behind the slots (bundle kernels)
}
template <typename T>
__global__ void k_scalar_export(const T *val, const int32_t *slot, const int32_t *ord, const int64_t *elem, void *const *ext_val,
                                int64_t n_leaf, int64_t n_exp) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n_exp) reinterpret_cast<T *>(ext_val[ord[k]])[elem[k]] = val[n_leaf + slot[k]];
}

struct RevP {
    const int32_t *rslot; const int64_t *reoff; const int32_t *esrc; const void *epart; const int64_t *rlevel_off;
    const int64_t *xoff; const int32_t *xord; const int64_t *xelem;
    const int32_t *leaf_ord; const int64_t *leaf_elem;
    void *const *ext_der; const int64_t *ext_size;
    void *der; int64_t R; int lv0, lv1;
};

// ---- R >= 16: the streaming kernel.  Lanes are seeds; ONE warp owns 32 seeds for the whole sweep and walks the slots in the
// reference's own order (reverse tape order; a leaf right after its last consumer).  Seeds never interact and a lane only ever
// reads adjoints it wrote itself, so there is NO synchronisation at all - not between CTAs, not between warps.  What is left
// is latency, and the host plan (scalar_block_build) removes it statically, per block of 32 schedule positions:
//   * the adjoint produced by the step right before is forwarded in a register;
//   * an adjoint produced in the same or the previous block is read from a 64-entry shared-memory ring (row = position & 63);
//   * an adjoint produced earlier ("far") lives in global memory ([row][seed], only slots with a far reader get a row) and is
//     brought into a shared-memory staging buffer by cp.async one block AHEAD of its use, as are the adjoints that later
//     instructions pushed into exported slots; the copy overlaps the 32 steps of the running block;
//   * the records of the next block (one per lane, coalesced) and their partials are loaded at the same time;
//   * the step loop is software-pipelined by hand: the record of step k+2 and the shared-memory operands of step k+1 are
//     loaded while step k computes (the plan guarantees they do not depend on step k except through the forwarded register).
// A step with at most two consumers, no exports and no leaf (the bulk of any scalar tape) is ~25 branch-free instructions; the
// rest takes one uniform branch into a general path.
constexpr int kBlk = 32;              // schedule positions per block (= lanes that load one record each)
constexpr int kStageF = 32;           // staged operands per block
constexpr int kRing = 64;             // two blocks of adjoints
constexpr int kInl = 4;               // edges whose partials are staged with the record; the rest loop over tsrc/epart
constexpr unsigned kSrcDirect = 0u, kSrcSmem = 1u, kSrcPrev = 3u;
inline int src_code(unsigned kind, unsigned payload) { return (int)((kind << 30) | payload); }
constexpr int kM_E0 = 1, kM_E1 = 2, kM_Prev0 = 4, kM_Prev1 = 8, kM_Rare = 16, kM_Dir0 = 64, kM_Dir1 = 128, kM_Exp = 256,
              kM_ExpStaged = 512, kM_Leaf = 1024;

struct StreamP {
    const int4 *recA, *recB, *recC, *recD; const int2 *pf; const int32_t *tsrc; const void *epart;
    const int64_t *xoff; const int32_t *xord; const int64_t *xelem;
    void *const *ext_der; const int64_t *ext_size;
    void *der; int64_t R; int64_t n_blocks; int leaves_zero;
};

template <int BYTES>
__device__ __forceinline__ void cp_async_small(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// record of schedule position q (shared-memory rows: [0, 64) ring, [64, 96) staging buffer of even blocks, [96, 128) of odd blocks):
//   A = (byte offset of the row holding the adjoint of edge 0's consumer; same for edge 1; kM_* flags; global row of this slot or -1)
//   B = (first edge in tsrc/epart; number of edges; instruction id (exports); staging entry of the export)
//   C = (source of edge 2; source of edge 3; global row read by edge 0 if kM_Dir0; by edge 1 if kM_Dir1)
//   D = leaf: (ordinal of the origin buffer; element, low word; element, high word; 0)
//   sources: kind << 30 | payload - kSrcSmem: shared-memory row, kSrcDirect: global row, kSrcPrev: the previous step's adjoint
template <typename T>
__global__ void __launch_bounds__(32) k_scalar_reverse_stream(const StreamP P) {
    __shared__ __align__(16) T comb[kRing + 2 * kStageF][32];
    __shared__ __align__(16) T spart[2][kBlk + 2][kInl];
    __shared__ int4 sA[2][kBlk + 2], sB[2][kBlk], sC[2][kBlk], sD[2][kBlk];
    __shared__ int2 spf[2][kStageF];
    const int lane = threadIdx.x;
    const int64_t R = P.R, r = blockIdx.x * 32LL + lane;
    const bool active = r < R;
    const int64_t rr = active ? r : 0;                 // idle lanes of a ragged last warp follow seed 0 and never store
    T *der = reinterpret_cast<T *>(P.der);
    const T *epart = reinterpret_cast<const T *>(P.epart);
    if (lane < 2) {                                    // the two records the pipeline reads past the end of a block
        sA[0][kBlk + lane] = sA[1][kBlk + lane] = make_int4(0, 0, 0, -1);
        for (int k = 0; k < kInl; ++k) spart[0][kBlk + lane][k] = spart[1][kBlk + lane][k] = (T)0;
    }

    auto load_block = [&](int64_t b, int buf) {
        const int64_t q = b * kBlk + lane;
        const int4 Bq = P.recB[q];
        sA[buf][lane] = P.recA[q]; sB[buf][lane] = Bq; sC[buf][lane] = P.recC[q]; sD[buf][lane] = P.recD[q];
#pragma unroll
        for (int k = 0; k < kInl; ++k) spart[buf][lane][k] = k < Bq.y ? epart[Bq.x + k] : (T)0;
        spf[buf][lane] = P.pf[b * kStageF + lane];
        __syncwarp();
        for (int f = 0; f < kStageF; ++f) {
            const int2 e = spf[buf][f];
            if (e.x == -1) break;
            const T *src = e.x == -2 ? der + (int64_t)e.y * R + rr
                                     : reinterpret_cast<const T *>(P.ext_der[e.x]) + e.y + rr * P.ext_size[e.x];
            cp_async_small<sizeof(T)>(&comb[kRing + buf * kStageF + f][lane], src);
        }
        cp_async_commit();
    };

    if (P.n_blocks > 0) load_block(0, 0);
    cp_async_wait_all();
    __syncwarp();
    const char *base = reinterpret_cast<const char *>(&comb[0][lane]);
    T prev = (T)0;
    for (int64_t b = 0; b < P.n_blocks; ++b) {
        const int cur = (int)(b & 1);
        if (b + 1 < P.n_blocks) load_block(b + 1, cur ^ 1);
        auto rd = [&](int c) -> T {
            const unsigned kind = (unsigned)c >> 30, pl = (unsigned)c & 0x3fffffffu;
            if (kind == kSrcSmem) return comb[pl][lane];
            if (kind == kSrcPrev) return prev;
            return der[(int64_t)pl * R + rr];
        };
        int4 r0 = sA[cur][0], r1 = sA[cur][1];
        T p00 = spart[cur][0][0], p01 = spart[cur][0][1], p10 = spart[cur][1][0], p11 = spart[cur][1][1];
        T v00 = *reinterpret_cast<const T *>(base + r0.x), v01 = *reinterpret_cast<const T *>(base + r0.y);
        T *ringrow = &comb[cur * kBlk][lane];
#pragma unroll 4
        for (int k = 0; k < kBlk; ++k) {
            const int4 r2 = sA[cur][k + 2];
            const T p20 = spart[cur][k + 2][0], p21 = spart[cur][k + 2][1];
            const T v10 = *reinterpret_cast<const T *>(base + r1.x), v11 = *reinterpret_cast<const T *>(base + r1.y);
            const int meta = r0.z;
            T x0 = (meta & kM_Prev0) ? prev : v00, x1 = (meta & kM_Prev1) ? prev : v01;
            T acc = (T)0;
            if (meta & kM_Rare) {
                const int4 Bq = sB[cur][k], C = sC[cur][k];
                T *dst = nullptr;
                if (meta & kM_Leaf) {
                    const int4 D = sD[cur][k];
                    const int64_t elem = (int64_t)(unsigned)D.y | ((int64_t)D.z << 32);
                    dst = reinterpret_cast<T *>(P.ext_der[D.x]) + elem + rr * P.ext_size[D.x];      // origin.deriv[index]
                    if (!P.leaves_zero) acc = *dst;                                                  // pull_deriv!
                } else if (meta & kM_Exp) {
                    // adjoints that instructions AFTER the block pushed into an exported slot (one TrackedReal, one deriv field)
                    if (meta & kM_ExpStaged) {
                        const int2 e = spf[cur][Bq.w];
                        acc = acc + comb[kRing + cur * kStageF + Bq.w][lane];
                        if (active) (reinterpret_cast<T *>(P.ext_der[e.x]) + e.y)[r * P.ext_size[e.x]] = (T)0;   // unseed!(output)
                    } else {
                        for (int64_t x = P.xoff[Bq.z]; x < P.xoff[Bq.z + 1]; ++x) {
                            const int o = P.xord[x];
                            T *p = reinterpret_cast<T *>(P.ext_der[o]) + P.xelem[x] + rr * P.ext_size[o];
                            acc = acc + *p;
                            if (active) *p = (T)0;
                        }
                    }
                }
                if (meta & kM_Dir0) x0 = der[(int64_t)C.z * R + rr];
                if (meta & kM_Dir1) x1 = der[(int64_t)C.w * R + rr];
                // increment_deriv!(slot, deriv(out_c) * partial) for every consumer c, in the reference's order
                if (meta & kM_E0) acc = acc + x0 * p00;
                if (meta & kM_E1) acc = acc + x1 * p01;
                const int ne = Bq.y;
                if (ne > 2) {
                    acc = acc + rd(C.x) * spart[cur][k][2];
                    if (ne > 3) acc = acc + rd(C.y) * spart[cur][k][3];
#pragma unroll 4
                    for (int e = kInl; e < ne; ++e) acc = acc + rd(P.tsrc[Bq.x + e]) * epart[Bq.x + e];
                }
                if (dst && active) *dst = acc;                                                       // push_deriv!
            } else {
                const T t0 = x0 * p00, t1 = x1 * p01;
                acc = (meta & kM_E0) ? acc + t0 : acc;
                acc = (meta & kM_E1) ? acc + t1 : acc;
            }
            ringrow[k * 32] = acc;
            if (r0.w >= 0 && active) der[(int64_t)r0.w * R + r] = acc;                                // a far reader will want it
            prev = acc;
            r0 = r1; r1 = r2; p00 = p10; p01 = p11; p10 = p20; p11 = p21; v00 = v10; v01 = v11;
        }
        cp_async_wait_all();
        __syncwarp();
    }
}

// R < 16 (gradient!: R = 1): threads are slots of a level; multi-level launches run on ONE CTA.
template <typename T>
__global__ void __launch_bounds__(kFwdThreads) k_scalar_reverse_slots(const RevP P) {
    T *der = reinterpret_cast<T *>(P.der);
    const T *epart = reinterpret_cast<const T *>(P.epart);
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    const int64_t R = P.R;
    for (int lv = P.lv0; lv < P.lv1; ++lv) {
        const int64_t j0 = P.rlevel_off[lv], j1 = P.rlevel_off[lv + 1];
        for (int64_t j = j0 + tid; j < j1; j += nt) {
            const int s = P.rslot[j];
            const int64_t e0 = P.reoff[j], e1 = P.reoff[j + 1];
            for (int64_t r = 0; r < R; ++r) {
                T acc = (T)0;
                T *dst;
                if (s >= 0) {
                    for (int64_t x = P.xoff[s]; x < P.xoff[s + 1]; ++x) {
                        const int o = P.xord[x];
                        T *p = reinterpret_cast<T *>(P.ext_der[o]) + P.xelem[x] + r * P.ext_size[o];
                        acc = acc + *p; *p = (T)0;
                    }
                    dst = der + (int64_t)s * R + r;
                } else {
                    const int l = ~s, o = P.leaf_ord[l];
                    dst = reinterpret_cast<T *>(P.ext_der[o]) + P.leaf_elem[l] + r * P.ext_size[o];
                    acc = *dst;
                }
                for (int64_t e = e0; e < e1; ++e) acc = acc + der[(int64_t)P.esrc[e] * R + r] * epart[e];
                *dst = acc;
            }
        }
        if (lv + 1 < P.lv1) __syncthreads();
    }
}

// =================================================================================================== bundle ("VLIW") kernels
// One seed has no parallelism but the tape's dataflow, and a scalar tape is full of long dependent chains (acc += term: one level
// per link), so the level-parallel kernels above pay a block barrier plus an L2 round trip per link.  For deep graphs the host
// list-schedules the dataflow into BUNDLES of up to 32 mutually independent slots (critical path first: a chain link goes out
// every bundle while the other lanes work ahead on independent terms) and ONE warp executes bundle after bundle: lanes are the
// slots of a bundle, __syncwarp() is the only synchronisation, values of the last 64 bundles sit in a shared-memory ring, the
// records of bundle t+2 and the global operands of bundle t+1 (leaves, literals, far values, partials) load while bundle t
// computes.  A chain link costs a shared-memory round trip plus its arithmetic.
constexpr int kVRing = 64;            // bundles whose results stay in the ring
constexpr unsigned kVGlobal = 0u, kVRingSrc = 1u, kVLiteral = 2u, kVNone = 3u;
constexpr int kVEmpty = INT_MIN;

// One scalar function evaluation = one ForwardDiff Dual{2} evaluation (scalars.jl:96-116).  Single-operation functions (what
// the DiffRules overloads record: +, -, *, /, ^, sin, ...) take a register-only path with the SAME formulas, in the same order,
// as the bytecode interpreter of expr_eval.cuh applied to unit partials; @forward closures run the interpreter.
template <typename T>
__device__ __forceinline__ void eval_scalar(const int4 *__restrict__ ops, int e_off, int n_ops, int n_args, const int4 op,
                                            const double *__restrict__ fc, const T *x, T &v, T *d) {
    if (n_ops != 1) {
        T h[1];
        eval_expr<T, 2, 1>(ops + e_off, n_ops, fc, x, v, d, h, n_args);
        return;
    }
    const int code = op.x, a = op.y & 1, b = op.z & 1, imm = op.w;
    const T xa = x[a], xb = x[b];
    const T da0 = a == 0 ? (T)1 : (T)0, da1 = a == 1 ? (T)1 : (T)0, db0 = b == 0 ? (T)1 : (T)0, db1 = b == 1 ? (T)1 : (T)0;
    T f0 = (T)0, f1 = (T)0;
    // + - * (unary -) ^2 make up most of any scalar tape and the lanes of a bundle hold all of them at once: evaluate the five
    // together, branch-free, and select (same formulas, same operation order as the cases below), so a bundle of basic
    // operations does not serialise into one pass per operation kind
    if (code == E_ADD || code == E_SUB || code == E_MUL || code == E_NEG || (code == E_POWI && imm == 2)) {
        const T sgn = code == E_SUB ? (T)-1 : (T)1;
        const T v_as = code == E_SUB ? xa - xb : xa + xb, d_as0 = da0 + sgn * db0, d_as1 = da1 + sgn * db1;
        const T v_m = xa * xb, d_m0 = da0 * xb + xa * db0, d_m1 = da1 * xb + xa * db1;
        const T u1 = code == E_NEG ? (T)-1 : (T)2 * xa;                      // f1 of the unary ones
        const T v_u = code == E_NEG ? -xa : xa * xa, d_u0 = u1 * da0, d_u1 = u1 * da1;
        const bool as = code == E_ADD || code == E_SUB, mu = code == E_MUL;
        v = as ? v_as : (mu ? v_m : v_u);
        d[0] = as ? d_as0 : (mu ? d_m0 : d_u0);
        d[1] = as ? d_as1 : (mu ? d_m1 : d_u1);
        return;
    }
    switch (code) {
        case E_CONST: v = (T)fc[imm]; d[0] = d[1] = (T)0; return;
        case E_ARG: v = x[imm & 1]; d[0] = (imm & 1) == 0 ? (T)1 : (T)0; d[1] = (imm & 1) == 1 ? (T)1 : (T)0; return;
        case E_ADD: v = xa + xb; d[0] = da0 + (T)1 * db0; d[1] = da1 + (T)1 * db1; return;
        case E_SUB: v = xa - xb; d[0] = da0 + (T)-1 * db0; d[1] = da1 + (T)-1 * db1; return;
        case E_MUL: v = xa * xb; d[0] = da0 * xb + xa * db0; d[1] = da1 * xb + xa * db1; return;
        case E_DIV: { const T q = xa / xb, ib = (T)1 / xb; v = q; d[0] = (da0 - q * db0) * ib; d[1] = (da1 - q * db1) * ib; } return;
        case E_POW: { const T Da[2] = {da0, da1}, Db[2] = {db0, db1}; pow_dual<T, 2>(xa, xb, imm, Da, Db, v, d); } return;
        case E_MAX: case E_MIN: { const bool pa = code == E_MAX ? (xa > xb) : (xa < xb); v = pa ? xa : xb; d[0] = pa ? da0 : db0; d[1] = pa ? da1 : db1; } return;
        case E_NEG: f0 = -xa; f1 = (T)-1; break;
        case E_POWI: f0 = powi(xa, imm); f1 = imm != 0 ? (T)imm * powi(xa, imm - 1) : (T)0; break;
        case E_TANH: { const T t = M<T>::tanh_(xa); f0 = t; f1 = (T)1 - t * t; } break;
        case E_EXP: { const T e = M<T>::exp_(xa); f0 = e; f1 = e; } break;
        case E_LOG: f0 = M<T>::log_(xa); f1 = (T)1 / xa; break;
        case E_SQRT: { const T q = M<T>::sqrt_(xa); f0 = q; f1 = (T)1 / ((T)2 * q); } break;
        case E_SIN: f0 = M<T>::sin_(xa); f1 = M<T>::cos_(xa); break;
        case E_COS: f0 = M<T>::cos_(xa); f1 = -M<T>::sin_(xa); break;
        default: {                                   // anything else: the interpreter is the definition
            T h[1];
            eval_expr<T, 2, 1>(ops + e_off, n_ops, fc, x, v, d, h, n_args);
            return;
        }
    }
    v = f0; d[0] = f1 * da0; d[1] = f1 * da1;
}

// forward record of (bundle, lane):  A = (instruction or kVEmpty, source a, source b, 0)   sources: kind << 30 | payload -
//   kVGlobal: index into val[] (a leaf, a literal - the pool is appended to val[] - or an instruction output older than the
//   ring), kVRingSrc: ring entry
//   B = (where partial a goes in epart or -1, same for b, bytecode offset, n_ops | n_args << 16)     C = the first bytecode operation
//   D = (where partial a goes in bpart - the bundle-ordered copy read by k_scalar_reverse_bundles - or -1, same for b, 0, 0)
// Software pipeline, one cp.async group per bundle: records travel kVRecAhead bundles ahead into a shared-memory ring of
// kVRecSlots, the global operands of a bundle kVOpAhead bundles ahead (their addresses come from the record, which has landed by
// then); wait_group keeps the last kVOpAhead - 1 groups in flight.
constexpr int kVRecAhead = 6, kVRecSlots = 8, kVOpAhead = 3, kVOpSlots = 4;
struct VFwdP {
    const int4 *recA, *recB, *recC, *recD; const int4 *ops; const double *fconst; void *val; void *epart; void *bpart; int64_t n_leaf; int64_t n_bundles;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_keep() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T>
__global__ void __launch_bounds__(32) k_scalar_forward_bundles(const VFwdP P) {
    __shared__ __align__(16) T ring[kVRing * 32];
    __shared__ int4 srec[kVRecSlots][4][32];
    __shared__ __align__(16) T sop[kVOpSlots][2][32];
    const int lane = threadIdx.x;
    T *val = reinterpret_cast<T *>(P.val);
    T *epart = reinterpret_cast<T *>(P.epart);
    T *bpart = reinterpret_cast<T *>(P.bpart);
    const int64_t T_ = P.n_bundles;
    for (int64_t t = -kVRecAhead; t < T_; ++t) {
        cp_async_wait_keep<kVOpAhead - 1>();                       // groups of iterations <= t - kVOpAhead have landed
        int4 A0 = make_int4(kVEmpty, 0, 0, 0), B0 = A0, C0 = A0, D0 = A0;
        if (t >= 0) { A0 = srec[t & (kVRecSlots - 1)][0][lane]; B0 = srec[t & (kVRecSlots - 1)][1][lane]; C0 = srec[t & (kVRecSlots - 1)][2][lane]; D0 = srec[t & (kVRecSlots - 1)][3][lane]; }
        {                                                          // records of bundle t + kVRecAhead (tables are padded)
            const int64_t tr = t + kVRecAhead, q = tr * 32 + lane;
            const int sl = (int)(tr & (kVRecSlots - 1));
            cp_async16(&srec[sl][0][lane], P.recA + q); cp_async16(&srec[sl][1][lane], P.recB + q); cp_async16(&srec[sl][2][lane], P.recC + q);
            cp_async16(&srec[sl][3][lane], P.recD + q);
        }
        const int64_t to = t + kVOpAhead;
        if (to >= 0 && to < T_) {                                  // pull_value! of bundle t + kVOpAhead: leaves, literals, old slots
            const int4 A = srec[to & (kVRecSlots - 1)][0][lane];
            if (A.x != kVEmpty) {
                if (((unsigned)A.y >> 30) == kVGlobal) cp_async_small<sizeof(T)>(&sop[to & (kVOpSlots - 1)][0][lane], val + (A.y & 0x3fffffff));
                if (((unsigned)A.z >> 30) == kVGlobal) cp_async_small<sizeof(T)>(&sop[to & (kVOpSlots - 1)][1][lane], val + (A.z & 0x3fffffff));
            }
        }
        cp_async_commit();
        if (A0.x != kVEmpty) {
            T x[2], v, d[2];
            const unsigned ka = (unsigned)A0.y >> 30, kb = (unsigned)A0.z >> 30;
            x[0] = ka == kVRingSrc ? ring[A0.y & 0x3fffffff] : (ka == kVGlobal ? sop[t & (kVOpSlots - 1)][0][lane] : (T)0);
            x[1] = kb == kVRingSrc ? ring[A0.z & 0x3fffffff] : (kb == kVGlobal ? sop[t & (kVOpSlots - 1)][1][lane] : (T)0);
            eval_scalar<T>(P.ops, B0.z, B0.w & 0xffff, B0.w >> 16, C0, P.fconst, x, v, d);
            val[P.n_leaf + A0.x] = v;                                       // value!(output, ...)
            ring[(int)(t & (kVRing - 1)) * 32 + lane] = v;
            if (B0.x >= 0) epart[B0.x] = d[0];                              // cache[] = partials, where the reverse sweep reads them
            if (B0.y >= 0) epart[B0.y] = d[1];
            if (D0.x >= 0) bpart[D0.x] = d[0];
            if (D0.y >= 0) bpart[D0.y] = d[1];
        }
        __syncwarp();
    }
    cp_async_wait_all();
}

// reverse record of (bundle, lane):  A = (slot: >= 0 instruction output, < 0 leaf ~l, kVEmpty; source of edge 0; source of edge 1;
//   number of edges | has-exports << 24)   sources: kVGlobal: consumer instruction (adjoint at der[c*R + r]), kVRingSrc: ring entry
//   B = (first edge in vsrc/epart; leaf: ordinal of the origin buffer; leaf: element, low word; high word)
struct VRevP {
    const int4 *recA, *recB; const int32_t *vsrc; const void *epart; const void *bpart;
    const int64_t *xoff; const int32_t *xord; const int64_t *xelem;
    void *const *ext_der; const int64_t *ext_size;
    void *der; int64_t R; int64_t n_bundles; int leaves_zero;
};

// grid.x = seeds: one warp per seed, the seeds share nothing.  Same pipeline as the forward kernel with a shallower ring and
// shorter distances (13 KB of shared memory: 13+ CTAs per SM when many seeds run at once); the staged operands of a bundle are
// its two partials, the adjoints it reads from global memory and a leaf's current origin.deriv[index].  Only adjoints with a
// reader more than kVRingR bundles away get a row in global memory (B.y).
constexpr int kVRingR = 16, kVRecAheadR = 6, kVRecSlotsR = 8, kVOpAheadR = 3, kVOpSlotsR = 4;
template <typename T>
__global__ void __launch_bounds__(32) k_scalar_reverse_bundles(const VRevP P) {
    constexpr int kVRing = kVRingR, kVRecAhead = kVRecAheadR, kVRecSlots = kVRecSlotsR, kVOpAhead = kVOpAheadR, kVOpSlots = kVOpSlotsR;
    __shared__ __align__(16) T ring[kVRing * 32];
    __shared__ int4 srec[kVRecSlots][2][32];
    __shared__ __align__(16) T sop[kVOpSlots][3][32];             // [0], [1]: adjoints read from global memory, [2]: a leaf's origin.deriv
    __shared__ __align__(16) T spp[kVOpSlots][32][2];             // the partials of edges 0 and 1: ONE coalesced copy per lane (bpart is bundle-ordered)
    const int lane = threadIdx.x;
    const int64_t R = P.R, r = blockIdx.x;
    T *der = reinterpret_cast<T *>(P.der);
    const T *epart = reinterpret_cast<const T *>(P.epart);
    const T *bpart = reinterpret_cast<const T *>(P.bpart);
    const int64_t T_ = P.n_bundles;
    auto leaf_ptr = [&](const int4 &Bq) -> T * {
        const int64_t elem = (int64_t)(unsigned)Bq.z | ((int64_t)Bq.w << 32);
        return reinterpret_cast<T *>(P.ext_der[Bq.y]) + elem + r * P.ext_size[Bq.y];              // origin.deriv[index]
    };
    for (int64_t t = -kVRecAhead; t < T_; ++t) {
        cp_async_wait_keep<kVOpAhead - 1>();
        int4 A0 = make_int4(kVEmpty, 0, 0, 0), B0 = A0;
        if (t >= 0) { A0 = srec[t & (kVRecSlots - 1)][0][lane]; B0 = srec[t & (kVRecSlots - 1)][1][lane]; }
        {
            const int64_t tr = t + kVRecAhead, q = tr * 32 + lane;
            const int sl = (int)(tr & (kVRecSlots - 1));
            cp_async16(&srec[sl][0][lane], P.recA + q); cp_async16(&srec[sl][1][lane], P.recB + q);
        }
        const int64_t to = t + kVOpAhead;
        if (to >= 0 && to < T_) {
            const int4 A = srec[to & (kVRecSlots - 1)][0][lane], Bq = srec[to & (kVRecSlots - 1)][1][lane];
            if (A.x != kVEmpty) {
                T(*o)[32] = sop[to & (kVOpSlots - 1)];
                const int ne = A.w & 0xffffff;
                if (ne > 0) {
                    cp_async_small<2 * sizeof(T)>(&spp[to & (kVOpSlots - 1)][lane][0], bpart + 2 * (to * 32 + lane));
                    if (((unsigned)A.y >> 30) == kVGlobal) cp_async_small<sizeof(T)>(&o[0][lane], der + (int64_t)(A.y & 0x3fffffff) * R + r);
                }
                if (ne > 1 && ((unsigned)A.z >> 30) == kVGlobal) cp_async_small<sizeof(T)>(&o[1][lane], der + (int64_t)(A.z & 0x3fffffff) * R + r);
                if (A.x < 0 && !P.leaves_zero) cp_async_small<sizeof(T)>(&o[2][lane], leaf_ptr(Bq));   // pull_deriv!
                if (A.x >= 0 && ((A.w >> 24) & 2))                                                       // what later instructions pushed into an exported slot
                    cp_async_small<sizeof(T)>(&o[2][lane], reinterpret_cast<const T *>(P.ext_der[Bq.w]) + Bq.z + r * P.ext_size[Bq.w]);
            }
        }
        cp_async_commit();
        if (A0.x != kVEmpty) {
            T(*o)[32] = sop[t & (kVOpSlots - 1)];
            const int ne = A0.w & 0xffffff;
            T acc = (A0.x < 0 && !P.leaves_zero) ? o[2][lane] : (T)0;
            const T p0 = spp[t & (kVOpSlots - 1)][lane][0], p1 = spp[t & (kVOpSlots - 1)][lane][1];
            if (A0.x >= 0 && ((A0.w >> 24) & 2)) {
                acc = acc + o[2][lane];
                (reinterpret_cast<T *>(P.ext_der[B0.w]) + B0.z)[r * P.ext_size[B0.w]] = (T)0;               // unseed!(output)
            } else if (A0.x >= 0 && (A0.w >> 24)) {
                // adjoints that instructions AFTER the block pushed into an exported slot (one TrackedReal, one deriv field)
                for (int64_t x = P.xoff[A0.x]; x < P.xoff[A0.x + 1]; ++x) {
                    const int oo = P.xord[x];
                    T *p = reinterpret_cast<T *>(P.ext_der[oo]) + P.xelem[x] + r * P.ext_size[oo];
                    acc = acc + *p; *p = (T)0;                                                      // ... and unseed!(output)
                }
            }
            // increment_deriv!(slot, deriv(out_c) * partial) for every consumer c, in the reference's order
            if (ne > 0) acc = acc + (((unsigned)A0.y >> 30) == kVRingSrc ? ring[A0.y & 0x3fffffff] : o[0][lane]) * p0;
            if (ne > 1) acc = acc + (((unsigned)A0.z >> 30) == kVRingSrc ? ring[A0.z & 0x3fffffff] : o[1][lane]) * p1;
            for (int e = 2; e < ne; ++e) {
                const int c = P.vsrc[B0.x + e];
                const T a = ((unsigned)c >> 30) == kVRingSrc ? ring[c & 0x3fffffff] : der[(int64_t)(c & 0x3fffffff) * R + r];
                acc = acc + a * epart[B0.x + e];
            }
            ring[(int)(t & (kVRing - 1)) * 32 + lane] = acc;
            if (A0.x < 0) *leaf_ptr(B0) = acc;                                                      // push_deriv!
            else if (B0.y >= 0) der[(int64_t)B0.y * R + r] = acc;                                   // a reader beyond the ring wants it
        }
        __syncwarp();
    }
    cp_async_wait_all();
}

// levels -> launches: a wide level alone on a full grid, runs of narrow levels on one CTA (block-level barrier between levels)
std::vector<ScalarSeg> plan_segments(const std::vector<int64_t> &off) {
    std::vector<ScalarSeg> segs;
    const int nl = (int)off.size() - 1;
    int lv = 0;
    while (lv < nl) {
        const int64_t w = off[lv + 1] - off[lv];
        if (w >= kWideLevel) {
            int grid = (int)std::min<int64_t>((w + kFwdThreads - 1) / kFwdThreads, 148 * 4);
            segs.push_back({lv, lv + 1, grid});
            ++lv;
        } else {
            int e = lv;
            while (e < nl && off[e + 1] - off[e] < kWideLevel) ++e;
            segs.push_back({lv, e, 1});
            lv = e;
        }
    }
    return segs;
}

// List scheduling into bundles of <= 32 nodes: a node runs in a LATER bundle than every predecessor; among the ready nodes the
// one with the longest path to a sink goes first (ties: lower id).  succ = CSR successor lists, indeg = predecessor edge counts.
struct BundlePlan { int64_t n_bundles = 0; std::vector<int32_t> bundle, lane; };
BundlePlan schedule_bundles(int64_t N, const std::vector<int64_t> &succ_off, const std::vector<int32_t> &succ, std::vector<int32_t> indeg,
                            const std::vector<int32_t> &height) {
    BundlePlan bp;
    bp.bundle.assign(N, -1); bp.lane.assign(N, -1);
    typedef std::pair<int32_t, int32_t> Key;                     // (height, -id): max-heap
    std::vector<Key> heap, pending;
    for (int64_t v = 0; v < N; ++v) if (indeg[v] == 0) heap.push_back(Key(height[v], -(int32_t)v));
    std::make_heap(heap.begin(), heap.end());
    int64_t done = 0, t = 0;
    while (done < N) {
        if (heap.empty()) break;                                  // cannot happen on a DAG
        int k = 0;
        pending.clear();
        while (k < 32 && !heap.empty()) {
            std::pop_heap(heap.begin(), heap.end());
            const int32_t v = -heap.back().second;
            heap.pop_back();
            bp.bundle[v] = (int32_t)t; bp.lane[v] = k++;
            ++done;
            for (int64_t e = succ_off[v]; e < succ_off[v + 1]; ++e)
                if (--indeg[succ[e]] == 0) pending.push_back(Key(height[succ[e]], -succ[e]));
        }
        for (const Key &x : pending) { heap.push_back(x); std::push_heap(heap.begin(), heap.end()); }
        ++t;
    }
    bp.n_bundles = done == N ? t : -1;
    return bp;
}

template <typename U>
U *carve(uint8_t *&p, size_t count) {
    U *r = reinterpret_cast<U *>(p);
    p += (count * sizeof(U) + 15) & ~size_t(15);
    return r;
}
}  // namespace

std::string scalar_block_build(ScalarBlock &B, const int64_t *w, int64_t len, const int32_t *expr, uint64_t n_expr, uint64_t n_fconst,
                               const std::vector<int64_t> &buf_size, size_t es, cudaStream_t s) {
    if (len < 4) return "scalar block: payload too short";
    const int64_t n = w[0], nref = w[1], nexp = w[2], nex = w[3];
    if (n < 0 || nref < 0 || nexp < 0 || nex < 0 || n > INT_MAX / 4 || nref > 1 << 20) return "scalar block: bad header";
    if (4 + nref + 3 * nex + 3 * nexp + 5 * n != len) return "scalar block: payload length mismatch";
    const int64_t *refs = w + 4, *etab = refs + nref, *xtab = etab + 3 * nex, *body = xtab + 3 * nexp;
    B.n = n; B.n_ref = (int)nref; B.n_exports = nexp;
    B.refbufs.resize(nref); B.ref_has_leaf.assign(nref, 0); B.ref_has_export.assign(nref, 0); B.ref_export_count.assign(nref, 0);
    for (int64_t k = 0; k < nref; ++k) {
        if (refs[k] < 0 || (size_t)refs[k] >= buf_size.size() || buf_size[refs[k]] < 0) return "scalar block: bad referenced buffer";
        B.refbufs[k] = (int32_t)refs[k];
    }
    // scalar functions --------------------------------------------------------------------------------------------------
    std::vector<int4> h_ops, h_exprs(nex);
    for (int64_t k = 0; k < nex; ++k) {
        const int64_t eo = etab[3 * k], el = etab[3 * k + 1], na = etab[3 * k + 2];
        if (eo < 0 || el < 1 || na < 0 || na > 2 || (uint64_t)(eo + 4 * el) > n_expr) return "scalar block: bad scalar function";
        if (na + el > kMaxExprRegs) return "scalar block: scalar function too large";
        h_exprs[k] = make_int4((int)h_ops.size(), (int)el, (int)na, 0);
        for (int64_t o = 0; o < el; ++o) {
            const int32_t *q = expr + eo + 4 * o;
            const int code = q[0], ra = q[1], rb = q[2], imm = q[3], reg = (int)(na + o);
            if (code < E_CONST || code > E_LAST) return "scalar block: scalar function is not on the whitelist";
            if (code == E_CONST) { if (imm < 0 || (uint64_t)imm >= n_fconst) return "scalar block: bad literal"; }
            else if (code == E_ARG) { if (imm < 0 || imm >= na) return "scalar block: bad arg ref"; }
            else {
                if (ra < 0 || ra >= reg) return "scalar block: bad SSA ref";
                const bool binary = code == E_ADD || code == E_SUB || code == E_MUL || code == E_DIV || code == E_POW || code == E_MAX || code == E_MIN;
                if (binary && (rb < 0 || rb >= reg)) return "scalar block: bad SSA ref";
            }
            h_ops.push_back(make_int4(code, ra, rb, imm));
        }
    }
    // operands: leaves (unique (buffer, element) pairs = one TrackedReal-with-origin each), literals, pool slots ---------------
    std::map<std::pair<int, int64_t>, int> leaf_id;
    std::vector<int32_t> leaf_ord; std::vector<int64_t> leaf_elem;
    std::vector<int> code_a(n), code_b(n);        // >= 0: pool slot i; <= -2-...: see below; kNone
    std::vector<int> kind_a(n), kind_b(n);        // 0 none, 1 pool, 2 leaf, 3 literal
    for (int64_t i = 0; i < n; ++i) {
        const int64_t *I = body + 5 * i;
        if (I[0] < 0 || I[0] >= nex) return "scalar block: bad scalar-function id";
        int nargs = 0;
        for (int k = 0; k < 2; ++k) {
            const int64_t sp = I[1 + 2 * k], ix = I[2 + 2 * k];
            int kind, code;
            if (sp == -3) { kind = 0; code = kNone; }
            else if (sp == -1) { if (ix < 0 || ix >= i) return "scalar block: operand refers to a later instruction"; kind = 1; code = (int)ix; ++nargs; }
            else if (sp == -2) { if (ix < 0 || (uint64_t)ix >= n_fconst) return "scalar block: bad literal operand"; kind = 3; code = (int)ix; ++nargs; }
            else if (sp >= 0 && sp < nref) {
                if (ix < 0 || ix >= buf_size[B.refbufs[sp]]) return "scalar block: operand element out of bounds";
                auto key = std::make_pair((int)sp, ix);
                auto it = leaf_id.find(key);
                if (it == leaf_id.end()) {
                    it = leaf_id.emplace(key, (int)leaf_ord.size()).first;
                    leaf_ord.push_back((int32_t)sp); leaf_elem.push_back(ix);
                    B.ref_has_leaf[sp] = 1;
                }
                kind = 2; code = it->second; ++nargs;
            } else return "scalar block: bad operand space";
            (k == 0 ? kind_a : kind_b)[i] = kind;
            (k == 0 ? code_a : code_b)[i] = code;
        }
        if (kind_a[i] == 0 && kind_b[i] != 0) return "scalar block: second operand without a first";
        if (nargs != h_exprs[I[0]].z) return "scalar block: operand count does not match the scalar function";
    }
    const int64_t nl = (int64_t)leaf_ord.size();
    B.n_leaf = nl;
    if (nl + n > INT_MAX / 2) return "scalar block: too many slots";
    // exports ------------------------------------------------------------------------------------------------------------
    std::vector<int64_t> xoff(n + 1, 0);
    for (int64_t k = 0; k < nexp; ++k) {
        const int64_t sl = xtab[3 * k], ro = xtab[3 * k + 1], el = xtab[3 * k + 2];
        if (sl < 0 || sl >= n || ro < 0 || ro >= nref || el < 0 || el >= buf_size[B.refbufs[ro]]) return "scalar block: bad export";
        if (leaf_id.count(std::make_pair((int)ro, el))) return "scalar block: export target is also read as an operand";
        xoff[sl + 1]++;
        B.ref_has_export[ro] = 1; B.ref_export_count[ro]++;
    }
    for (int64_t i = 0; i < n; ++i) xoff[i + 1] += xoff[i];
    std::vector<int32_t> xord(nexp), exp_slot(nexp), exp_ord(nexp); std::vector<int64_t> xelem(nexp), exp_elem(nexp);
    {
        std::vector<int64_t> fill(xoff.begin(), xoff.end() - 1);
        for (int64_t k = 0; k < nexp; ++k) {
            const int64_t sl = xtab[3 * k], at = fill[sl]++;
            xord[at] = (int32_t)xtab[3 * k + 1]; xelem[at] = xtab[3 * k + 2];
            exp_slot[k] = (int32_t)sl; exp_ord[k] = (int32_t)xtab[3 * k + 1]; exp_elem[k] = xtab[3 * k + 2];
        }
    }
    // forward levels --------------------------------------------------------------------------------------------------------
    std::vector<int> lf(n);
    int nfl = 0;
    for (int64_t i = 0; i < n; ++i) {
        int l = 0;
        if (kind_a[i] == 1) l = std::max(l, lf[code_a[i]] + 1);
        if (kind_b[i] == 1) l = std::max(l, lf[code_b[i]] + 1);
        lf[i] = l; nfl = std::max(nfl, l + 1);
    }
    std::vector<int64_t> flevel_off(nfl + 1, 0);
    for (int64_t i = 0; i < n; ++i) flevel_off[lf[i] + 1]++;
    for (int l = 0; l < nfl; ++l) flevel_off[l + 1] += flevel_off[l];
    // reverse (pull) graph: slot ids  [0, n) instruction outputs, [n, n + nl) leaves ---------------------------------------------
    const int64_t ns = n + nl;
    std::vector<int> lr(ns, 0);
    std::vector<int64_t> ecount(ns, 0);
    auto slot_of = [&](int kind, int code) -> int64_t { return kind == 1 ? code : (kind == 2 ? n + code : -1); };
    for (int64_t i = n - 1; i >= 0; --i) {
        for (int k = 0; k < 2; ++k) {
            const int64_t sl = slot_of(k == 0 ? kind_a[i] : kind_b[i], k == 0 ? code_a[i] : code_b[i]);
            if (sl < 0) continue;
            lr[sl] = std::max(lr[sl], lr[i] + 1);
            ecount[sl]++;
        }
    }
    int nrl = 0;
    for (int64_t sl = 0; sl < ns; ++sl) nrl = std::max(nrl, lr[sl] + 1);
    std::vector<int64_t> rlevel_off(nrl + 1, 0);
    for (int64_t sl = 0; sl < ns; ++sl) rlevel_off[lr[sl] + 1]++;
    for (int l = 0; l < nrl; ++l) rlevel_off[l + 1] += rlevel_off[l];
    std::vector<int32_t> rslot(ns);
    std::vector<int64_t> rpos(ns);
    {
        std::vector<int64_t> fill(rlevel_off.begin(), rlevel_off.end() - 1);
        for (int64_t sl = 0; sl < ns; ++sl) {
            const int64_t at = fill[lr[sl]]++;
            rpos[sl] = at;
            rslot[at] = sl < n ? (int32_t)sl : ~(int32_t)(sl - n);
        }
    }
    std::vector<int64_t> reoff(ns + 1, 0);
    for (int64_t sl = 0; sl < ns; ++sl) reoff[rpos[sl] + 1] = ecount[sl];
    for (int64_t j = 0; j < ns; ++j) reoff[j + 1] += reoff[j];
    const int64_t ne = reoff[ns];
    if (ne > INT_MAX) return "scalar block: too many operand edges";
    B.n_edges = ne; B.n_rslots = ns;
    std::vector<int32_t> esrc(ne), epos(2 * n, -1);
    {
        std::vector<int64_t> fill(ns);
        for (int64_t sl = 0; sl < ns; ++sl) fill[sl] = reoff[rpos[sl]];
        for (int64_t i = n - 1; i >= 0; --i) {         // reverse tape order, a before b: the order of the reference's increments
            for (int k = 0; k < 2; ++k) {
                const int64_t sl = slot_of(k == 0 ? kind_a[i] : kind_b[i], k == 0 ? code_a[i] : code_b[i]);
                if (sl < 0) continue;
                const int64_t at = fill[sl]++;
                esrc[at] = (int32_t)i;
                epos[2 * i + k] = (int32_t)at;
            }
        }
    }
    // streaming plan for the many-seed reverse kernel (k_scalar_reverse_stream) -------------------------------------------------
    // schedule: instructions in reverse tape order, every leaf right after its lowest consumer (all its increments are in by then)
    const int64_t nblk = (ns + kBlk - 1) / kBlk, npos = nblk * kBlk;
    std::vector<int4> recA(npos, make_int4(0, 0, 0, -1)), recB(npos, make_int4(0, 0, 0, 0)), recC(npos, make_int4(0, 0, 0, 0)),
        recD(npos, make_int4(0, 0, 0, 0));
    std::vector<int2> pf(nblk * kStageF, make_int2(-1, 0));
    std::vector<int32_t> tsrc(ne, 0);
    B.n_sblocks = nblk; B.n_grow = 0; B.n_near = B.n_staged = B.n_direct = 0;
    {
        std::vector<int64_t> lowc(nl, -1), lhead(n + 1, 0);
        for (int64_t i = 0; i < n; ++i) {
            if (kind_a[i] == 2 && lowc[code_a[i]] < 0) lowc[code_a[i]] = i;
            if (kind_b[i] == 2 && lowc[code_b[i]] < 0) lowc[code_b[i]] = i;
        }
        for (int64_t l = 0; l < nl; ++l) lhead[lowc[l] + 1]++;
        for (int64_t i = 0; i < n; ++i) lhead[i + 1] += lhead[i];
        std::vector<int64_t> lsorted(nl), fill(lhead.begin(), lhead.end() - 1);
        for (int64_t l = 0; l < nl; ++l) lsorted[fill[lowc[l]]++] = l;
        std::vector<int64_t> order;
        order.reserve(ns);
        std::vector<int64_t> qpos(ns);
        for (int64_t i = n - 1; i >= 0; --i) {
            qpos[i] = (int64_t)order.size(); order.push_back(i);
            for (int64_t x = lhead[i]; x < lhead[i + 1]; ++x) { qpos[n + lsorted[x]] = (int64_t)order.size(); order.push_back(n + lsorted[x]); }
        }
        // pass 1: which instruction outputs are read two or more blocks after they are produced -> they need a global row
        std::vector<int32_t> grow(n, -1);
        for (int64_t q = 0; q < ns; ++q) {
            const int64_t sl = order[q], e0 = reoff[rpos[sl]];
            for (int64_t k = 0; k < ecount[sl]; ++k) {
                const int c = esrc[e0 + k];
                if (q / kBlk - qpos[c] / kBlk > 1 && grow[c] < 0) grow[c] = 0;
            }
        }
        for (int64_t i = 0; i < n; ++i) if (grow[i] == 0) grow[i] = (int32_t)B.n_grow++;
        // pass 2: records, operand sources, staging lists
        const int rowbytes = 32 * (int)es;
        std::vector<int> nstage(nblk, 0);
        for (int64_t q = 0; q < ns; ++q) {
            const int64_t sl = order[q], e0 = reoff[rpos[sl]], cnt = ecount[sl], blk = q / kBlk;
            if (cnt > INT_MAX / 2) return "scalar block: a slot has too many consumers";
            int meta = 0, off[2] = {0, 0}, dir[2] = {0, 0}, src23[2] = {0, 0};
            for (int64_t k = 0; k < cnt; ++k) {
                const int c = esrc[e0 + k];
                const int64_t qc = qpos[c];
                int code;
                if (q - qc == 1) { code = src_code(kSrcPrev, 0); B.n_near++; }
                else if (blk - qc / kBlk <= 1) { code = src_code(kSrcSmem, (unsigned)(qc & (kRing - 1))); B.n_near++; }
                else if (k < kInl && nstage[blk] < kStageF) {
                    const int f = nstage[blk]++;
                    pf[blk * kStageF + f] = make_int2(-2, grow[c]);
                    code = src_code(kSrcSmem, (unsigned)(kRing + (blk & 1) * kStageF + f)); B.n_staged++;
                } else { code = src_code(kSrcDirect, (unsigned)grow[c]); B.n_direct++; }
                tsrc[e0 + k] = code;
                const unsigned kind = (unsigned)code >> 30, pl = (unsigned)code & 0x3fffffffu;
                if (k < 2) {
                    meta |= k == 0 ? kM_E0 : kM_E1;
                    if (kind == kSrcPrev) meta |= k == 0 ? kM_Prev0 : kM_Prev1;
                    else if (kind == kSrcSmem) off[k] = (int)pl * rowbytes;
                    else { meta |= (k == 0 ? kM_Dir0 : kM_Dir1) | kM_Rare; dir[k] = (int)pl; }
                } else if (k < kInl) src23[k - 2] = code;
            }
            if (cnt > 2) meta |= kM_Rare;
            int xstage = 0;
            if (sl < n && xoff[sl + 1] > xoff[sl]) {
                meta |= kM_Exp | kM_Rare;
                if (xoff[sl + 1] - xoff[sl] == 1 && xelem[xoff[sl]] < INT_MAX && nstage[blk] < kStageF) {
                    xstage = nstage[blk]++;
                    pf[blk * kStageF + xstage] = make_int2(xord[xoff[sl]], (int)xelem[xoff[sl]]);
                    meta |= kM_ExpStaged;
                }
            }
            if (sl >= n) {
                const int64_t l = sl - n;
                meta |= kM_Leaf | kM_Rare;
                recD[q] = make_int4(leaf_ord[l], (int)(uint32_t)(leaf_elem[l] & 0xffffffffLL), (int)(leaf_elem[l] >> 32), 0);
            }
            recA[q] = make_int4(off[0], off[1], meta, sl < n ? grow[sl] : -1);
            recB[q] = make_int4((int)e0, (int)cnt, sl < n ? (int)sl : 0, xstage);
            recC[q] = make_int4(src23[0], src23[1], dir[0], dir[1]);
        }
        if (getenv("RDB_SCALAR_DEBUG"))
            fprintf(stderr, "[rdb scalar] n=%lld leaves=%lld edges=%lld blocks=%lld global_rows=%lld near=%lld staged=%lld direct=%lld\n", (long long)n,
                    (long long)nl, (long long)ne, (long long)nblk, (long long)B.n_grow, (long long)B.n_near, (long long)B.n_staged, (long long)B.n_direct);
    }
    // bundle plans (k_scalar_forward_bundles / k_scalar_reverse_bundles) for deep graphs ------------------------------------------
    std::vector<int4> vfA, vfB, vfC, vfD, vrA, vrB;
    std::vector<int32_t> bpos(2 * n, -1);         // where partial (i, a|b) goes in the bundle-ordered copy, or -1
    std::vector<int64_t> fq;                       // forward bundle position of instruction i
    int64_t n_bpart = 0;
    std::vector<int32_t> vsrc(ne, 0);
    B.n_fbundles = B.n_rbundles = 0;
    B.fwd_bundles = B.rev_bundles = false;
    {
        static const int force = [] { const char *e = getenv("RDB_SCALAR_BUNDLES"); return e ? atoi(e) : -1; }();   // 0 = never, 1 = always
        const bool deep_f = force >= 0 ? force == 1 : nfl > 48, deep_r = force >= 0 ? force == 1 : nrl > 48;
        if (deep_f && n > 0 && (uint64_t)(nl + n) + n_fconst < (1u << 30)) {
            // forward DAG: instruction -> the later instructions that read its slot
            std::vector<int64_t> soff(n + 1, 0);
            std::vector<int32_t> indeg(n, 0), height(n, 1);
            for (int64_t i = 0; i < n; ++i) {
                if (kind_a[i] == 1) { soff[code_a[i] + 1]++; indeg[i]++; }
                if (kind_b[i] == 1) { soff[code_b[i] + 1]++; indeg[i]++; }
            }
            for (int64_t i = 0; i < n; ++i) soff[i + 1] += soff[i];
            std::vector<int32_t> succ(soff[n]);
            {
                std::vector<int64_t> fill(soff.begin(), soff.end() - 1);
                for (int64_t i = 0; i < n; ++i) {
                    if (kind_a[i] == 1) succ[fill[code_a[i]]++] = (int32_t)i;
                    if (kind_b[i] == 1) succ[fill[code_b[i]]++] = (int32_t)i;
                }
            }
            for (int64_t i = n - 1; i >= 0; --i)
                for (int64_t e = soff[i]; e < soff[i + 1]; ++e) height[i] = std::max(height[i], height[succ[e]] + 1);
            BundlePlan bp = schedule_bundles(n, soff, succ, indeg, height);
            if (bp.n_bundles > 0) {
                const int64_t T = bp.n_bundles;
                fq.assign(n, 0);
                vfA.assign((T + kVRecAhead + 1) * 32, make_int4(kVEmpty, 0, 0, 0)); vfB.assign((T + kVRecAhead + 1) * 32, make_int4(-1, -1, 0, 0)); vfC.assign((T + kVRecAhead + 1) * 32, make_int4(0, 0, 0, 0));
                for (int64_t i = 0; i < n; ++i) {
                    auto src = [&](int kind, int code) -> int {
                        if (kind == 0) return (int)(kVNone << 30);
                        if (kind == 2) return (int)((kVGlobal << 30) | (unsigned)code);                    // leaf: val[code]
                        if (kind == 3) return (int)((kVGlobal << 30) | (unsigned)(nl + n + code));             // literal pool appended to val[]
                        if (bp.bundle[i] - bp.bundle[code] < kVRing)                                        // still in the ring
                            return (int)((kVRingSrc << 30) | (unsigned)((bp.bundle[code] & (kVRing - 1)) * 32 + bp.lane[code]));
                        return (int)((kVGlobal << 30) | (unsigned)(nl + code));
                    };
                    const int64_t q = (int64_t)bp.bundle[i] * 32 + bp.lane[i];
                    const int4 E = h_exprs[body[5 * i]];
                    vfA[q] = make_int4((int)i, src(kind_a[i], code_a[i]), src(kind_b[i], code_b[i]), 0);
                    vfB[q] = make_int4(epos[2 * i], epos[2 * i + 1], E.x, E.y | (E.z << 16));
                    vfC[q] = h_ops[E.x];
                    fq[i] = q;
                }
                B.n_fbundles = T; B.fwd_bundles = true;
            }
        }
        if (deep_r && ns > 0) {
            // pull DAG: consumer instruction c -> its operand slots (they accumulate deriv(out_c))
            std::vector<int64_t> soff(ns + 1, 0);
            std::vector<int32_t> indeg(ns, 0), height(ns, 1);
            for (int64_t i = 0; i < n; ++i)
                for (int k = 0; k < 2; ++k) {
                    const int64_t sl = slot_of(k == 0 ? kind_a[i] : kind_b[i], k == 0 ? code_a[i] : code_b[i]);
                    if (sl >= 0) { soff[i + 1]++; indeg[sl]++; }
                }
            for (int64_t v = 0; v < ns; ++v) soff[v + 1] += soff[v];
            std::vector<int32_t> succ(soff[ns]);
            {
                std::vector<int64_t> fill(soff.begin(), soff.end() - 1);
                for (int64_t i = 0; i < n; ++i)
                    for (int k = 0; k < 2; ++k) {
                        const int64_t sl = slot_of(k == 0 ? kind_a[i] : kind_b[i], k == 0 ? code_a[i] : code_b[i]);
                        if (sl >= 0) succ[fill[i]++] = (int32_t)sl;
                    }
            }
            // operands have lower instruction ids (or are leaves, which have no successors): ascending order is reverse-topological
            for (int64_t i = 0; i < n; ++i)
                for (int64_t e = soff[i]; e < soff[i + 1]; ++e) height[i] = std::max(height[i], height[succ[e]] + 1);
            BundlePlan bp = schedule_bundles(ns, soff, succ, indeg, height);
            if (bp.n_bundles > 0) {
                const int64_t T = bp.n_bundles;
                vrA.assign((T + kVRecAhead + 1) * 32, make_int4(kVEmpty, 0, 0, 0)); vrB.assign((T + kVRecAhead + 1) * 32, make_int4(0, -1, 0, 0));
                std::vector<int32_t> brow(n, -1);                 // global row of an instruction output read from beyond the ring
                B.n_brow = 0;
                for (int64_t sl = 0; sl < ns; ++sl) {
                    const int64_t e0 = reoff[rpos[sl]];
                    for (int64_t k = 0; k < ecount[sl]; ++k) {
                        const int c = esrc[e0 + k];
                        if (bp.bundle[sl] - bp.bundle[c] >= kVRingR && brow[c] < 0) brow[c] = (int32_t)B.n_brow++;
                    }
                }
                for (int64_t sl = 0; sl < ns; ++sl) {
                    const int64_t e0 = reoff[rpos[sl]], cnt = ecount[sl];
                    if (cnt > 0xffffff) { vrA.clear(); break; }
                    int s01[2] = {(int)(kVNone << 30), (int)(kVNone << 30)};
                    for (int64_t k = 0; k < cnt; ++k) {
                        const int c = esrc[e0 + k];
                        const int code = bp.bundle[sl] - bp.bundle[c] < kVRingR
                                             ? (int)((kVRingSrc << 30) | (unsigned)((bp.bundle[c] & (kVRingR - 1)) * 32 + bp.lane[c]))
                                             : (int)((kVGlobal << 30) | (unsigned)brow[c]);
                        vsrc[e0 + k] = code;
                        if (k < 2) s01[k] = code;
                    }
                    const int64_t q = (int64_t)bp.bundle[sl] * 32 + bp.lane[sl];
                    int flags = (sl < n && xoff[sl + 1] > xoff[sl]) ? 1 : 0;
                    const bool one_export = flags && xoff[sl + 1] - xoff[sl] == 1 && xelem[xoff[sl]] < INT_MAX;
                    if (one_export) flags |= 2;                    // its adjoint-so-far is staged like a leaf's (B.z = element, B.w = buffer)
                    vrA[q] = make_int4(sl < n ? (int)sl : ~(int)(sl - n), s01[0], s01[1], (int)cnt | (flags << 24));
                    if (sl < n) vrB[q] = make_int4((int)e0, brow[sl], one_export ? (int)xelem[xoff[sl]] : 0, one_export ? xord[xoff[sl]] : 0);
                    else {
                        const int64_t l = sl - n;
                        vrB[q] = make_int4((int)e0, leaf_ord[l], (int)(uint32_t)(leaf_elem[l] & 0xffffffffLL), (int)(leaf_elem[l] >> 32));
                    }
                }
                if (!vrA.empty()) {
                    B.n_rbundles = T; B.rev_bundles = true;
                    n_bpart = (T + kVRecAhead + 1) * 64;
                    // partial of (instruction i, operand k) = edge e of the operand's slot: edges 0 and 1 of the slot at bundle
                    // position q live at bpart[2q], bpart[2q + 1]
                    for (int64_t i = 0; i < n; ++i)
                        for (int k = 0; k < 2; ++k) {
                            const int64_t sl = slot_of(k == 0 ? kind_a[i] : kind_b[i], k == 0 ? code_a[i] : code_b[i]);
                            if (sl < 0 || epos[2 * i + k] < 0) continue;
                            const int64_t j = epos[2 * i + k] - reoff[rpos[sl]];
                            if (j < 2) bpos[2 * i + k] = (int32_t)(2 * ((int64_t)bp.bundle[sl] * 32 + bp.lane[sl]) + j);
                        }
                }
            }
        }
        if (B.fwd_bundles) {
            vfD.assign(vfA.size(), make_int4(-1, -1, 0, 0));
            for (int64_t i = 0; i < n; ++i) vfD[fq[i]] = make_int4(bpos[2 * i], bpos[2 * i + 1], 0, 0);
        }
        if (getenv("RDB_SCALAR_DEBUG"))
            fprintf(stderr, "[rdb scalar] levels fwd=%d rev=%d bundles fwd=%lld rev=%lld\n", nfl, nrl, (long long)B.n_fbundles, (long long)B.n_rbundles);
    }
    // forward instruction table, level-sorted ------------------------------------------------------------------------------------
    std::vector<int4> finstr(n);
    {
        std::vector<int64_t> fill(flevel_off.begin(), flevel_off.end() - 1);
        auto enc = [&](int kind, int code) -> int {
            if (kind == 0) return kNone;
            if (kind == 1) return (int)(nl + code);     // value slots: leaves first, then instruction outputs
            if (kind == 2) return code;
            return -2 - code;                           // literal: fconst[code]
        };
        for (int64_t i = 0; i < n; ++i)
            finstr[fill[lf[i]]++] = make_int4((int)i, enc(kind_a[i], code_a[i]), enc(kind_b[i], code_b[i]), (int)body[5 * i]);
    }
    B.n_flevels = nfl; B.n_rlevels = nrl;
    B.fsegs = plan_segments(flevel_off);
    B.rsegs = plan_segments(rlevel_off);
    // device tables: one allocation ------------------------------------------------------------------------------------------------
    auto pad = [](size_t b) { return (b + 15) & ~size_t(15); };
    size_t bytes = pad(n * 16) + pad((nfl + 1) * 8) + pad(h_ops.size() * 16) + pad(nex * 16) + pad(2 * n * 4) + pad(nl * 4) + pad(nl * 8) +
                   pad(nexp * 4) * 2 + pad(nexp * 8) + pad(ns * 4) + pad((ns + 1) * 8) + pad(ne * 4) + pad((nrl + 1) * 8) + pad((n + 1) * 8) +
                   pad(nexp * 4) + pad(nexp * 8) + pad(nref * 8) * 3 + pad(npos * 16) * 4 + pad(nblk * kStageF * 8) + pad(ne * 4) * 2 + pad(vfA.size() * 16) * 4 + pad(vrA.size() * 16) * 2 + pad(2 * n * 4) + 1024;
    std::vector<uint8_t> host(bytes, 0);
    if (cudaMalloc(&B.d_tables, bytes) != cudaSuccess) { cudaGetLastError(); return "scalar block: out of device memory (tables)"; }
    uint8_t *hp = host.data();
    auto put = [&](auto *&dptr, const auto &vec) {
        using U = typename std::remove_reference<decltype(vec[0])>::type;
        using V = typename std::remove_const<U>::type;
        uint8_t *before = hp;
        V *dst = carve<V>(hp, vec.size());
        if (!vec.empty()) memcpy(dst, vec.data(), vec.size() * sizeof(V));
        dptr = reinterpret_cast<typename std::remove_reference<decltype(dptr)>::type>((uint8_t *)B.d_tables + (before - host.data()));
    };
    put(B.d_finstr, finstr); put(B.d_flevel_off, flevel_off); put(B.d_ops, h_ops); put(B.d_exprs, h_exprs); put(B.d_epos, epos);
    put(B.d_leaf_ord, leaf_ord); put(B.d_leaf_elem, leaf_elem);
    put(B.d_exp_slot, exp_slot); put(B.d_exp_ord, exp_ord); put(B.d_exp_elem, exp_elem);
    put(B.d_rslot, rslot); put(B.d_reoff, reoff); put(B.d_esrc, esrc); put(B.d_rlevel_off, rlevel_off);
    put(B.d_xoff, xoff); put(B.d_xord, xord); put(B.d_xelem, xelem);
    put(B.d_recA, recA); put(B.d_recB, recB); put(B.d_recC, recC); put(B.d_recD, recD); put(B.d_pf, pf); put(B.d_tsrc, tsrc);
    put(B.d_vfA, vfA); put(B.d_vfB, vfB); put(B.d_vfC, vfC); put(B.d_vfD, vfD); put(B.d_bpos, bpos); put(B.d_vrA, vrA); put(B.d_vrB, vrB); put(B.d_vsrc, vsrc);
    std::vector<void *> nulls(nref, nullptr); std::vector<int64_t> zeros(nref, 0);
    put(B.d_ext_val, nulls); put(B.d_ext_der, nulls); put(B.d_ext_size, zeros);
    if ((size_t)(hp - host.data()) > bytes) return "scalar block: internal table overflow";
    if (cudaMemcpyAsync(B.d_tables, host.data(), bytes, cudaMemcpyHostToDevice, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
        cudaGetLastError(); return "scalar block: table upload failed";
    }
    B.n_lit = B.fwd_bundles ? (int64_t)n_fconst : 0;
    if (cudaMalloc(&B.d_val, std::max<size_t>((size_t)(nl + n + B.n_lit) * es, 16)) != cudaSuccess ||
        cudaMalloc(&B.d_epart, std::max<size_t>((size_t)ne * es, 16)) != cudaSuccess ||
        cudaMalloc(&B.d_bpart, std::max<size_t>((size_t)n_bpart * es, 16)) != cudaSuccess) { cudaGetLastError(); return "scalar block: out of device memory (pools)"; }
    B.h_ext_val.assign(nref, nullptr); B.h_ext_der.assign(nref, nullptr); B.h_ext_size.assign(nref, 0);
    return "";
}

void scalar_block_free(ScalarBlock &B) {
    cudaFree(B.d_tables); cudaFree(B.d_val); cudaFree(B.d_epart); cudaFree(B.d_bpart); cudaFree(B.d_der);
    B.d_tables = B.d_val = B.d_epart = B.d_bpart = B.d_der = nullptr;
    B.der_bytes = 0;
}

cudaError_t scalar_block_upload_ext(ScalarBlock &B, cudaStream_t s) {
    if (B.n_ref == 0) return cudaSuccess;
    cudaError_t e = cudaMemcpyAsync(B.d_ext_val, B.h_ext_val.data(), B.n_ref * sizeof(void *), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(B.d_ext_der, B.h_ext_der.data(), B.n_ref * sizeof(void *), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(B.d_ext_size, B.h_ext_size.data(), B.n_ref * sizeof(int64_t), cudaMemcpyHostToDevice, s);
    return e;
}

// Which kernel takes a reverse sweep of R >= 16 seeds.  Fitted on B200 (profiles/README.md): the streaming kernel costs ~290 clk
// per schedule position whatever R is (one warp per 32 seeds); one bundle warp per seed costs ~(1100 + 100 W) clk per bundle
// with W warps resident on an SM (<= 13), so it wins while the seeds fit a few rounds of a short schedule (n <= 4096 for the
// literal Hessian tape).  Its global adjoint rows must also stay small.
bool scalar_reverse_streams(const ScalarBlock &B, int64_t R, size_t es) {
    if (R < 16) return false;
    if (!B.rev_bundles) return true;
    static const int force = [] { const char *e = getenv("RDB_SCALAR_MANYSEED"); return !e ? 0 : (e[0] == 's' ? 1 : (e[0] == 'b' ? 2 : 0)); }();
    if (force) return force == 1;
    if ((double)B.n_brow * (double)R * (double)es > 4.0 * (1 << 30)) return true;
    const double W = std::min(13.0, std::ceil((double)R / 148.0)), rounds = std::ceil((double)R / (148.0 * 13.0));
    const double t_stream = 290.0 * (double)B.n_sblocks * 32.0, t_bundles = (1100.0 + 100.0 * W) * (double)B.n_rbundles * rounds;
    return t_stream <= t_bundles;
}

cudaError_t scalar_block_ensure_R(ScalarBlock &B, int64_t R, size_t es) {
    // adjoint rows in global memory, [row*R + r]: the level kernel (R < 16, shallow graphs) indexes them by instruction; the
    // bundle and streaming kernels only keep the rows that are read from beyond their shared-memory rings
    const int64_t few = std::min<int64_t>(R, 15);
    size_t elems = 16;
    if (!B.rev_bundles) elems = std::max<size_t>(elems, (size_t)B.n * few);
    else elems = std::max<size_t>(elems, (size_t)B.n_brow * few);
    if (R >= 16) elems = std::max<size_t>(elems, (size_t)(scalar_reverse_streams(B, R, es) ? B.n_grow : B.n_brow) * R);
    const size_t need = elems * es;
    if (need > B.der_bytes) {
        cudaFree(B.d_der);
        B.d_der = nullptr; B.der_bytes = 0;
        cudaError_t e = cudaMalloc(&B.d_der, need);
        if (e != cudaSuccess) return e;
        B.der_bytes = need;
    }
    B.R_alloc = std::max(B.R_alloc, R);
    return cudaSuccess;
}

template <typename T>
cudaError_t scalar_block_forward(cudaStream_t s, ScalarBlock &B, const double *fconst, int *launches) {
    int nl = 0;
    const int64_t n_lit = B.fwd_bundles ? B.n_lit : 0;
    if (B.n_leaf + n_lit > 0) {
        k_scalar_pull<T><<<(unsigned)((B.n_leaf + n_lit + 255) / 256), 256, 0, s>>>((T *)B.d_val, B.d_leaf_ord, B.d_leaf_elem, B.d_ext_val, B.n_leaf,
                                                                                   B.n, fconst, n_lit);
        ++nl;
    }
    if (B.fwd_bundles) {
        VFwdP V{B.d_vfA, B.d_vfB, B.d_vfC, B.d_vfD, B.d_ops, fconst, B.d_val, B.d_epart, B.d_bpart, B.n_leaf, B.n_fbundles};
        k_scalar_forward_bundles<T><<<1, 32, 0, s>>>(V);
        ++nl;
    } else {
        FwdP P{B.d_finstr, B.d_flevel_off, B.d_ops, B.d_exprs, B.d_epos, B.rev_bundles ? B.d_bpos : nullptr, B.d_bpart, fconst, B.d_val, B.d_epart, B.n_leaf, 0, 0};
        for (const ScalarSeg &g : B.fsegs) {
            P.lv0 = g.lv0; P.lv1 = g.lv1;
            k_scalar_forward<T><<<g.grid, kFwdThreads, 0, s>>>(P);
            ++nl;
        }
    }
    if (B.n_exports > 0) {
        k_scalar_export<T><<<(unsigned)((B.n_exports + 255) / 256), 256, 0, s>>>((const T *)B.d_val, B.d_exp_slot, B.d_exp_ord, B.d_exp_elem,
                                                                                 B.d_ext_val, B.n_leaf, B.n_exports);
        ++nl;
    }
    if (launches) *launches = nl;
    return cudaGetLastError();
}

template <typename T>
cudaError_t scalar_block_reverse(cudaStream_t s, ScalarBlock &B, int64_t R, bool leaves_zero, int *launches) {
    int nl = 0;
    const bool stream = scalar_reverse_streams(B, R, sizeof(T));
    if (stream) {
        StreamP P{B.d_recA, B.d_recB, B.d_recC, B.d_recD, B.d_pf, B.d_tsrc, B.d_epart, B.d_xoff, B.d_xord, B.d_xelem,
                  B.d_ext_der, B.d_ext_size, B.d_der, R, B.n_sblocks, leaves_zero ? 1 : 0};
        k_scalar_reverse_stream<T><<<(unsigned)((R + 31) / 32), 32, 0, s>>>(P);
        ++nl;
    } else if (B.rev_bundles) {
        VRevP V{B.d_vrA, B.d_vrB, B.d_vsrc, B.d_epart, B.d_bpart, B.d_xoff, B.d_xord, B.d_xelem, B.d_ext_der, B.d_ext_size, B.d_der, R, B.n_rbundles,
                leaves_zero ? 1 : 0};
        k_scalar_reverse_bundles<T><<<(unsigned)R, 32, 0, s>>>(V);
        ++nl;
    } else {
        RevP P{B.d_rslot, B.d_reoff, B.d_esrc, B.d_epart, B.d_rlevel_off, B.d_xoff, B.d_xord, B.d_xelem, B.d_leaf_ord, B.d_leaf_elem,
               B.d_ext_der, B.d_ext_size, B.d_der, R, 0, B.n_rlevels};
        for (const ScalarSeg &g : B.rsegs) {
            P.lv0 = g.lv0; P.lv1 = g.lv1;
            k_scalar_reverse_slots<T><<<g.grid, kFwdThreads, 0, s>>>(P);
            ++nl;
        }
    }
    if (launches) *launches = nl;
    return cudaGetLastError();
}

// host-only entry for unit tests of the bundle scheduler (no device needed): a DAG as CSR successor lists; returns the number of
// bundles (or -1 on a cycle) and fills bundle[v], lane[v]
int64_t scalar_schedule_bundles_host(int64_t n_nodes, const int64_t *succ_off, const int32_t *succ, int32_t *bundle, int32_t *lane) {
    if (n_nodes < 0 || !succ_off || (!succ && succ_off[n_nodes] > 0) || !bundle || !lane) return -1;
    std::vector<int64_t> off(succ_off, succ_off + n_nodes + 1);
    std::vector<int32_t> su(succ, succ + off[n_nodes]), indeg(n_nodes, 0), height(n_nodes, 1);
    for (int32_t x : su) { if (x < 0 || x >= n_nodes) return -1; indeg[x]++; }
    // heights by reverse topological order (Kahn on the reversed problem is not needed: relax until stable along a topological order)
    std::vector<int32_t> order; order.reserve(n_nodes);
    {
        std::vector<int32_t> deg(indeg);
        for (int64_t v = 0; v < n_nodes; ++v) if (deg[v] == 0) order.push_back((int32_t)v);
        for (size_t i = 0; i < order.size(); ++i)
            for (int64_t e = off[order[i]]; e < off[order[i] + 1]; ++e) if (--deg[su[e]] == 0) order.push_back(su[e]);
        if ((int64_t)order.size() != n_nodes) return -1;
    }
    for (size_t i = order.size(); i-- > 0;) {
        const int32_t v = order[i];
        for (int64_t e = off[v]; e < off[v + 1]; ++e) height[v] = std::max(height[v], height[su[e]] + 1);
    }
    BundlePlan bp = schedule_bundles(n_nodes, off, su, indeg, height);
    if (bp.n_bundles < 0) return -1;
    for (int64_t v = 0; v < n_nodes; ++v) { bundle[v] = bp.bundle[v]; lane[v] = bp.lane[v]; }
    return bp.n_bundles;
}

template cudaError_t scalar_block_forward<double>(cudaStream_t, ScalarBlock &, const double *, int *);
template cudaError_t scalar_block_forward<float>(cudaStream_t, ScalarBlock &, const double *, int *);
template cudaError_t scalar_block_reverse<double>(cudaStream_t, ScalarBlock &, int64_t, bool, int *);
template cudaError_t scalar_block_reverse<float>(cudaStream_t, ScalarBlock &, int64_t, bool, int *);

}  // namespace rdb
