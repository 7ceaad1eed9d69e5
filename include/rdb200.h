/*
 * rdb200.h — C ABI of librdb200.so, the B200 (sm_100a) replay engine for ReverseDiff compiled tapes.
 *
 * The reference (JuliaDiff/ReverseDiff.jl v1.16.2, pure Julia) has no FFI; its seam for the tape-replay
 * path is multiple dispatch on the tape object.  Each entry point below replaces one such method; the
 * Julia shim (julia/ReverseDiffB200.jl) binds them with `ccall`.  All citations are
 * relative to the reference tree.
 *
 * Conventions: every function returning `int` returns 0 on success and a negative RDB_E* code on
 * failure; the message is available from rdb_last_error() (thread-local).  No exceptions cross the ABI.
 * Arrays are column-major and contiguous (the reference requires IndexLinear storage, src/tracked.jl:77).
 * The caller owns host arrays; the engine owns all device memory.  A tape handle is NOT thread-safe (the
 * reference mutates tapes in place, src/tracked.jl:155-173); calls are synchronous on return unless
 * stated otherwise.  There is no CPU fallback: without a CUDA device every compute entry point fails
 * with RDB_ENODEVICE.
 */
#ifndef RDB200_H
#define RDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDB_OK 0
#define RDB_EINVAL (-1)       /* malformed program / bad argument                                 */
#define RDB_EUNSUPPORTED (-2) /* instruction kind that cannot be lowered (e.g. @grad closures)     */
#define RDB_ECUDA (-3)        /* CUDA runtime error                                                */
#define RDB_ENODEVICE (-4)    /* no CUDA device: the engine has no CPU path                        */
#define RDB_ENOMEM (-5)

#define RDB_F64 0
#define RDB_F32 1

typedef struct rdb_ctx rdb_ctx;
typedef struct rdb_tape rdb_tape;
typedef struct rdb_comm rdb_comm;

typedef struct rdb_options {
    int32_t fuse_elementwise; /* 1 = fuse runs of adjacent elementwise instructions (default 1)     */
    int32_t seed_block;       /* R: seed rows / Hessian columns processed per batched sweep (0=auto) */
    int32_t use_graph;        /* 1 = replay gradient! through a captured CUDA graph (default 1)      */
    int32_t profile;          /* 1 = record per-step CUDA-event timings for rdb_tape_stats           */
    int32_t gemm_f64_mode;    /* FP64 `*` kernels: 0 = auto (exact int8 slices on tcgen05 when eligible and certified
                                 accurate, DMMA otherwise - see rdb_tape_guard_stats), 1 = DMMA only,
                                 2 = int8 slices without the DMMA replay (measurement)                 */
    int32_t ozaki_slices;     /* int8 slices per FP64 operand for the tcgen05 path (0 = default: 7 of 8 bits) */
    int32_t reserved[2];
} rdb_options;

/* ---- library / context --------------------------------------------------------------------------- */
const char *rdb_version(void);
const char *rdb_last_error(void);
/* One context per device and stream; all state of the kernels (workspaces, sliced-operand cache) belongs to the context, so
 * several contexts - on one device or on several, from one host thread each - can be live in a process. */
rdb_ctx *rdb_ctx_create(int device_id);
void rdb_ctx_destroy(rdb_ctx *);
/* Adopt a caller stream (e.g. torch's current stream) instead of the context's own. */
int rdb_ctx_set_stream(rdb_ctx *, void *cuda_stream);
int rdb_ctx_synchronize(rdb_ctx *);
/* FP64 kernel selection for rdb_gemm on this context (tapes carry their own in rdb_options): mode 0 = auto,
 * 1 = DMMA only; slices = int8 slices per operand of the tcgen05 int8 path (0 = default: 7 slices of 8 bits, at most 8). */
int rdb_ctx_set_gemm_f64_mode(rdb_ctx *, int mode, int slices);

/* Accuracy certificate of the int8-slice FP64 GEMM (the default `*` kernel for m, n, k >= 256).  Its error is norm-wise
 * (relative to the row / column maxima of the operands), not component-wise like DGEMM's, so every launch records
 * max |alpha op(A) op(B)| and a rigorous bound of its own error; a launch is FLAGGED when bound > 2^-40 * max |product|
 * (9.1e-13, inside the 1e-12 contract).  Tapes act on it themselves (rdb_tape_guard_stats); for rdb_gemm the caller reads the
 * counters here: launches checked / flagged since the last call and the worst bound/|product| ratio.  Synchronises. */
int rdb_ctx_gemm_guard(rdb_ctx *, int64_t *checked, int64_t *flagged, double *worst_ratio);

/* ---- tape life cycle ----------------------------------------------------------------------------- */
/* Replaces ReverseDiff.compile(t::AbstractTape) (src/api/tape.jl:146-149) and the CompiledTape it builds
 * (:55-59,101-106).  `program` is the flat tape program of reversediff.jl_b200/program.py (= Appendix C
 * of SURVEY.md); it is validated, CONSTs and index maps are copied to the device.  Returns NULL on
 * failure.  Unsupported instruction kinds make the load fail (no silent fallback). */
rdb_tape *rdb_tape_load(rdb_ctx *, const void *program, size_t nbytes, const rdb_options *opts);
void rdb_tape_destroy(rdb_tape *);

/* value!(input_hook(tape)[slot], x) (src/api/tape.jl:40-44 -> src/tracked.jl:156-163).  The copy is only ENQUEUED (large host
 * matrices upload in column panels on a separate stream and the sweep consumes them panel by panel): `data` must stay alive and
 * unmodified until the next rdb_forward / rdb_gradient / rdb_jacobian / rdb_hessian on this tape has returned. */
int rdb_tape_set_input(rdb_tape *, int input_slot, const void *data, int is_device);

/* forward_pass!(::CompiledTape) (src/api/tape.jl:122-127; src/tape.jl:75-80). */
int rdb_forward(rdb_tape *);
/* seeded_reverse_pass! for a scalar output minus result extraction: pull_value!(output);
 * unseed!(input); seed!(output); reverse_pass!(tape) (src/api/utils.jl:27-33; src/api/tape.jl:129-134). */
int rdb_reverse_scalar_seed(rdb_tape *);

/* gradient!(result, tape::CompiledGradient, input) (src/api/gradients.jl:78-82) after set_input:
 * forward sweep + seeded reverse sweep + extract_result! (src/api/utils.jl:77-100).
 * result_ptrs[i] receives deriv(input[i]) (NULL = skip); value_out (host, element type of the tape, may be
 * NULL) receives value(output) as the DiffResult methods do (src/api/utils.jl:96-100). */
int rdb_gradient(rdb_tape *, void *const *result_ptrs, int results_on_device, void *value_out);

/* The same call, only ENQUEUED: it returns as soon as the sweeps and the copies into `result_ptrs` / `value_out` are on the
 * context's streams (host arrays should be page-locked, or the copies block the caller).  rdb_tape_wait(tape) completes it: it waits
 * for the device, runs the accuracy check of the call's int8-slice GEMMs and, if that fails, replays the call with the DMMA kernels
 * before returning - so results are valid only after rdb_tape_wait.  Between the two calls the tape, its inputs (rdb_tape_set_input
 * arrays) and its results must not be touched, and the tape's context must not run other tapes.  Two tapes of the same program on
 * two contexts pipeline consecutive evaluations: the upload of evaluation k+1 and the download of k-1 overlap the sweeps of k
 * (gradient! at 4096 x 4096 moves 268 MB each way per evaluation, more than half of its compute time over PCIe). */
int rdb_gradient_async(rdb_tape *, void *const *result_ptrs, int results_on_device, void *value_out);
int rdb_tape_wait(rdb_tape *);

/* jacobian!(result, tape::CompiledJacobian, input) (src/api/jacobians.jl:120-124) restricted to the seed
 * rows [row_begin,row_end) of seeded_reverse_pass! (src/api/utils.jl:44-58): one forward sweep, then the
 * rows are replayed in blocks of `seed_block` one-hot seeds.  `result` addresses element (row_begin, 0) of
 * a column-major (rows x length(input[slot])) matrix with leading dimension `ld` (ld = length(output) for
 * the full matrix, src/api/utils.jl:45).  value_out (may be NULL; host or device like `result`) receives
 * value(output) (extract_result_value!, src/api/utils.jl:60-64). */
int rdb_jacobian(rdb_tape *, int input_slot, int64_t row_begin, int64_t row_end, void *result, int64_t ld,
                 int result_on_device, void *value_out);

/* hessian!(result, tape::CompiledHessian, input) (src/api/hessians.jl:86-90) restricted to columns
 * [col_begin,col_end).  Mechanism differs from the reference on purpose (BASELINE.json north_star): the
 * reference replays a scalar reverse-over-reverse tape n times (src/api/tape.jl:285-293); here the inner
 * array tape of f is replayed forward-over-reverse with `seed_block` tangent directions per sweep.
 * `result` addresses element (0, col_begin) of the column-major n x n matrix, leading dimension `ld`.
 * grad_out / value_out (may be NULL) receive the gradient and primal as the DiffResult method does
 * (src/api/hessians.jl:92-97); grad_out may be a host or a device pointer, value_out is a host pointer. */
int rdb_hessian(rdb_tape *, int64_t col_begin, int64_t col_end, void *result, int64_t ld, int result_on_device,
                void *grad_out, void *value_out);

/* Debug / parity accessors: value(x), deriv(x) of any tape buffer, copied to host. */
int rdb_get_value(rdb_tape *, int buffer_id, void *dst_host);
int rdb_get_deriv(rdb_tape *, int buffer_id, void *dst_host);
int64_t rdb_buffer_size(rdb_tape *, int buffer_id);
/* Per-step report (kernel name, time, algorithmic bytes / flops) as JSON; needs opts.profile = 1. */
int rdb_tape_stats(rdb_tape *, char *json, size_t cap);
/* on != 0: rdb_hessian calls whose result is a device array and that return nothing to the host (no grad_out / value_out) are
 * only ENQUEUED on the context's stream; the caller orders later work after them on that stream or calls rdb_ctx_synchronize.
 * Inputs passed by device pointer must stay alive until then.  (A sharded hessian! call is tens of microseconds of device work;
 * returning without a host round trip lets consecutive calls overlap their launch latency.) */
int rdb_tape_set_async(rdb_tape *, int on);
/* Totals since load of the int8-slice GEMM's accuracy certificate on this tape (see rdb_ctx_gemm_guard): launches checked,
 * launches flagged, and how many API calls were therefore replayed with the DMMA kernels before returning (gemm_f64_mode 0;
 * mul!, src/derivatives/linalg/arithmetic.jl:245-285, is DGEMM in the reference).  Any pointer may be NULL. */
int rdb_tape_guard_stats(rdb_tape *, int64_t *checked, int64_t *flagged, int64_t *fallbacks, double *worst_ratio);
/* int8 slice products the tcgen05 `*` kernels of this tape issued since load, summed over launches (folded in with the certificate
 * after every call).  A dense FP64 operand pair costs S(S+1)/2 = 28 of them per launch; operands whose low digit planes are
 * identically zero (constant fills such as the adjoint of `sum`, one-hot seeds, Float32 data held in Float64) cost fewer - the
 * kernels skip those planes, with bit-identical results.  launches checked (rdb_tape_guard_stats) is the number of launches. */
int64_t rdb_tape_slice_pairs(rdb_tape *);
/* Number of kernel launches issued by the engine on this tape since load (bench "gpu_launches"). */
int64_t rdb_tape_launch_count(rdb_tape *);

/* ---- multi-GPU result assembly (one process per GPU) -------------------------------- */
/* jacobian! shards by output-seed rows and hessian! by columns; every rank replays its own tape copy and calls
 * rdb_jacobian(row_begin, row_end) / rdb_hessian(col_begin, col_end) into a device block of its own - no collective on the data
 * path.  The only exchange is the gather of those blocks into the reference's result: the column-major
 * length(output) x length(input) matrix `result_matrix` of seeded_reverse_pass! (src/api/utils.jl:44-58), resp. the n x n Hessian
 * (src/api/hessians.jl:86-90).  Blocks travel over NCCL (NVLink / NVSwitch), reached through dlopen("libnccl.so.2").
 *
 * rdb_comm_unique_id: rank 0 obtains RDB_COMM_ID_BYTES bytes (ncclGetUniqueId) and the HOST distributes them to the other ranks
 * (MPI.jl, torch.distributed, a shared file - the engine has no bootstrap of its own); then every rank calls rdb_comm_create. */
#define RDB_COMM_ID_BYTES 128
int rdb_comm_unique_id(void *id);
rdb_comm *rdb_comm_create(rdb_ctx *, int world, int rank, const void *id);
void rdb_comm_destroy(rdb_comm *);
/* kind 0: rank r holds rows [bounds[r], bounds[r+1]) of a column-major (bounds[world] x other_dim) matrix as a contiguous
 *         (rows_r x other_dim) device block (what rdb_jacobian writes with ld = rows_r); result has leading dimension bounds[world].
 * kind 1: rank r holds columns [bounds[r], bounds[r+1]) of a column-major (other_dim x bounds[world]) matrix (what rdb_hessian
 *         writes with ld = other_dim).
 * bounds (world + 1 entries, the same on every rank) need not be even.  root = -1: every rank receives the full matrix in `result`
 * (device memory); root >= 0: only that rank does (`result` may be NULL elsewhere).  Collective; synchronous on return. */
int rdb_gather_result(rdb_comm *, int dtype, int kind, const void *block, const int64_t *bounds, int64_t other_dim, void *result,
                      int root);
/* The same placement for ONE block in HOST memory (no device, no communicator): for callers that moved the blocks with a
 * transport of their own, and for the CPU test of the N > 1 host logic.  result is the full m x n column-major matrix. */
int rdb_place_block(int dtype, int kind, const void *block, int64_t begin, int64_t end, int64_t m, int64_t n, void *result);

/* ---- single-kernel entry points (kernel tests and per-kernel roofline measurement) ----------------- */
/* C <- alpha*op(A)*op(B) + beta*C, column-major device pointers; the kernel behind every `*` instruction
 * (mul!, src/derivatives/linalg/arithmetic.jl:245-255,272-285). Asynchronous on the context stream. */
int rdb_gemm(rdb_ctx *, int dtype, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha,
             const void *A, int64_t lda, const void *B, int64_t ldb, double beta, void *C, int64_t ldc);
/* out = a + sign*b and block-reduced sum(out) (array +/- fused with sum, arithmetic.jl:55-61 +
 * reductions.jl:23-27); sum_out is a device scalar. */
int rdb_add_sum(rdb_ctx *, int dtype, int64_t n, const void *a, const void *b, double sign, void *out, void *sum_out);

/* ---- host-only planning helper (no device needed; unit-tested on the CPU) --------------------------- */
/* Column panels used for large host arrays (panelled uploads in rdb_tape_set_input, panelled last accumulations with early D2H
 * copies in rdb_gradient): widths[0..3] = columns per panel (0 = unused), chosen so each panel's 128x64 GEMM tiles fill whole
 * waves of the 148 SMs.  Returns the number of panels. */
int rdb_plan_col_panels(int64_t rows, int64_t cols, int64_t widths[4]);
/* The list scheduler behind the scalar-block bundle kernels (reversediff.jl_b200/csrc/scalar.cu): a DAG given as CSR successor
 * lists is packed into bundles of <= 32 mutually independent nodes, every node in a later bundle than all its predecessors, longest
 * path to a sink first.  Fills bundle[v] and lane[v]; returns the number of bundles, -1 for a cyclic or malformed graph. */
int64_t rdb_schedule_bundles(int64_t n_nodes, const int64_t *succ_off, const int32_t *succ, int32_t *bundle, int32_t *lane);

#ifdef __cplusplus
}
#endif
#endif /* RDB200_H */
