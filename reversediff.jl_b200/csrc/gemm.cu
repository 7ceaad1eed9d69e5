// GEMM dispatch and the non-tcgen05 kernels behind every `*` instruction of a ReverseDiff tape and its two adjoints
// (reference: mul! in src/derivatives/linalg/arithmetic.jl:245-255 forward, :272-285 reverse).
//
//   C(MxN) <- alpha * op(A)(MxK) * op(B)(KxN) + beta * C,   everything column-major (Julia layout).
//
// gemm_f64:  vector operand            -> GEMV kernels below (HBM-bound)
//            m,n,k >= 256 (mode auto)  -> exact int8 slices on the tcgen05 tensor pipe (umma.cu)
//            otherwise / mode = DMMA   -> `mma.sync ... f64` (SASS DMMA.8x8x4, the native FP64 tensor op of sm_100a; tcgen05 has no
//                                         f64 kind): 128x64 CTA tile, 4 warps (64x32 warp tiles = 8x4 DMMA tiles), 2 CTAs/SM,
//                                         3-stage cp.async (LDGSTS.128) ring, padded shared-memory rows (conflict-free fragment
//                                         LDS.64), copy plan hoisted out of the k-loop.  94-96 % of cuBLAS DGEMM.
//            The beta=1 epilogue is what folds `increment_deriv!(a, mul!(a_tmp, ...))` (arithmetic.jl:279-284) into the GEMM:
//            tmp = A*B is rounded to double in the accumulator, then added — same arithmetic.
// gemm_f32:  m,n >= 128, k >= 32 -> 3xTF32 on tcgen05 (umma.cu); otherwise the SIMT FFMA 128x128x16 kernel below.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include "kernels.h"

namespace rdb {

thread_local int g_gemm_launches = 0;

// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = pred ? 16 : 0;   // src-size 0 => zero-fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_8(void *smem, const void *gmem, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_4(void *smem, const void *gmem, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = pred ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ==================================================================================================
// FP64 DMMA kernel
// ==================================================================================================
namespace d64 {
constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256, PAD = 4;
constexpr int STAGES = 4;
// tile of op(X) with R rows in the "outer" dim (m or n) and BK in k:
//   KMAJOR (k contiguous in global):  smem [R][BK+PAD]
//   else   (m/n contiguous in global): smem [BK][R+PAD]
template <bool KMAJOR, int R>
struct Tile {
    static constexpr int STRIDE = KMAJOR ? (BK + PAD) : (R + PAD);
    static constexpr int ELEMS = KMAJOR ? R * (BK + PAD) : BK * (R + PAD);
    __device__ __forceinline__ static int at(int r, int k) { return KMAJOR ? r * STRIDE + k : k * STRIDE + r; }
};

// Per-thread copy plan for one operand tile: global pointers, shared offsets and row predicates are computed ONCE;
// a k-tile then costs one compare + one cp.async + one pointer bump per 16-byte chunk (ncu on the first version showed
// ~150 integer instructions per thread per k-tile in front of the DMMAs, issued right after the barrier).
// `g` points at element (row 0, k 0) of op(X).  KMAJOR: element (r,k) at g[k + r*ld]; else at g[r + k*ld].
template <bool KMAJOR, int R, bool VEC, int NTHREADS>
struct TileLoader {
    using TL = Tile<KMAJOR, R>;
    static constexpr int THREADS = NTHREADS;
    static constexpr int PER = VEC ? (R * BK / 2) / THREADS : (R * BK) / THREADS;
    const double *ptr[PER];
    const double *base;
    int soff[PER];
    int kofs[PER];
    unsigned rowok;
    int64_t kstep;
    __device__ __forceinline__ void init(const double *g, int64_t ld, int r0, int Rtot, int tid) {
        base = g;
        rowok = 0;
        kstep = KMAJOR ? (int64_t)BK : (int64_t)BK * ld;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            int c = tid + i * THREADS, r, k;
            if (VEC) {
                if (KMAJOR) { r = c / (BK / 2); k = (c % (BK / 2)) * 2; } else { k = c / (R / 2); r = (c % (R / 2)) * 2; }
            } else {
                if (KMAJOR) { r = c / BK; k = c % BK; } else { k = c / R; r = c % R; }
            }
            bool ok = r0 + r < Rtot;
            int rr = ok ? r0 + r : 0;
            ptr[i] = KMAJOR ? g + (int64_t)k + (int64_t)rr * ld : g + (int64_t)rr + (int64_t)k * ld;
            soff[i] = TL::at(r, k);
            kofs[i] = k;
            rowok |= (ok ? 1u : 0u) << i;
        }
    }
    // copy k-tile number `kt` (k0 = kt*BK) into stage buffer s, then advance the pointers to the next k-tile
    __device__ __forceinline__ void issue(double *s, int k0, int Ktot) {
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            bool p = ((rowok >> i) & 1u) && (k0 + kofs[i] < Ktot);
            const double *src = p ? ptr[i] : base;
            if (VEC) cp_async_16(s + soff[i], src, p); else cp_async_8(s + soff[i], src, p);
            ptr[i] += kstep;
        }
    }
};

// NWN = warps along N: 4 -> 128x128 tile, 256 threads, 4 stages, 1 CTA/SM; 2 -> 128x64 tile, 128 threads, 3 stages,
// 2 CTAs/SM (two independently progressing CTAs hide each other's barrier and fixed-latency stalls).
template <int NWN> struct Cfg {
    static constexpr int BN_ = 32 * NWN;
    static constexpr int THREADS_ = 64 * NWN;
    static constexpr int STAGES_ = NWN == 4 ? 4 : 3;
    static constexpr int MINB = NWN == 4 ? 1 : 2;
};
template <bool TA, bool TB, int NWN>
constexpr int smem_bytes() {
    return Cfg<NWN>::STAGES_ * (Tile<TA, BM>::ELEMS + Tile<!TB, Cfg<NWN>::BN_>::ELEMS) * (int)sizeof(double);
}

// op(A) is MxK.  !TA: A stored MxK (m contiguous)  -> not KMAJOR.   TA: A stored KxM (k contiguous) -> KMAJOR.
// op(B) is KxN.  !TB: B stored KxN (k contiguous)  -> KMAJOR.       TB: B stored NxK (n contiguous) -> not KMAJOR.
// Grouped tile order (see umma.cu raster()): consecutive CTAs walk 8 m-tiles x all n-tiles so a wave's A panel stays in L2.
__device__ __forceinline__ void raster8(int &m_tile, int &n_tile) {
    const int cid = blockIdx.x + gridDim.x * blockIdx.y, tiles_m = gridDim.x, tiles_n = gridDim.y;
    const int per_group = 8 * tiles_n, group = cid / per_group, first_m = group * 8;
    const int gsz = min(8, tiles_m - first_m), in_group = cid - group * per_group;
    m_tile = first_m + in_group % gsz;
    n_tile = in_group / gsz;
}
template <bool TA, bool TB, bool VEC, int NWN>
__global__ void __launch_bounds__(Cfg<NWN>::THREADS_, Cfg<NWN>::MINB)
dgemm_dmma_kernel(int M, int N, int K, double alpha, const double *__restrict__ A, int64_t lda,
                  const double *__restrict__ B, int64_t ldb, double beta, double *__restrict__ C, int64_t ldc) {
    extern __shared__ __align__(16) double smem[];
    constexpr int BN = Cfg<NWN>::BN_, THREADS = Cfg<NWN>::THREADS_, STAGES = Cfg<NWN>::STAGES_;
    using TLA = Tile<TA, BM>;
    using TLB = Tile<!TB, BN>;
    double *sA = smem;
    double *sB = smem + STAGES * TLA::ELEMS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / NWN, wn = warp % NWN;   // 2 x NWN warps; warp tile 64 x 32
    int m_tile, n_tile;
    raster8(m_tile, n_tile);
    const int m0 = m_tile * BM, n0 = n_tile * BN;
    const int KT = (K + BK - 1) / BK;

    TileLoader<TA, BM, VEC, THREADS> la;
    TileLoader<!TB, BN, VEC, THREADS> lb;
    la.init(A, lda, m0, M, tid);
    lb.init(B, ldb, n0, N, tid);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) {
            la.issue(sA + s * TLA::ELEMS, s * BK, K);
            lb.issue(sB + s * TLB::ELEMS, s * BK, K);
        }
        cp_async_commit();
    }

    // fragment base offsets (loop invariant)
    const int a_off = TLA::at(wm * 64 + g, t);
    const int b_off = TLB::at(wn * 32 + g, t);
    constexpr int A_ROW8 = TA ? 8 * TLA::STRIDE : 8;      // +8 rows of op(A)
    constexpr int B_ROW8 = !TB ? 8 * TLB::STRIDE : 8;
    constexpr int A_K4 = TA ? 4 : 4 * TLA::STRIDE;        // +4 in k
    constexpr int B_K4 = !TB ? 4 : 4 * TLB::STRIDE;

    int stage = 0;
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const double *a_s = sA + stage * TLA::ELEMS + a_off;
        const double *b_s = sB + stage * TLB::ELEMS + b_off;
        double af[8], bf[4];
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = a_s[i * A_ROW8 + kk * A_K4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = b_s[j * B_ROW8 + kk * B_K4];
            if (kk == 1) {
                // refill the stage consumed in the previous iteration; issued behind the first DMMA batch so the tensor
                // pipe already has work queued while the copies are set up
                int nk = kt + STAGES - 1;
                if (nk < KT) {
                    int ns = stage == 0 ? STAGES - 1 : stage - 1;
                    la.issue(sA + ns * TLA::ELEMS, nk * BK, K);
                    lb.issue(sB + ns * TLB::ELEMS, nk * BK, K);
                }
                cp_async_commit();
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        stage = stage + 1 == STAGES ? 0 : stage + 1;
    }
    cp_async_wait<0>();

    // epilogue: C fragment of tile (i,j): rows m0+wm*64+i*8+g, cols n0+wn*32+j*8+2t+{0,1}
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            int col = n0 + wn * 32 + j * 8 + 2 * t + c;
            if (col >= N) continue;
            double *cp = C + (int64_t)col * ldc;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int row = m0 + wm * 64 + i * 8 + g;
                if (row < M) {
                    double v = alpha * acc[i][j][c];
                    if (beta != 0.0) v += beta * cp[row];
                    cp[row] = v;
                }
            }
        }
    }
}

// small-problem variant: 64x64x16 tile, 4 warps (2x2), 32x32 warp tile — more CTAs for the 100x100 config
constexpr int SBM = 64, SBN = 64, STHREADS = 128, SSTAGES = 2;
template <bool KMAJOR, int R>
__device__ __forceinline__ void load_tile_small(double *s, const double *g, int64_t ld, int r0, int k0, int Rtot,
                                                int Ktot, int tid) {
    using TL = Tile<KMAJOR, R>;
    constexpr int ELEMS = R * BK;
#pragma unroll
    for (int i = 0; i < ELEMS / STHREADS; ++i) {
        int c = tid + i * STHREADS;
        int r, k;
        if (KMAJOR) { r = c / BK; k = c % BK; } else { k = c / R; r = c % R; }
        bool p = (r0 + r < Rtot) && (k0 + k < Ktot);
        const double *src = p ? (KMAJOR ? g + (int64_t)(k0 + k) + (int64_t)(r0 + r) * ld
                                        : g + (int64_t)(r0 + r) + (int64_t)(k0 + k) * ld)
                              : g;
        cp_async_8(s + TL::at(r, k), src, p);
    }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(STHREADS)
dgemm_dmma_small_kernel(int M, int N, int K, double alpha, const double *__restrict__ A, int64_t lda,
                        const double *__restrict__ B, int64_t ldb, double beta, double *__restrict__ C, int64_t ldc) {
    using TLA = Tile<TA, SBM>;
    using TLB = Tile<!TB, SBN>;
    __shared__ __align__(16) double sA[SSTAGES * TLA::ELEMS];
    __shared__ __align__(16) double sB[SSTAGES * TLB::ELEMS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;
    const int KT = (K + BK - 1) / BK;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
    for (int s = 0; s < SSTAGES - 1; ++s) {
        if (s < KT) {
            load_tile_small<TA, SBM>(sA + s * TLA::ELEMS, A, lda, m0, s * BK, M, K, tid);
            load_tile_small<!TB, SBN>(sB + s * TLB::ELEMS, B, ldb, n0, s * BK, N, K, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<SSTAGES - 2>();
        __syncthreads();
        {
            int nk = kt + SSTAGES - 1;
            if (nk < KT) {
                int s = nk % SSTAGES;
                load_tile_small<TA, SBM>(sA + s * TLA::ELEMS, A, lda, m0, nk * BK, M, K, tid);
                load_tile_small<!TB, SBN>(sB + s * TLB::ELEMS, B, ldb, n0, nk * BK, N, K, tid);
            }
            cp_async_commit();
        }
        const double *a_s = sA + (kt % SSTAGES) * TLA::ELEMS;
        const double *b_s = sB + (kt % SSTAGES) * TLB::ELEMS;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = a_s[TLA::at(wm * 32 + i * 8 + g, kk + t)];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = b_s[TLB::at(wn * 32 + j * 8 + g, kk + t)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            int col = n0 + wn * 32 + j * 8 + 2 * t + c;
            if (col >= N) continue;
            double *cp = C + (int64_t)col * ldc;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int row = m0 + wm * 32 + i * 8 + g;
                if (row < M) {
                    double v = alpha * acc[i][j][c];
                    if (beta != 0.0) v += beta * cp[row];
                    cp[row] = v;
                }
            }
        }
}

template <bool TA, bool TB, bool VEC, int NWN>
static cudaError_t launch_cfg(cudaStream_t st, int64_t M, int64_t N, int64_t K, double alpha, const double *A, int64_t lda,
                              const double *B, int64_t ldb, double beta, double *C, int64_t ldc) {
    constexpr int smem = smem_bytes<TA, TB, NWN>();
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(dgemm_dmma_kernel<TA, TB, VEC, NWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + Cfg<NWN>::BN_ - 1) / Cfg<NWN>::BN_));
    dgemm_dmma_kernel<TA, TB, VEC, NWN><<<grid, Cfg<NWN>::THREADS_, smem, st>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb,
                                                                               beta, C, ldc);
    return cudaGetLastError();
}

static int gemm_cfg() {
    static int cfg = -1;
    if (cfg < 0) {
        const char *e = getenv("RDB_GEMM_CFG");
        cfg = e ? atoi(e) : 1;
    }
    return cfg;
}

template <bool TA, bool TB>
static cudaError_t launch(cudaStream_t st, int64_t M, int64_t N, int64_t K, double alpha, const double *A, int64_t lda,
                          const double *B, int64_t ldb, double beta, double *C, int64_t ldc) {
    // small problems: more, smaller CTAs
    int64_t big_ctas = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    if (big_ctas < 64) {
        dim3 grid((unsigned)((M + SBM - 1) / SBM), (unsigned)((N + SBN - 1) / SBN));
        dgemm_dmma_small_kernel<TA, TB><<<grid, STHREADS, 0, st>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta,
                                                                    C, ldc);
        return cudaGetLastError();
    }
    // 16-byte cp.async needs even contiguous extents, even leading dims and 16B-aligned bases
    bool vec = (lda % 2 == 0) && (ldb % 2 == 0) && (((uintptr_t)A & 15) == 0) && (((uintptr_t)B & 15) == 0) &&
               ((TA ? K : M) % 2 == 0) && ((TB ? N : K) % 2 == 0);
    if (gemm_cfg() == 0) {
        if (vec) return launch_cfg<TA, TB, true, 4>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
        return launch_cfg<TA, TB, false, 4>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    }
    if (vec) return launch_cfg<TA, TB, true, 2>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    return launch_cfg<TA, TB, false, 2>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}
}  // namespace d64


// ==================================================================================================
// GEMV: `*` with a vector operand (forward sweep of the jacobian! config: u = A*x, v = B*x).  HBM-bound (the matrix is read
// once); a 128x64-tile GEMM kernel with one useful column wasted ~8x that.
//   trans:   y(cols) = alpha * A' * x(rows) + beta*y   one warp per column, lanes stride the contiguous column, shuffle reduce
//   notrans: y(rows) = alpha * A  * x(cols) + beta*y   threads along rows (coalesced), K split over blockIdx.y, deterministic
//                                                      two-pass reduction through a scratch buffer (no atomics)
// ==================================================================================================
namespace gv {
template <typename T>
__global__ void __launch_bounds__(256) gemv_t_kernel(int rows, int cols, double alpha, const T *__restrict__ A, int64_t lda,
                                                     const T *__restrict__ x, int64_t incx, double beta, T *__restrict__ y, int64_t incy) {
    int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= cols) return;
    const T *col = A + (int64_t)j * lda;
    double acc = 0.0;
    for (int i = lane; i < rows; i += 32) acc += (double)col[i] * (double)x[(int64_t)i * incx];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        double v = alpha * acc;
        if (beta != 0.0) v += beta * (double)y[(int64_t)j * incy];
        y[(int64_t)j * incy] = (T)v;
    }
}
template <typename T>
__global__ void __launch_bounds__(256) gemv_n_partial(int rows, int cols, const T *__restrict__ A, int64_t lda, const T *__restrict__ x,
                                                      int64_t incx, double *__restrict__ part, int kchunk) {
    int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= rows) return;
    int k0 = blockIdx.y * kchunk, k1 = min(cols, k0 + kchunk);
    double acc = 0.0;
#pragma unroll 4
    for (int k = k0; k < k1; ++k) acc += (double)A[r + (int64_t)k * lda] * (double)x[(int64_t)k * incx];
    part[(int64_t)blockIdx.y * rows + r] = acc;
}
template <typename T>
__global__ void __launch_bounds__(256) gemv_n_finish(int rows, int chunks, double alpha, const double *__restrict__ part, double beta,
                                                     T *__restrict__ y, int64_t incy) {
    int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= rows) return;
    double acc = 0.0;
    for (int c = 0; c < chunks; ++c) acc += part[(int64_t)c * rows + r];
    double v = alpha * acc;
    if (beta != 0.0) v += beta * (double)y[(int64_t)r * incy];
    y[(int64_t)r * incy] = (T)v;
}
template <typename T>
static cudaError_t gemv(cudaStream_t st, bool trans, int64_t rows, int64_t cols, double alpha, const T *A, int64_t lda, const T *x,
                        int64_t incx, double beta, T *y, int64_t incy) {
    if (trans) {
        gemv_t_kernel<T><<<(unsigned)((cols + 7) / 8), 256, 0, st>>>((int)rows, (int)cols, alpha, A, lda, x, incx, beta, y, incy);
        g_gemm_launches = 1;
        return cudaGetLastError();
    }
    int64_t bx = (rows + 255) / 256;
    int64_t chunks = (148 * 8 + bx - 1) / bx;                 // ~8 CTAs per SM
    if (chunks > (cols + 31) / 32) chunks = (cols + 31) / 32;
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) chunks = 65535;
    int kchunk = (int)((cols + chunks - 1) / chunks);
    chunks = (cols + kchunk - 1) / kchunk;
    size_t need = (size_t)chunks * rows;
    GemmState &G = gemm_state();                    // per-context scratch (kernels.h)
    double *&g_scratch = G.gemv_scratch;
    if (need > G.gemv_scratch_elems) {
        if (g_scratch) cudaFree(g_scratch);
        g_scratch = nullptr; G.gemv_scratch_elems = 0;
        G.alloc_gen++;
        if (cudaMalloc(&g_scratch, need * sizeof(double)) != cudaSuccess) return cudaErrorMemoryAllocation;
        G.gemv_scratch_elems = need;
    }
    gemv_n_partial<T><<<dim3((unsigned)bx, (unsigned)chunks), 256, 0, st>>>((int)rows, (int)cols, A, lda, x, incx, g_scratch, kchunk);
    gemv_n_finish<T><<<(unsigned)bx, 256, 0, st>>>((int)rows, (int)chunks, alpha, g_scratch, beta, y, incy);
    g_gemm_launches = 2;
    return cudaGetLastError();
}
// C(MxN) = alpha*op(A)*op(B) + beta*C with N == 1 or M == 1, expressed as a GEMV.  Returns false if not applicable.
template <typename T>
static bool try_gemv(cudaStream_t st, bool ta, bool tb, int64_t M, int64_t N, int64_t K, double alpha, const T *A, int64_t lda,
                     const T *B, int64_t ldb, double beta, T *C, int64_t ldc, cudaError_t *err) {
    if (K < 1 || (M < 64 && N < 64)) return false;
    if (N == 1 && M > 1) {
        // y(M) = op(A) * b ;  op(B) is Kx1: stored Kx1 (contiguous) or 1xK with leading dim ldb
        const int64_t incb = tb ? ldb : 1;
        *err = ta ? gemv<T>(st, true, K, M, alpha, A, lda, B, incb, beta, C, 1)       // A stored KxM: y = A' b
                  : gemv<T>(st, false, M, K, alpha, A, lda, B, incb, beta, C, 1);     // A stored MxK: y = A b
        return true;
    }
    if (M == 1 && N > 1) {
        // c(N) = op(B)' * a ;  op(A) is 1xK: stored 1xK (stride lda) or Kx1 (contiguous)
        const int64_t inca = ta ? 1 : lda;
        *err = tb ? gemv<T>(st, false, N, K, alpha, B, ldb, A, inca, beta, C, ldc)    // B stored NxK: c = B a
                  : gemv<T>(st, true, K, N, alpha, B, ldb, A, inca, beta, C, ldc);    // B stored KxN: c = B' a
        return true;
    }
    return false;
}
}  // namespace gv

// mode: 0 = auto (exact int8 slices on tcgen05 when the shape is eligible, DMMA otherwise), 1 = DMMA only, 2 = same as auto.
// RDB_F64_GEMM=dmma|ozaki overrides (experiments / A-B timing).
cudaError_t gemm_f64(cudaStream_t st, bool ta, bool tb, int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                     int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc, int mode, int slices) {
    g_gemm_launches = 0;
    const uint64_t tag_a = g_tag_a, tag_b = g_tag_b;      // operand identity hints are valid for exactly one call
    const OperandWindow win_a = g_win_a, win_b = g_win_b;
    g_tag_a = g_tag_b = 0;
    g_win_a = g_win_b = OperandWindow();
    if (M <= 0 || N <= 0) return cudaSuccess;
    if (M > INT32_MAX || N > INT32_MAX || K > INT32_MAX) return cudaErrorInvalidValue;
    g_gemm_launches = 1;
    {
        cudaError_t ge;
        if (gv::try_gemv<double>(st, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, &ge)) return ge;
    }
    static int env_mode = -1;
    if (env_mode < 0) { const char *e = getenv("RDB_F64_GEMM"); env_mode = !e ? 0 : (e[0] == 'd' ? 1 : (e[0] == 'o' ? 2 : 0)); }
    if (env_mode) mode = env_mode;
    if (mode != 1) {
        cudaError_t e = gemm_f64_ozaki(st, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, slices > 0 ? slices : ozaki_default_slices(),
                                       tag_a, tag_b, win_a, win_b);
        if (e != cudaErrorNotSupported) return e;
    }
    if (!ta && !tb) return d64::launch<false, false>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    if (!ta && tb) return d64::launch<false, true>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    if (ta && !tb) return d64::launch<true, false>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    return d64::launch<true, true>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}

// ==================================================================================================
// FP32 SIMT kernel (128x128x16, 8x8 per thread).  Shared tiles are always [k][m] / [k][n].
// ==================================================================================================
namespace s32 {
constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256, PAD = 4, STAGES = 3;
constexpr int STRIDE = BM + PAD;
constexpr int TILE = BK * STRIDE;

// KMAJOR: element (r,k) at g[k + r*ld]  -> 4-byte copies (transposing scatter)
// else:   element (r,k) at g[r + k*ld]  -> 16-byte copies when VEC
template <bool KMAJOR, bool VEC>
__device__ __forceinline__ void load_tile(float *s, const float *g, int64_t ld, int r0, int k0, int Rtot, int Ktot,
                                          int tid) {
    if (!KMAJOR && VEC) {
#pragma unroll
        for (int i = 0; i < (BM * BK / 4) / THREADS; ++i) {
            int c = tid + i * THREADS;
            int k = c / (BM / 4), rc = (c % (BM / 4)) * 4;
            bool p = (r0 + rc < Rtot) && (k0 + k < Ktot);
            const float *src = p ? g + (int64_t)(r0 + rc) + (int64_t)(k0 + k) * ld : g;
            cp_async_16(s + k * STRIDE + rc, src, p);
        }
    } else {
#pragma unroll
        for (int i = 0; i < (BM * BK) / THREADS; ++i) {
            int c = tid + i * THREADS;
            int r, k;
            if (KMAJOR) { r = c / BK; k = c % BK; } else { k = c / BM; r = c % BM; }
            bool p = (r0 + r < Rtot) && (k0 + k < Ktot);
            const float *src = p ? (KMAJOR ? g + (int64_t)(k0 + k) + (int64_t)(r0 + r) * ld
                                           : g + (int64_t)(r0 + r) + (int64_t)(k0 + k) * ld)
                                 : g;
            cp_async_4(s + k * STRIDE + r, src, p);
        }
    }
}

template <bool TA, bool TB, bool VEC>
__global__ void __launch_bounds__(THREADS, 2)
sgemm_simt_kernel(int M, int N, int K, float alpha, const float *__restrict__ A, int64_t lda,
                  const float *__restrict__ B, int64_t ldb, float beta, float *__restrict__ C, int64_t ldc) {
    extern __shared__ __align__(16) float smemf[];
    float *sA = smemf, *sB = smemf + STAGES * TILE;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;   // thread owns rows tx*4..+3 and 64+tx*4..+3 ; cols ty*4..+3 and 64+ty*4..
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int KT = (K + BK - 1) / BK;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) {
            load_tile<TA, VEC>(sA + s * TILE, A, lda, m0, s * BK, M, K, tid);
            load_tile<!TB, VEC>(sB + s * TILE, B, ldb, n0, s * BK, N, K, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT) {
                int s = nk % STAGES;
                load_tile<TA, VEC>(sA + s * TILE, A, lda, m0, nk * BK, M, K, tid);
                load_tile<!TB, VEC>(sB + s * TILE, B, ldb, n0, nk * BK, N, K, tid);
            }
            cp_async_commit();
        }
        const float *a_s = sA + (kt % STAGES) * TILE;
        const float *b_s = sB + (kt % STAGES) * TILE;
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4 *>(a_s + k * STRIDE + tx * 4);
            float4 a1 = *reinterpret_cast<const float4 *>(a_s + k * STRIDE + 64 + tx * 4);
            float4 b0 = *reinterpret_cast<const float4 *>(b_s + k * STRIDE + ty * 4);
            float4 b1 = *reinterpret_cast<const float4 *>(b_s + k * STRIDE + 64 + ty * 4);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int col = n0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
        if (col >= N) continue;
        float *cp = C + (int64_t)col * ldc;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int row = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
            if (row < M) {
                float v = alpha * acc[i][j];
                if (beta != 0.f) v += beta * cp[row];
                cp[row] = v;
            }
        }
    }
}

template <bool TA, bool TB>
static cudaError_t launch(cudaStream_t st, int64_t M, int64_t N, int64_t K, float alpha, const float *A, int64_t lda,
                          const float *B, int64_t ldb, float beta, float *C, int64_t ldc) {
    bool vec = (lda % 4 == 0) && (ldb % 4 == 0) && (((uintptr_t)A & 15) == 0) && (((uintptr_t)B & 15) == 0) &&
               (M % 4 == 0) && (N % 4 == 0);
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
    constexpr int smem = 2 * STAGES * TILE * (int)sizeof(float);
    cudaError_t e;
    if (vec) {
        static bool done = false;
        if (!done) {
            e = cudaFuncSetAttribute(sgemm_simt_kernel<TA, TB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            done = true;
        }
        sgemm_simt_kernel<TA, TB, true><<<grid, THREADS, smem, st>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta,
                                                                      C, ldc);
    } else {
        static bool done = false;
        if (!done) {
            e = cudaFuncSetAttribute(sgemm_simt_kernel<TA, TB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            done = true;
        }
        sgemm_simt_kernel<TA, TB, false><<<grid, THREADS, smem, st>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb,
                                                                       beta, C, ldc);
    }
    return cudaGetLastError();
}
}  // namespace s32

cudaError_t gemm_f32(cudaStream_t st, bool ta, bool tb, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                     int64_t lda, const float *B, int64_t ldb, float beta, float *C, int64_t ldc) {
    g_gemm_launches = 0;
    g_tag_a = g_tag_b = 0;
    g_win_a = g_win_b = OperandWindow();
    if (M <= 0 || N <= 0) return cudaSuccess;
    if (M > INT32_MAX || N > INT32_MAX || K > INT32_MAX) return cudaErrorInvalidValue;
    g_gemm_launches = 1;
    {
        cudaError_t ge;
        if (gv::try_gemv<float>(st, ta, tb, M, N, K, (double)alpha, A, lda, B, ldb, (double)beta, C, ldc, &ge)) return ge;
    }
    static int force_simt = -1;
    if (force_simt < 0) { const char *e = getenv("RDB_F32_GEMM"); force_simt = (e && e[0] == 's') ? 1 : 0; }
    if (!force_simt) {
        cudaError_t e = gemm_f32_umma(st, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);   // 3xTF32 on tcgen05
        if (e != cudaErrorNotSupported) return e;
    }
    if (!ta && !tb) return s32::launch<false, false>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    if (!ta && tb) return s32::launch<false, true>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    if (ta && !tb) return s32::launch<true, false>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    return s32::launch<true, true>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}

}  // namespace rdb
