// TMA-fed tcgen05 GEMM (sm_100a): the tensor-core path behind `*` instructions and their adjoints
// (reference: mul! in src/derivatives/linalg/arithmetic.jl:245-255,272-285).
//
// One generic kernel: D(128 x 256 tile, TMEM accumulator) = sum over a list of K-segments of A_seg * B_seg^T, where every
// operand slice is a K-major [rows][Kp] matrix reached through a 3-D TMA tensor map (k, row, slice).
//   * warp 0 : TMA producer  (cp.async.bulk.tensor.3d, 128B swizzle, mbarrier complete_tx)
//   * warp 1 : TMEM allocator + single-thread tcgen05.mma issuer (smem descriptors, tcgen05.commit -> mbarrier)
//   * warps 2-5 : epilogue (tcgen05.ld 32x32b -> registers -> alpha/beta -> coalesced column-major stores)
// Two instantiations:
//   kind::tf32  FP32 tapes.  1xTF32 (10-bit mantissa) cannot hold the 1e-5 bound, so operands are split hi = RN_tf32(a),
//               lo = a - hi and three segments (lo*hi, hi*lo, hi*hi) accumulate in one FP32 TMEM accumulator ("3xTF32").
//   kind::i8    FP64 tapes, Ozaki-style exact slicing: each FP64 operand row is scaled by a power of two and cut into
//               signed 7-bit slices; slice products accumulate EXACTLY in int32 TMEM accumulators; the epilogue rescales
//               into the FP64 result.  tcgen05 has no f64 kind — this is how the FP64 GEMMs reach the tcgen05 pipe.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kernels.h"

namespace rdb {
thread_local uint64_t g_tag_a = 0, g_tag_b = 0;
thread_local OperandWindow g_win_a, g_win_b;
thread_local GemmState *g_gemm = nullptr;
GemmState &gemm_state() {
    static GemmState fallback;               // only reached by code that runs outside any context
    return g_gemm ? *g_gemm : fallback;
}
void GemmState::release() {
    if (ws) cudaFree(ws);
    if (gemv_scratch) cudaFree(gemv_scratch);
    if (side_e) cudaFree(side_e);
    for (auto &en : cache) { if (en.q) cudaFree(en.q); if (en.ready) cudaEventDestroy(en.ready); }
    cache.clear();
    if (guard) cudaFree(guard);
    if (guard_sum) cudaFree(guard_sum);
    if (guard_host) cudaFreeHost(guard_host);
    ws = nullptr; gemv_scratch = nullptr; side_e = nullptr; guard = nullptr; guard_sum = nullptr; guard_host = nullptr;
    ws_bytes = gemv_scratch_elems = 0; side_e_n = 0; cache_bytes = 0; guard_used = 0;
}
namespace umma {

constexpr int BM = 128, BN = 256, BKB = 128;          // tile rows / cols / K-bytes per stage (= one 128B swizzle atom)
constexpr int STAGES = 4;
constexpr int A_STAGE = BM * BKB, B_STAGE = BN * BKB;  // 16 KB + 32 KB
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int MAX_SEG = 8;
constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) + 1024 /*align slack*/ + 256 /*barriers*/;

enum Kind { KIND_TF32 = 0, KIND_I8 = 1 };

struct Segs {
    int n;
    int sa[MAX_SEG], sb[MAX_SEG];
};

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int KIND>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == KIND_TF32) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
            "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
            "r"(accumulate)
            : "memory");
    }
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row swizzle atoms of 1024 B (SBO), LBO unused.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);         // start address      bits [0,14)
    d |= (uint64_t)0 << 16;                          // leading byte offs  bits [16,30)  (one atom along K)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // layout type: SWIZZLE_128B
    return d;
}
template <int KIND>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
    // c_format [4,6): 1 = F32, 2 = S32 ; a/b_format [7,10)/[10,13): TF32 = 2, signed int8 = 1 ; a/b major bits 15/16 = 0 (K-major)
    // n_dim [17,23) = N >> 3 ; m_dim [24,29) = M >> 4
    return (KIND == KIND_TF32 ? (1u << 4) | (2u << 7) | (2u << 10) : (2u << 4) | (1u << 7) | (1u << 10)) |
           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ epilogues
struct EpiF32 {           // C = alpha*acc + beta*C   (FP32, column-major)
    float *C;
    int64_t ldc;
    float alpha, beta;
    using prev_t = float;
    __device__ __forceinline__ float load(int row, int col) const { return beta != 0.f ? C[row + (int64_t)col * ldc] : 0.f; }
    __device__ __forceinline__ void store(int row, int col, uint32_t raw, float prev) const {
        C[row + (int64_t)col * ldc] = alpha * __uint_as_float(raw) + beta * prev;
    }
};
// ------------------------------------------------------------------------------------------ the kernel
// Tile rasterisation: CTAs are dispatched in linear block order (x fastest).  With x = m-tile a wave of 148 CTAs touched every
// m-tile, i.e. ALL slices of A (134 MB at 4096^2) per wave, and ncu showed A re-read from DRAM once per wave (2.06 GB per launch for
// 0.27 GB of operands).  Remap so that consecutive (cluster) ids walk GROUP_M m-tiles x all n-tiles: a wave then works on
// ~GROUP_M m-tiles of A plus a few n-tiles of B, which stay in the 126 MB L2 for the waves that follow.
__device__ __forceinline__ void raster(int cid, int tiles_m, int tiles_n, int GROUP_M, int &m_tile, int &n_tile) {
    const int per_group = GROUP_M * tiles_n;
    const int group = cid / per_group, first_m = group * GROUP_M;
    const int gsz = min(GROUP_M, tiles_m - first_m);
    const int in_group = cid - group * per_group;
    m_tile = first_m + in_group % gsz;
    n_tile = in_group / gsz;
}

__device__ __forceinline__ void mbar_arrive1(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// The tensor core ACCUMULATES with truncation, not round-to-nearest: every tcgen05.mma step shrinks |accumulator| by up to an
// ulp, a BIAS that grows linearly with the number of steps chained into one accumulator (measured on B200, all-positive data:
// -2.7e-6 relative at K = 1024, -1.3e-5 at 4096, -5.4e-5 at 16384 - profiles/README.md; the FP32 contract is 1e-5).  So a chain
// is never longer than one K-CHUNK of KCHUNK_BLOCKS stages (512 elements x 3 segments = 192 steps): the two TMEM accumulators
// take alternate chunks, and EIGHT epilogue warps (two per TMEM lane quarter, 128 columns each) drain a finished chunk into
// running sums held in registers with round-to-nearest FP32 adds while the MMA warp is already filling the other accumulator;
// C is read and written once, at the end.  The error then no longer depends on K (~3e-6 worst case, measured).
constexpr int KCHUNK_BLOCKS = 16;
constexpr int THREADS_F32 = 64 + 8 * 32;     // TMA warp, MMA warp, 8 epilogue warps

// (10 warps are allocated as 12 - warp allocation granularity 4 - so the budget is 65536 / (12 * 32) = 170 registers per thread)
template <int KIND, typename Epi>
__global__ void __launch_bounds__(THREADS_F32, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                 Segs segs, Epi epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);      // SWIZZLE_128B needs 1024B alignment
    uint8_t *sA = smem;
    uint8_t *sB = smem + STAGES * A_STAGE;
    uint64_t *full = (uint64_t *)(smem + STAGES * (A_STAGE + B_STAGE));
    uint64_t *empty = full + STAGES;
    uint64_t *tmem_full = empty + STAGES;      // [2]
    uint64_t *tmem_empty = tmem_full + 2;      // [2]
    uint32_t *tmem_ptr = (uint32_t *)(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int m_tile, n_tile;
    raster((int)(blockIdx.x + gridDim.x * blockIdx.y), (int)gridDim.x, (int)gridDim.y, 8, m_tile, n_tile);
    const int m0 = m_tile * BM, n0 = n_tile * BN;
    constexpr int ELEMS = KIND == KIND_TF32 ? BKB / 4 : BKB;      // K elements per stage
    constexpr int KSTEP = KIND == KIND_TF32 ? 8 : 32;              // K elements per tcgen05.mma (32 bytes)
    const int kblocks = (K + ELEMS - 1) / ELEMS;
    const int chunks = (kblocks + KCHUNK_BLOCKS - 1) / KCHUNK_BLOCKS;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(2 * TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int c = 0; c < chunks; ++c) {
                const int kb0 = c * KCHUNK_BLOCKS, kb1 = min(kblocks, kb0 + KCHUNK_BLOCKS);
                for (int seg = 0; seg < segs.n; ++seg)
                    for (int kb = kb0; kb < kb1; ++kb, ++it) {
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], A_STAGE + B_STAGE);
                        tma_load_3d(sA + s * A_STAGE, &tmA, &full[s], kb * ELEMS, m0, segs.sa[seg]);
                        tma_load_3d(sB + s * B_STAGE, &tmB, &full[s], kb * ELEMS, n0, segs.sb[seg]);
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc<KIND>();
            int it = 0;
            for (int c = 0; c < chunks; ++c) {
                const int buf = c & 1;
                mbar_wait(&tmem_empty[buf], ((c >> 1) & 1) ^ 1);          // the epilogue drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
                const int kb0 = c * KCHUNK_BLOCKS, kb1 = min(kblocks, kb0 + KCHUNK_BLOCKS);
                const int iters = segs.n * (kb1 - kb0);
                for (int j = 0; j < iters; ++j, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint64_t ad = make_smem_desc(smem_u32(sA + s * A_STAGE));
                    const uint64_t bd = make_smem_desc(smem_u32(sB + s * B_STAGE));
#pragma unroll
                    for (int k = 0; k < ELEMS / KSTEP; ++k) {
                        // advance 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
                        tc_mma<KIND>(acc, ad + 2 * k, bd + 2 * k, idesc, (j > 0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(&empty[s]);            // frees the smem stage when these MMAs retire
                }
                tc_commit(&tmem_full[buf]);          // this chunk's accumulator is complete
            }
        }
    } else {
        // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32); warps 2-5 take columns [0,128), warps 6-9 [128,256)
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int row = m0 + q * 32 + lane;
        constexpr int HC = BN / 2;
        float run[HC];
#pragma unroll
        for (int j = 0; j < HC; ++j) run[j] = 0.f;
        for (int c = 0; c < chunks; ++c) {
            const int buf = c & 1;
            mbar_wait(&tmem_full[buf], (c >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < HC; c0 += 16) {
                uint32_t v[16];
                tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * HC + c0), v);
#pragma unroll
                for (int j = 0; j < 16; ++j) run[c0 + j] += __uint_as_float(v[j]);        // FP32 round-to-nearest
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive1(&tmem_empty[buf]);
        }
        if (row < M) {
#pragma unroll
            for (int c0 = 0; c0 < HC; c0 += 16) {
                // the loads of the old C are issued in batches before the stores (a load->store->load chain cost ~0.4 ms per GEMM)
                typename Epi::prev_t prev[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = n0 + half * HC + c0 + j;
                    if (col < N) prev[j] = epi.load(row, col);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = n0 + half * HC + c0 + j;
                    if (col < N) epi.store(row, col, __float_as_uint(run[c0 + j]), prev[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * TMEM_COLS));
    }
}

// ------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// [slices][rows][Kp] K-major operand, element size es; box = (BKB/es, box_rows, 1), 128B swizzle
static bool make_map(CUtensorMap *map, CUtensorMapDataType dt, int es, void *base, int64_t Kp, int64_t rows, int slices, int box_rows,
                     int box_k_bytes = BKB, int box_slices = 1, int64_t rows_stride = 0) {      // rows_stride: rows between slices (a window into a larger operand)
    if (rows_stride <= 0) rows_stride = rows;
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)slices};
    cuuint64_t strides[2] = {(cuuint64_t)Kp * es, (cuuint64_t)Kp * es * rows_stride};
    cuuint32_t box[3] = {(cuuint32_t)(box_k_bytes / es), (cuuint32_t)box_rows, (cuuint32_t)box_slices};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, dt, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              box_k_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (box_k_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// grow-only device workspace of the current context (GemmState, kernels.h)
static void *workspace(size_t bytes) {
    GemmState &G = gemm_state();
    if (bytes > G.ws_bytes) {
        if (G.ws) cudaFree(G.ws);
        G.ws = nullptr;
        G.ws_bytes = 0;
        G.alloc_gen++;
        if (cudaMalloc(&G.ws, bytes) != cudaSuccess) return nullptr;
        G.ws_bytes = bytes;
    }
    return G.ws;
}

// ------------------------------------------------------------------------------------------ 3xTF32 operand split
// dst[0] = hi = RN_tf32(x), dst[1] = lo = x - hi, both [rows][Kp] with k contiguous (zero padded to Kp).
// KMAJOR source: element (r,k) at src[k + r*ld]; otherwise at src[r + k*ld] (transposed through shared memory).
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
template <bool KMAJOR>
__global__ void __launch_bounds__(256) k_split_tf32(float *__restrict__ dst, const float *__restrict__ src, int64_t ld, int rows,
                                                    int K, int Kp) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    const int64_t plane = (int64_t)rows * Kp;
    if (KMAJOR) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            int r = r0 + ty + i, k = k0 + tx;
            if (r < rows && k < Kp) {
                float x = k < K ? src[k + (int64_t)r * ld] : 0.f;
                float hi = tf32_rn(x);
                dst[(int64_t)r * Kp + k] = hi;
                dst[plane + (int64_t)r * Kp + k] = x - hi;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            int k = k0 + ty + i, r = r0 + tx;
            tile[ty + i][tx] = (r < rows && k < K) ? src[r + (int64_t)k * ld] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            int r = r0 + ty + i, k = k0 + tx;
            if (r < rows && k < Kp) {
                float x = tile[tx][ty + i];
                float hi = tf32_rn(x);
                dst[(int64_t)r * Kp + k] = hi;
                dst[plane + (int64_t)r * Kp + k] = x - hi;
            }
        }
    }
}


// ==================================================================================================
// FP64 on the tcgen05 INT8 pipe: Ozaki-style error-free slicing
// ==================================================================================================
// op(A)(i,k) = 2^ea[i] * sum_p qa_p(i,k) 2^(-7(p+1)) + O(2^(ea[i]-7s-1)),  qa_p in [-64,64] (signed 7-bit slices), same for
// the columns of op(B).  Slice products are summed EXACTLY in int32 (K*(d+1)*2^12 < 2^31), grouped by d = p+q:
//     C(i,j) = 2^(ea[i]+eb[j]) * sum_d 2^(-7(d+2)) * D_d(i,j),   D_d = sum_{p+q=d} Qa_p * Qb_q^T   (int32)
// Groups are processed from the smallest magnitude (d = s-1) to the largest (d = 0); each group is one accumulation chain in
// one of two TMEM buffers, and the epilogue of group d (FP64 read-modify-write of the C tile, L2 resident) overlaps the MMAs of
// the next group.  Only pairs with p+q <= s-1 are formed: the dropped ones are below the slicing truncation.
constexpr int OZ_MAX_SLICES = 8;    // 7*8 = 56 fixed-point bits fit an int64 with headroom; 8 x 64 TMEM columns = 512

// ---- pass 1: per-row max |x| (bit pattern of a non-negative double orders like an unsigned integer), then the binary
// exponent e = ilogb(max) + 2 (so |x| 2^-e <= 0.5) and the scale 2^e
// The same pass also takes the row's 1-norm and its number of non-zero elements: the accuracy certificate (guard) bounds the
// error of element (i,j) through them instead of through K * max, which keeps it sharp for sparse operands (the one-hot seed
// matrices of jacobian!) and for rows dominated by a few large elements.  stats = [max bits | 1-norm | nnz], `rows` apart.
template <bool KMAJOR>
__global__ void __launch_bounds__(256) k_row_absmax(const double *__restrict__ src, int64_t ld, int rows, int K,
                                                    unsigned long long *__restrict__ mx_bits) {
    double *l1 = reinterpret_cast<double *>(mx_bits) + rows, *nz = l1 + rows;
    if (KMAJOR) {                               // one warp per row, k contiguous
        int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
        if (r >= rows) return;
        const double *p = src + (int64_t)r * ld;
        double mx = 0.0, s1 = 0.0, cnt = 0.0;
        // NaN/Inf anywhere in the row poisons the row: fmax() would silently drop a NaN, so non-finite elements count as +Inf
        int kdone = 0;
        if ((ld & 1) == 0 && ((uintptr_t)src & 15) == 0) {          // 16-byte loads, four in flight per lane
            const int Kv = K & ~1;
#pragma unroll 4
            for (int k = 2 * lane; k < Kv; k += 64) {
                const double2 t = *reinterpret_cast<const double2 *>(p + k);
                const double v0 = fabs(t.x), v1 = fabs(t.y);
                mx = fmax(mx, v0 <= 1.7976931348623157e308 ? v0 : __longlong_as_double(0x7FF0000000000000LL));
                mx = fmax(mx, v1 <= 1.7976931348623157e308 ? v1 : __longlong_as_double(0x7FF0000000000000LL));
                s1 += v0; s1 += v1; cnt += v0 != 0.0 ? 1.0 : 0.0; cnt += v1 != 0.0 ? 1.0 : 0.0;
            }
            kdone = Kv;
        }
        for (int k = kdone + lane; k < K; k += 32) {
            double v = fabs(p[k]);
            mx = fmax(mx, v <= 1.7976931348623157e308 ? v : __longlong_as_double(0x7FF0000000000000LL));
            s1 += v; cnt += v != 0.0 ? 1.0 : 0.0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            s1 += __shfl_xor_sync(0xffffffffu, s1, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (lane == 0) { mx_bits[r] = (unsigned long long)__double_as_longlong(mx); l1[r] = s1; nz[r] = cnt; }
    } else {                                    // threads along rows (coalesced), blockIdx.y splits K into chunks of 64
        int r = blockIdx.x * 256 + threadIdx.x;
        if (r >= rows) return;
        int k0 = blockIdx.y * 64, k1 = min(K, k0 + 64);
        double mx = 0.0, s1 = 0.0, cnt = 0.0;
#pragma unroll 8
        for (int k = k0; k < k1; ++k) {
            double v = fabs(src[r + (int64_t)k * ld]);
            mx = fmax(mx, v <= 1.7976931348623157e308 ? v : __longlong_as_double(0x7FF0000000000000LL));
            s1 += v; cnt += v != 0.0 ? 1.0 : 0.0;
        }
        atomicMax(mx_bits + r, (unsigned long long)__double_as_longlong(mx));
        if (cnt != 0.0) { atomicAdd(l1 + r, s1); atomicAdd(nz + r, cnt); }
    }
}
__global__ void __launch_bounds__(256) k_row_exponent(const unsigned long long *__restrict__ mx_bits, int rows, int bits, int *__restrict__ e_out,
                                                      double *__restrict__ scale_out) {
    int r = blockIdx.x * 256 + threadIdx.x;
    if (r == 0) reinterpret_cast<unsigned long long *>(scale_out)[3 * (int64_t)rows] = 0ull;     // plane word (planes_note), filled by the slice pass
    if (r >= rows) return;
    double mx = __longlong_as_double((long long)mx_bits[r]);
    const bool finite = isfinite(mx);
    // 7-bit digits: |x| 2^-e < 1/2, leading digit <= 64.  8-bit digits: one more bit of headroom (|x| 2^-e < 1/4) so that the
    // leading digit plus the carry from below (<= 65) stays inside int8
    int e = (mx > 0.0 && finite) ? ilogb(mx) + (bits == 8 ? 3 : 2) : 0;
    e_out[r] = e;
    // a row/column holding NaN or Inf gets a NaN scale: every output it touches becomes NaN, as DGEMM would propagate it
    // an all-zero row gets scale 0 (its digits are all zero anyway): it then adds nothing to the guard's error bound
    scale_out[r] = finite ? (mx > 0.0 ? scalbn(1.0, e) : 0.0) : __longlong_as_double(0x7FF8000000000000LL);
    // 1-norm in units of the scale (<= nnz / 4); 1 % covers the rounding of the running sums
    double *l1 = scale_out + rows;
    l1[r] = (finite && mx > 0.0) ? 1.01 * scalbn(l1[r], -e) : 0.0;
}

// ---- significant planes ("plane word").  Every slicing pass also ORs the fixed-point values X of the whole operand into one
// 64-bit word stored behind the row statistics (index 3 * rows of the scale array).  If the low 8 z bits of that word are zero,
// the z least significant digit planes of the operand are identically zero (a zero low byte of X is digit 0 with no carry), so
// every slice pair that touches them contributes exactly nothing: the GEMM kernels read the word and neither load nor multiply
// those planes - bit-identical results with fewer products.  Operands of few significant bits have such planes: constant fills
// (the adjoint of `sum`: one plane), one-hot / integer-valued seeds, Float32 data held in Float64 (24 bits: four planes).
__device__ __forceinline__ void planes_publish_warp(unsigned long long *word, unsigned long long orv) {
    unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)orv), hi = __reduce_or_sync(0xffffffffu, (unsigned)(orv >> 32));
    if ((threadIdx.x & 31) == 0) {
        const unsigned long long v = ((unsigned long long)hi << 32) | lo;
        if (v & ~*(volatile unsigned long long *)word) atomicOr(word, v);      // the plain read is only a filter (bits are only ever set)
    }
}
__device__ __forceinline__ void planes_publish_block(unsigned long long *word, unsigned long long orv, unsigned long long *sh) {
    unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)orv), hi = __reduce_or_sync(0xffffffffu, (unsigned)(orv >> 32));
    if (threadIdx.x == 0) *sh = 0ull;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && (lo | hi)) atomicOr(sh, ((unsigned long long)hi << 32) | lo);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long v = *sh;
        if (v & ~*(volatile unsigned long long *)word) atomicOr(word, v);
    }
}
// planes of an operand that can hold a non-zero digit, from its plane word (1 for an all-zero operand)
__device__ __forceinline__ int planes_of(const unsigned long long *word, int nslices, int bits) {
    if (!word) return nslices;
    const unsigned long long v = *word;
    if (!v) return 1;
    const int z = (__ffsll((long long)v) - 1) / bits;
    return z >= nslices ? 1 : nslices - z;
}

// 2^h as the product of two normal doubles (h = bits * S - e >= -971 because e <= 1027; h > 1000 only for rows of denormal
// magnitude): X = rint((v * m1) * m2) equals rint(scalbn(v, h)) bit for bit - both products are exact unless the result is
// below 2^-1022, where either way rounds to X = 0 - at two multiplications instead of a library call per element.
__device__ __forceinline__ void pow2_pair(int h, double &m1, double &m2) {
    const int h1 = h > 1000 ? 1000 : h;
    m1 = __hiloint2double((1023 + h1) << 20, 0);
    m2 = __hiloint2double((1023 + (h - h1)) << 20, 0);
}

// Balanced base-256 digits of four fixed-point values at once (bits == 8).  Adding 0x80 at every digit position below the
// leading one turns "d >= 128 ? d - 256, carry" into plain byte extraction: byte p of Y = X + 0x80..80 is digit_p + 128, i.e. the
// digit's int8 bit pattern is that byte with its top bit flipped; the leading digit is the (signed) rest.  Bit-identical to the
// digit loop.  Stores plane s (s = 0 leading) of the four consecutive k of one row as one packed 32-bit word.
// NS > 0: the number of slices is a compile-time constant (the default S = 7 is instantiated: byte selectors, plane offsets and the
// bias become immediates - the generic kernels spent most of their issue slots computing them); NS = 0: `nslices` at run time.
template <int NS = 0>
__device__ __forceinline__ void store_digits8(int8_t *__restrict__ dst_row_k, int64_t plane, int nslices_rt, const long long X[4]) {
    const int nslices = NS ? NS : nslices_rt;
    const unsigned long long bias = nslices > 1 ? (0x8080808080808080ULL >> (8 * (9 - nslices))) : 0ULL;
    unsigned lo[4], hi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned long long y = (unsigned long long)X[j] + bias;
        lo[j] = (unsigned)y; hi[j] = (unsigned)(y >> 32);
    }
#pragma unroll
    for (int p = 0; p < (NS ? NS : 8); ++p) {                    // byte position p <-> slice index nslices - 1 - p
        if (p >= nslices) break;
        const unsigned *w = p < 4 ? lo : hi;
        const unsigned q = (unsigned)(p & 3), sel = q | ((4u + q) << 4);
        const unsigned h01 = __byte_perm(w[0], w[1], sel), h23 = __byte_perm(w[2], w[3], sel);
        unsigned packed = __byte_perm(h01, h23, 0x5410);
        if (p != nslices - 1) packed ^= 0x80808080u;
        *reinterpret_cast<uint32_t *>(dst_row_k + (int64_t)(nslices - 1 - p) * plane) = packed;
    }
}

// ---- pass 2: slices.  dst is [s][rows][Kp] int8, k contiguous.  Block = 32 rows x 128 k; thread = 1 row x 4 consecutive k per
// iteration, so every slice store is one packed 32-bit word and a warp writes 128 contiguous bytes of one row.
template <bool KMAJOR, int NS8 = 0>      // NS8 > 0: bits == 8 and nslices == NS8 are compile-time facts
__global__ void __launch_bounds__(256, 4) k_slice_i8(int8_t *__restrict__ dst, const double *__restrict__ src, int64_t ld, int rows, int K,
                                                  int Kp, int nslices_rt, int bits_rt, const int *__restrict__ e_in, unsigned long long *plane_word) {
    const int nslices = NS8 ? NS8 : nslices_rt, bits = NS8 ? 8 : bits_rt;
    __shared__ unsigned long long sh_or;
    // [k % 4][k / 4][r]: a thread later reads 4 consecutive k of one row, i.e. one element of each k%4 plane at row k/4 = lane -
    // 33-double rows make both the write (lanes along r) and that read (lanes along k/4) bank-conflict free
    __shared__ double tile[KMAJOR ? 1 : 4][KMAJOR ? 1 : 32][KMAJOR ? 1 : 33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 128;
    if (!KMAJOR) {
        // tile[k][r] <- src[(r0+r) + (k0+k)*ld], lanes along r (coalesced)
        for (int k = w; k < 128; k += 8) {
            int r = r0 + lane, kk = k0 + k;
            tile[k & 3][k >> 2][lane] = (r < rows && kk < K) ? src[r + (int64_t)kk * ld] : 0.0;
        }
        __syncthreads();
    }
    const int64_t plane = (int64_t)rows * Kp;
    const int k = k0 + 4 * lane;
    // fixed point, b = bits per digit (8: the full int8 range; 7: the first-generation layout): X = rint(v * 2^(bS - e)),
    // |X| <= 2^(bS-1); balanced base-2^b digits q_j in [-2^(b-1), 2^(b-1) - 1] from the least significant end
    // (d = X & (2^b - 1); X >>= b; if d >= 2^(b-1): d -= 2^b, X += 1) so that v 2^-e = sum_j q_j 2^(-b(j+1)) + O(2^(-bS-1)).
    // All four rows of a thread are loaded before any digit work so four 32-byte loads are in flight per thread.
    long long X[4][4];
    bool ok[4];
    const int radix = 1 << bits, half = radix >> 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rl = w + 8 * i, r = r0 + rl;
        ok[i] = r < rows && k < Kp;
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        int e = 0;
        if (ok[i]) {
            e = e_in[r];
            if (KMAJOR) {
                const double *p = src + (int64_t)r * ld + k;
                if (k + 3 < K && ((((uintptr_t)p) & 15) == 0)) {
                    double2 lo = *reinterpret_cast<const double2 *>(p), hi = *reinterpret_cast<const double2 *>(p + 2);
                    v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = (k + j < K) ? p[j] : 0.0;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = tile[j][lane][rl];
            }
        }
        double m1, m2;
        pow2_pair(bits * nslices - e, m1, m2);
#pragma unroll
        for (int j = 0; j < 4; ++j) X[i][j] = isfinite(v[j]) ? __double2ll_rn((v[j] * m1) * m2) : 0LL;
    }
    {
        unsigned long long orv = 0ull;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) orv |= (unsigned long long)X[i][j];
        planes_publish_block(plane_word, orv, &sh_or);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!ok[i]) continue;
        const int r = r0 + w + 8 * i;
        if (bits == 8) { store_digits8<NS8>(dst + (int64_t)r * Kp + k, plane, nslices, X[i]); continue; }
        for (int sidx = nslices - 1; sidx >= 0; --sidx) {
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int d = (int)(X[i][j] & (long long)(radix - 1));
                X[i][j] >>= bits;
                if (sidx > 0 && d >= half) { d -= radix; X[i][j] += 1; }
                if (sidx == 0) d = (int)((X[i][j] << bits) | d);       // most significant digit keeps the sign: |d| <= 2^(b-2) + 1
                packed |= ((uint32_t)(uint8_t)(int8_t)d) << (8 * j);
            }
            *reinterpret_cast<uint32_t *>(dst + sidx * plane + (int64_t)r * Kp + k) = packed;
        }
    }
}

// k-contiguous source, ONE pass: a warp owns a row, brings it into shared memory with cp.async (16-byte pieces), takes the row
// maximum there, derives the exponent and cuts the digits from the same copy - the row is read from HBM once instead of twice
// (k_row_absmax + k_row_exponent + k_slice_i8<true>).  blockDim = 32 x rows per CTA; dynamic shared memory = rows x Kp doubles.
__global__ void __launch_bounds__(128) k_slice_fused_kmajor(int8_t *__restrict__ dst, const double *__restrict__ src, int64_t ld, int rows, int K,
                                                            int Kp, int nslices, int bits, double *__restrict__ scale_out) {
    // (the plane word at scale_out[3 * rows] is zeroed by the launcher: planes_publish_warp only ever sets bits)
    extern __shared__ __align__(16) double srows[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int r = blockIdx.x * wpc + w;
    if (r >= rows) return;
    double *row = srows + (size_t)w * Kp;
    const double *p = src + (int64_t)r * ld;
    for (int c = lane; c < K / 2; c += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(row + 2 * c)), "l"(p + 2 * c) : "memory");
    for (int k = K + lane; k < Kp; k += 32) row[k] = 0.0;                       // (K is even here) zero padding up to Kp
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    double mx = 0.0, s1 = 0.0, cnt = 0.0;
    for (int k = 2 * lane; k < K; k += 64) {
        const double2 v = *reinterpret_cast<const double2 *>(row + k);
        const double a0 = fabs(v.x), a1 = fabs(v.y);
        // NaN/Inf anywhere in the row poisons the row (fmax would drop a NaN): non-finite elements count as +Inf
        mx = fmax(mx, a0 <= 1.7976931348623157e308 ? a0 : __longlong_as_double(0x7FF0000000000000LL));
        mx = fmax(mx, a1 <= 1.7976931348623157e308 ? a1 : __longlong_as_double(0x7FF0000000000000LL));
        s1 += a0 + a1; cnt += (a0 != 0.0 ? 1.0 : 0.0) + (a1 != 0.0 ? 1.0 : 0.0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        s1 += __shfl_xor_sync(0xffffffffu, s1, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    const bool finite = isfinite(mx);
    const int e = (mx > 0.0 && finite) ? ilogb(mx) + (bits == 8 ? 3 : 2) : 0;
    if (lane == 0) {
        scale_out[r] = finite ? (mx > 0.0 ? scalbn(1.0, e) : 0.0) : __longlong_as_double(0x7FF8000000000000LL);
        scale_out[rows + r] = (finite && mx > 0.0) ? 1.01 * scalbn(s1, -e) : 0.0;       // guard statistics, see k_row_absmax
        scale_out[2 * rows + r] = cnt;
    }
    const int64_t plane = (int64_t)rows * Kp;
    const int radix = 1 << bits, half = radix >> 1;
    unsigned long long orv = 0ull;
    double m1, m2;
    pow2_pair(bits * nslices - e, m1, m2);
    for (int k = 4 * lane; k < Kp; k += 128) {
        const double2 lo = *reinterpret_cast<const double2 *>(row + k), hi = *reinterpret_cast<const double2 *>(row + k + 2);
        const double v[4] = {lo.x, lo.y, hi.x, hi.y};
        long long X[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { X[j] = isfinite(v[j]) ? __double2ll_rn((v[j] * m1) * m2) : 0LL; orv |= (unsigned long long)X[j]; }
        if (bits == 8) { store_digits8(dst + (int64_t)r * Kp + k, plane, nslices, X); continue; }
        for (int sidx = nslices - 1; sidx >= 0; --sidx) {
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int d = (int)(X[j] & (long long)(radix - 1));
                X[j] >>= bits;
                if (sidx > 0 && d >= half) { d -= radix; X[j] += 1; }
                if (sidx == 0) d = (int)((X[j] << bits) | d);
                packed |= ((uint32_t)(uint8_t)(int8_t)d) << (8 * j);
            }
            *reinterpret_cast<uint32_t *>(dst + sidx * plane + (int64_t)r * Kp + k) = packed;
        }
    }
    planes_publish_warp(reinterpret_cast<unsigned long long *>(scale_out) + 3 * (int64_t)rows, orv);
}

// A UNIFORM operand (every element equals src[0]; GemmState::uniform): all rows share one exponent and one digit string, so the
// planes are fills: 117 MB of stores for a 4096^2 operand instead of 2 x 134 MB read and 117 MB written.  Same e,
// X, digits, scale and plane word as k_row_absmax + k_row_exponent + k_slice_i8 on the materialised matrix; the 1-norm statistic
// is K |v| instead of a running sum (it only feeds the error bound, with 1 % of slack).  Grid-stride over rows x Kp / 16.
__global__ void __launch_bounds__(256) k_slice_uniform(int8_t *__restrict__ dst, const double *__restrict__ src, int rows, int K, int Kp,
                                                       int nslices, int bits, double *__restrict__ scale_out) {
    const double v = src[0], av = fabs(v);
    const bool finite = av <= 1.7976931348623157e308;
    const int e = (av > 0.0 && finite) ? ilogb(av) + (bits == 8 ? 3 : 2) : 0;
    double m1, m2;
    pow2_pair(bits * nslices - e, m1, m2);
    const long long X = finite ? __double2ll_rn((v * m1) * m2) : 0LL;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = tid; r < rows; r += nth) {
        scale_out[r] = finite ? (av > 0.0 ? scalbn(1.0, e) : 0.0) : __longlong_as_double(0x7FF8000000000000LL);
        scale_out[rows + r] = (finite && av > 0.0) ? 1.01 * scalbn((double)K * av, -e) : 0.0;
        scale_out[2 * (int64_t)rows + r] = av != 0.0 ? (double)K : 0.0;
    }
    if (tid == 0) reinterpret_cast<unsigned long long *>(scale_out)[3 * (int64_t)rows] = K > 0 ? (unsigned long long)X : 0ull;
    // digit of plane p (p = 0 leading), as store_digits8 / the digit loop cut it
    const int radix = 1 << bits, half = radix >> 1;
    long long Y = X;
    int digit[OZ_MAX_SLICES];
    for (int sidx = nslices - 1; sidx >= 0; --sidx) {
        int d = (int)(Y & (long long)(radix - 1));
        Y >>= bits;
        if (sidx > 0 && d >= half) { d -= radix; Y += 1; }
        if (sidx == 0) d = (int)((Y << bits) | d);
        digit[sidx] = d;
    }
    const int64_t plane = (int64_t)rows * Kp, words_per_row = Kp / 16, nwords = (int64_t)rows * words_per_row;
    for (int p = 0; p < nslices; ++p) {                                  // every plane: not every GEMM kernel reads the plane word
        const uint32_t b = (uint32_t)(uint8_t)(int8_t)digit[p], w = b * 0x01010101u;
        uint4 *out = reinterpret_cast<uint4 *>(dst + p * plane);
        for (int64_t i = tid; i < nwords; i += nth) {
            const int k0 = (int)(i % words_per_row) * 16;
            uint4 o = make_uint4(w, w, w, w);
            if (k0 + 16 > K) {                                          // zero padding behind K
                uint32_t t[4] = {w, w, w, w};
                for (int j = 0; j < 16; ++j) if (k0 + j >= K) t[j >> 2] &= ~(0xFFu << (8 * (j & 3)));
                o = make_uint4(t[0], t[1], t[2], t[3]);
            }
            out[i] = o;
        }
    }
}

// Column-major source (rows strided by ld... i.e. consecutive ROWS are contiguous, k strided): tile of 128 rows x 32 k, so every
// column read is 1 KB contiguous (the 32-row x 128-k tile of the generic kernel reads 256-byte pieces 32 KB apart: every piece a
// different DRAM page).  Transposed through shared memory ([k % 4][k / 4][row]); a thread then owns 4 rows x 4 consecutive k and
// stores one packed 32-bit word per slice and row - a warp writes full 32-byte sectors of 4 rows.
template <int NS8 = 0>
__global__ void __launch_bounds__(256, 4) k_slice_i8_cm(int8_t *__restrict__ dst, const double *__restrict__ src, int64_t ld, int rows, int K,
                                                     int Kp, int nslices_rt, int bits_rt, const int *__restrict__ e_in, unsigned long long *plane_word) {
    const int nslices = NS8 ? NS8 : nslices_rt, bits = NS8 ? 8 : bits_rt;
    __shared__ double tile[4][8][132];
    __shared__ unsigned long long sh_or;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int r0 = blockIdx.y * 128, k0 = blockIdx.x * 32;
    {
        const int rr = tid & 127, kb = tid >> 7;                 // 128 threads along the rows, two columns per pass
        const int r = r0 + rr;
        double v[16];                                                // all sixteen column reads in flight before the first store
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int kk = k0 + kb + 2 * i;
            v[i] = (r < rows && kk < K) ? src[r + (int64_t)kk * ld] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) { const int k = kb + 2 * i; tile[k & 3][k >> 2][rr] = v[i]; }
    }
    __syncthreads();
    const int64_t plane = (int64_t)rows * Kp;
    const int kg = lane & 7, k = k0 + 4 * kg;
    const int radix = 1 << bits, half = radix >> 1;
    long long X[4][4];
    bool ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rl = w * 4 + (lane >> 3) + 32 * i, r = r0 + rl;
        ok[i] = r < rows && k < Kp;
        const int e = ok[i] ? e_in[r] : 0;
        double m1, m2;
        pow2_pair(bits * nslices - e, m1, m2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double v = tile[j][kg][rl];
            X[i][j] = (ok[i] && isfinite(v)) ? __double2ll_rn((v * m1) * m2) : 0LL;
        }
    }
    {
        unsigned long long orv = 0ull;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) orv |= (unsigned long long)X[i][j];
        planes_publish_block(plane_word, orv, &sh_or);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!ok[i]) continue;
        const int r = r0 + w * 4 + (lane >> 3) + 32 * i;
        if (bits == 8) { store_digits8<NS8>(dst + (int64_t)r * Kp + k, plane, nslices, X[i]); continue; }
        for (int sidx = nslices - 1; sidx >= 0; --sidx) {
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int d = (int)(X[i][j] & (long long)(radix - 1));
                X[i][j] >>= bits;
                if (sidx > 0 && d >= half) { d -= radix; X[i][j] += 1; }
                if (sidx == 0) d = (int)((X[i][j] << bits) | d);
                packed |= ((uint32_t)(uint8_t)(int8_t)d) << (8 * j);
            }
            *reinterpret_cast<uint32_t *>(dst + sidx * plane + (int64_t)r * Kp + k) = packed;
        }
    }
}

struct OzParams {
    double *C;
    int64_t ldc;
    double alpha, beta;
    const double *sa, *sb;      // 2^ea[row], 2^eb[col]
    int nslices;
    int group_m;                // tile rasterisation group (see raster())
    int bits;                   // bits per digit: slice pair (p, q) weighs 2^(-bits (p + q + 2))
    // accuracy certificate ("guard"): slot[0] <- max |alpha * (op(A) op(B))(i,j)| as computed, slot[1] <- max over (i,j) of the
    // rigorous error bound gscale * |alpha| * 2^ea[i] * 2^eb[j], gscale = K * c(bits, S) (guard_coeff); NULL = not recorded
    unsigned long long *guard;
    const double *la, *na, *lb, *nb;   // per row of op(A) / column of op(B): 1-norm in units of the scale, number of non-zeros
    double g_rho, g_h, g_drop;         // guard_coeffs()
    int dbg;                           // timing experiments only (RDB_OZAKI_DBG): 1 = no TMA loads, 2 = no epilogue; results are garbage
    // plane words of op(A) / op(B) (planes_publish_*; NULL = every plane is live) and the running count of slice pairs issued
    const unsigned long long *pa, *pb;
    unsigned long long *pairs;
    int allow_swap;                    // the trimmed path may swap the roles (both tilings have the same number of 2-CTA clusters: column tiles are padded to pairs)
    int pair_first;                    // umma_ozaki2tp_kernel was launched in front: one-plane products are its work, not the cluster kernel's
    int wide;                          // trimmed cluster path: 128-byte K-blocks when the A role has one live plane
};

// |x - x~| <= 2^(e - bS - 1) per element (rint), |x| <= 2^e / 4 (b = 8) or 2^e / 2 (b = 7), and the slice pairs with p + q >= S
// that the kernels drop weigh at most (S-1)/4 * 2^(-bS) per k (digits <= 2^(b-1)); 1 % covers the second-order terms:
// |(A B)(i,j) - computed| <= K * c * 2^ea[i] * 2^eb[j] + (a few ulps of the computed value, added in guard_fold).
// Per element, with s_a = 2^ea[i], s_b = 2^eb[j], abar = |a_i|_1 / s_a, bbar = |b_j|_1 / s_b, na / nb = non-zeros of the row / column
// (zeros are represented exactly and meet no dropped pair), h = 1/4 (b = 8) or 1/2 (b = 7) = max |x| / scale:
//   |(A B)(i,j) - computed| <= s_a s_b [ rho (min(abar, h nb) + min(bbar, h na)) + drop min(na, nb) ] + a few ulps of the value
// rho = 1.01 * 2^(-bS-1), drop = 1.01 (S-1)/4 2^(-bS).  Dense operands give the familiar K c s_a s_b.
static void guard_coeffs(int bits, int S, double &rho, double &h, double &drop) {
    rho = 1.01 * ldexp(1.0, -bits * S - 1);
    h = bits == 8 ? 0.25 : 0.5;
    drop = 1.01 * 0.25 * (S - 1) * ldexp(1.0, -bits * S);
}
__device__ __forceinline__ void guard_record(unsigned long long *slot, double gmax, double bmax, int lane) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
    }
    if (lane == 0) {
        // bit patterns of non-negative doubles order like unsigned integers; the plain read is only a filter (monotone slot)
        const unsigned long long g = (unsigned long long)__double_as_longlong(gmax), b = (unsigned long long)__double_as_longlong(bmax);
        if (g > *(volatile unsigned long long *)&slot[0]) atomicMax(&slot[0], g);
        if (b > *(volatile unsigned long long *)&slot[1]) atomicMax(&slot[1], b);
    }
}
// one thread per used slot: flagged when bound + 16 ulp of the value exceeds tau * value, tau = 2^-40 (9.1e-13 < the 1e-12 contract)
__global__ void k_guard_fold(unsigned long long *slots, int n, GuardSummary *sum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double prod = __longlong_as_double((long long)slots[2 * i]), bnd0 = __longlong_as_double((long long)slots[2 * i + 1]);
    slots[2 * i] = 0ull; slots[2 * i + 1] = 0ull;
    atomicAdd(&sum->checked, 1ull);
    if (!(bnd0 > 0.0)) return;                                  // all-zero operand (scale 0): the product is exactly zero
    const double bnd = bnd0 + 16.0 * 1.1102230246251565e-16 * prod;
    const double ratio = prod > 0.0 ? bnd / prod : __longlong_as_double(0x7FF0000000000000LL);
    if (ratio > 9.094947017729282e-13) atomicAdd(&sum->flagged, 1ull);
    atomicMax((unsigned long long *)&sum->worst_ratio, (unsigned long long)__double_as_longlong(ratio));
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
umma_ozaki_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K, OzParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = smem + STAGES * A_STAGE;
    uint64_t *full = (uint64_t *)(smem + STAGES * (A_STAGE + B_STAGE));
    uint64_t *empty = full + STAGES;
    uint64_t *tmem_full = empty + STAGES;      // [2]
    uint64_t *tmem_empty = tmem_full + 2;      // [2]
    uint32_t *tmem_ptr = (uint32_t *)(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kblocks = (K + BKB - 1) / BKB;
    const int S = P.nslices;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int d = S - 1; d >= 0; --d)
                for (int p = 0; p <= d; ++p)
                    for (int kb = 0; kb < kblocks; ++kb, ++it) {
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], A_STAGE + B_STAGE);
                        tma_load_3d(sA + s * A_STAGE, &tmA, &full[s], kb * BKB, m0, p);
                        tma_load_3d(sB + s * B_STAGE, &tmB, &full[s], kb * BKB, n0, d - p);
                    }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc<KIND_I8>();
            int it = 0, gi = 0;
            for (int d = S - 1; d >= 0; --d, ++gi) {
                const int buf = gi & 1;
                mbar_wait(&tmem_empty[buf], ((gi >> 1) & 1) ^ 1);          // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
                const int iters = (d + 1) * kblocks;
                for (int j = 0; j < iters; ++j, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint64_t ad = make_smem_desc(smem_u32(sA + s * A_STAGE));
                    const uint64_t bd = make_smem_desc(smem_u32(sB + s * B_STAGE));
#pragma unroll
                    for (int k = 0; k < BKB / 32; ++k) tc_mma<KIND_I8>(acc, ad + 2 * k, bd + 2 * k, idesc, (j > 0 || k > 0) ? 1u : 0u);
                    tc_commit(&empty[s]);
                }
                tc_commit(&tmem_full[buf]);
            }
        }
    } else {
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        const double sa = row < M ? P.sa[row] : 0.0;
        int gi = 0;
        for (int d = S - 1; d >= 0; --d, ++gi) {
            const int buf = gi & 1;
            mbar_wait(&tmem_full[buf], (gi >> 1) & 1);
            tc_fence_after();
            const double rs = P.alpha * sa * scalbn(1.0, -P.bits * (d + 2));     // alpha * 2^(ea[row] - bits(d+2))
            const bool first = gi == 0;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0), v);
                if (row < M) {
                    double prev[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = n0 + c0 + j;
                        prev[j] = 0.0;
                        if (col < N && !(first && P.beta == 0.0)) prev[j] = P.C[row + (int64_t)col * P.ldc];
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = n0 + c0 + j;
                        if (col < N) {
                            double t = (double)(int)v[j] * rs * P.sb[col];
                            P.C[row + (int64_t)col * P.ldc] = (first ? P.beta * prev[j] : prev[j]) + t;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}


// --------------------------------------------------------------------------------------------------
// Second generation: every loaded slice is reused for ALL of its pairs.  Tile 128 x 64, one TMEM accumulator per slice-sum d
// (S x 64 columns <= 512), K-blocks of 64 bytes (64B swizzle).  One stage = all S slices of the A tile and of the B tile for one
// K-block (S*(8 KB + 4 KB)); the MMA thread then issues the S(S+1)/2 pair products on it.  L2->SM traffic per MAC drops 2.25x
// against the pair-at-a-time kernel above (which ncu/timing showed L2-bound), and the FP64 result is assembled in registers and
// written ONCE instead of read-modify-written per group.
namespace oz2 {
constexpr int BM2 = 128, BN2 = 64;
// KB = K bytes per stage: 64 (64B swizzle, 2 stages at S = 8) or 32 (32B swizzle, 4 stages: a deeper ring hides the TMA latency
// that a 2-deep ring of 96 KB stages exposed)
__host__ __device__ constexpr int stages_for(int KB) { return KB == 64 ? 2 : 4; }
__host__ __device__ constexpr int stage_bytes(int S, int KB) { return S * (BM2 + BN2) * KB; }
__host__ __device__ constexpr int smem_bytes(int S, int KB) { return stages_for(KB) * stage_bytes(S, KB) + 1024 + 256; }
__host__ __device__ constexpr int smem_bytes_cluster(int S, int CL) { return 2 * (((S + CL - 1) / CL) * CL * BM2 * 64 + S * BN2 * 64) + 1024 + 256; }

// K-major swizzled tile with rows of KB bytes: 8-row atoms of 8*KB bytes (SBO); layout type 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
template <int KB>
__device__ __forceinline__ uint64_t make_desc_kb(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((8 * KB) >> 4) << 32;            // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                          // descriptor version
    d |= (uint64_t)(KB == 64 ? 4 : 6) << 61;
    return d;
}
__device__ __forceinline__ constexpr uint32_t make_idesc_i8(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM2 >> 4) << 24);
}

__device__ __forceinline__ void tc_ld8_issue(uint32_t taddr, uint32_t (&v)[8]) {        // no wait: batch several, then tc_wait_ld()
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Epilogue warps of the gen-2 kernels: EPI_CB column blocks x 4 TMEM lane quarters.  With 4 warps (one per scheduler) the epilogue
// was a single dependent instruction stream per scheduler and cost 0.15 ms of a 1.30 ms 4096^3 launch (timing experiment with the
// epilogue switched off, profiles/README.md); 16 warps hide that latency.
constexpr int EPI_CB = 4;
constexpr int OZ_THREADS = 64 + 128 * EPI_CB;

// Epilogue shared by the gen-2 kernels: the S int32 accumulators of a 128 x 64 tile (TMEM columns d*64 .. d*64+63) are assembled
// into FP64, scaled by alpha * 2^(ea[row] + eb[col]) and written once.  Warp (q4, cb) owns TMEM lanes [32 q4, 32 q4 + 32) = tile rows
// and the columns [cb CW, cb CW + CW), CW = 64 / EPI_CB.
//  * the S tcgen05.ld of an 8-column block are issued back to back and waited for once;
//  * accumulators are first combined EXACTLY three at a time in 64-bit integers (D_d carries weight 2^(-b(d+2)), so
//    D_{3g} 2^(2b) + D_{3g+1} 2^b + D_{3g+2} < 2^(31+2b) <= 2^47 shares the weight 2^(-b(3g+4))): ceil(S/3) int->double
//    conversions and FP64 adds per element instead of S of each, and fewer roundings (smallest group first);
//  * the guard's two maxima ride along: max |term| through the high word of the double (integer max; an underestimate by at most
//    2^-20, i.e. on the safe side), the error bound once per thread from the column maxima of the warp's block (G is monotone in them).
// nd = number of accumulators the main loop wrote (S unless planes were trimmed: na + nb - 1); the others count as zero
// EpiRole: which operand supplies the tile's rows.  The trimmed cluster path may compute the tile of C^T = op(B)^T op(A)^T instead
// (rows = columns of C, the operand with fewer live planes in the A role, see oz2_trimmed_cluster): CT = true addresses C transposed.
struct EpiRole { const double *srow, *scol, *lrow, *nrow, *lcol, *ncol; };
__device__ __forceinline__ EpiRole epi_role(const OzParams &P, bool swapped = false) {
    return swapped ? EpiRole{P.sb, P.sa, P.lb, P.nb, P.la, P.na} : EpiRole{P.sa, P.sb, P.la, P.na, P.lb, P.nb};
}
template <int S, bool CT = false>
__device__ __forceinline__ void oz2_epilogue(uint32_t tmem_base, int q4, int cb, int lane, int m0, int n0, int M, int N, const OzParams &P, int nd,
                                             const EpiRole R) {
    constexpr int CW = BN2 / EPI_CB;
    const int row = m0 + q4 * 32 + lane;
    const bool live = row < M;
    const double rs = live ? P.alpha * R.srow[row] : 0.0;
    const int64_t rstride = CT ? P.ldc : 1, cstride = CT ? 1 : P.ldc;
    constexpr int NG = (S + 2) / 3;
    double wg[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const int last = (3 * g + 2 < S) ? 3 * g + 2 : S - 1;                       // weight of the group's LAST member
        wg[g] = scalbn(1.0, -P.bits * (last + 2));
    }
    int gmax_hi = 0;
#pragma unroll 1
    for (int c0 = cb * CW; c0 < (cb + 1) * CW; c0 += 8) {
        uint32_t v[S][8];
#pragma unroll
        for (int d = 0; d < S; ++d) tc_ld8_issue(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(d * BN2 + c0), v[d]);
        double prev[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {                                                  // the old C (beta != 0) travels under the TMEM loads
            const int col = n0 + c0 + j;
            prev[j] = (live && col < N && P.beta != 0.0) ? P.C[row * rstride + col * cstride] : 0.0;
        }
        tc_wait_ld();
        if (!live) continue;
        // transposed addressing: a thread's eight columns are eight CONSECUTIVE doubles of C - 16-byte stores when aligned
        const bool vec = CT && n0 + c0 + 8 <= N && (P.ldc & 1) == 0 && (((uintptr_t)P.C) & 15) == 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + c0 + j;
            if (col >= N) continue;
            double sum = 0.0;
#pragma unroll
            for (int g = NG - 1; g >= 0; --g) {                                        // smallest weight first
                const int last = (3 * g + 2 < S) ? 3 * g + 2 : S - 1;
                long long acc = 0;
#pragma unroll
                for (int d = 3 * g; d <= last; ++d) acc += (long long)(d < nd ? (int)v[d][j] : 0) << (P.bits * (last - d));   // (unwritten accumulators hold stale TMEM)
                sum += (double)acc * wg[g];
            }
            const double term = sum * (rs * R.scol[col]);
            gmax_hi = max(gmax_hi, __double2hiint(term) & 0x7fffffff);
            if (CT) prev[j] = P.beta * prev[j] + term;
            else P.C[row * rstride + col * cstride] = P.beta * prev[j] + term;
        }
        if (CT) {
            double *dst = P.C + row * rstride + (n0 + c0);
            if (vec) {
#pragma unroll
                for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2 *>(dst + j) = make_double2(prev[j], prev[j + 1]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (n0 + c0 + j < N) dst[j] = prev[j];
            }
        }
    }
    if (P.guard) {
        // column maxima of (scale, 1-norm, nnz) over the warp's CW columns, then ONE bound per thread (row)
        double sbm = 0.0, lbm = 0.0, nbm = 0.0;
        {
            const int col = n0 + cb * CW + lane;
            if (lane < CW && col < N) { sbm = fabs(R.scol[col]); lbm = R.lcol[col]; nbm = R.ncol[col]; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sbm = fmax(sbm, __shfl_xor_sync(0xffffffffu, sbm, o));
            lbm = fmax(lbm, __shfl_xor_sync(0xffffffffu, lbm, o));
            nbm = fmax(nbm, __shfl_xor_sync(0xffffffffu, nbm, o));
        }
        double bnd = 0.0;
        if (live) {
            const double la = R.lrow[row], na = R.nrow[row];
            const double G = P.g_rho * (fmin(la, P.g_h * nbm) + fmin(lbm, P.g_h * na)) + P.g_drop * fmin(na, nbm);
            bnd = fabs(rs) * sbm * G;
        }
        gmax_hi = __reduce_max_sync(0xffffffffu, gmax_hi);
        guard_record(P.guard, __hiloint2double(gmax_hi, 0), bnd, lane);
    }
}

// ---- trimmed main loop: only na planes of op(A) and nb planes of op(B) can hold non-zero digits (plane words, see
// planes_publish_*).  Pairs (p, q) with p < na, q < nb, p + q <= S-1 are issued, merged over q as in the dense loop; accumulator
// d = p + q is first written (accumulate = 0) by the smallest p that reaches it.  Dynamic loops: this path trades the unrolled
// issue sequence for up to 28x fewer products.
template <int S, int KB>
__device__ __forceinline__ void oz2_issue_trimmed(uint32_t a0, uint32_t b0, uint32_t tmem_base, int na, int nb, bool first_kblock) {
    constexpr int A_SLICE = BM2 * KB, B_SLICE = BN2 * KB;
    int init = 0;                                                  // accumulators written so far in this tile (first K-block only)
    for (int p = 0; p < na; ++p) {
        const int nqt = nb < S - p ? nb : S - p;
        const uint64_t ad = make_desc_kb<KB>(a0 + p * A_SLICE);
        for (int q = 0; q < nqt;) {
            int nq = nqt - q < 4 ? nqt - q : 4;
            uint32_t acc0 = 1u;
            if (first_kblock) {
                if (p + q >= init) acc0 = 0u;                      // d = p + q and everything above it in this chunk is new
                else if (init - (p + q) < nq) nq = init - (p + q); // the already written accumulators first, the new one next round
            }
            const uint64_t bd = make_desc_kb<KB>(b0 + q * B_SLICE);
            const uint32_t acc = tmem_base + (uint32_t)((p + q) * BN2);
            const uint32_t idesc = make_idesc_i8(nq * BN2);
#pragma unroll
            for (int k = 0; k < KB / 32; ++k) tc_mma<KIND_I8>(acc, ad + 2 * k, bd + 2 * k, idesc, k > 0 ? 1u : acc0);
            q += nq;
        }
        if (first_kblock && p + nqt > init) init = p + nqt;
    }
}
__device__ __forceinline__ int oz2_pairs(int S, int na, int nb) {
    int n = 0;
    for (int p = 0; p < na; ++p) n += nb < S - p ? nb : S - p;
    return n;
}

template <int S, int KB>
__global__ void __launch_bounds__(OZ_THREADS, 1)
umma_ozaki2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmA1,
                   const __grid_constant__ CUtensorMap tmB1, int M, int N, int K, OzParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int STAGE = stage_bytes(S, KB), STAGES2 = stages_for(KB), BKB2 = KB, A_SLICE = BM2 * KB, B_SLICE = BN2 * KB;
    uint64_t *full = (uint64_t *)(smem + STAGES2 * STAGE);
    uint64_t *empty = full + STAGES2;
    uint64_t *tmem_full = empty + STAGES2;
    uint32_t *tmem_ptr = (uint32_t *)(tmem_full + 1);
    constexpr int TCOLS = S * BN2 <= 32 ? 32 : S * BN2 <= 64 ? 64 : S * BN2 <= 128 ? 128 : S * BN2 <= 256 ? 256 : 512;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int m_tile, n_tile;
    raster((int)(blockIdx.x + gridDim.x * blockIdx.y), (int)gridDim.x, (int)gridDim.y, P.group_m, m_tile, n_tile);
    const int m0 = m_tile * BM2, n0 = n_tile * BN2;
    const int kblocks = (K + BKB2 - 1) / BKB2;
    const int na = planes_of(P.pa, S, P.bits), nb = planes_of(P.pb, S, P.bits);     // live planes: tmA1 / tmB1 load them one by one
    const bool dense = na == S && nb == S;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES2;
                const uint32_t ph = (kb / STAGES2) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                if (!dense) {
                    mbar_expect_tx(&full[s], na * A_SLICE + nb * B_SLICE);
                    for (int p = 0; p < na; ++p) tma_load_3d(smem + s * STAGE + p * A_SLICE, &tmA1, &full[s], kb * BKB2, m0, p);
                    for (int q = 0; q < nb; ++q) tma_load_3d(smem + s * STAGE + S * A_SLICE + q * B_SLICE, &tmB1, &full[s], kb * BKB2, n0, q);
                    continue;
                }
                mbar_expect_tx(&full[s], STAGE);
                // one 3-D box per operand: (64 bytes of k) x rows x all S slices
                tma_load_3d(smem + s * STAGE, &tmA, &full[s], kb * BKB2, m0, 0);
                tma_load_3d(smem + s * STAGE + S * A_SLICE, &tmB, &full[s], kb * BKB2, n0, 0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES2;
                const uint32_t ph = (kb / STAGES2) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + s * STAGE), b0 = a0 + S * A_SLICE;
                if (!dense) { oz2_issue_trimmed<S, KB>(a0, b0, tmem_base, na, nb, kb == 0); tc_commit(&empty[s]); continue; }
                // For a fixed A slice p the partner slices q = 0..S-1-p are CONTIGUOUS rows of the B stage ([slice][64 rows]) and
                // their accumulators d = p+q are CONTIGUOUS TMEM column blocks, so up to four pairs go out as ONE MMA with
                // N = 64*(#q) <= 256: A is read from shared memory once per four pairs (N=64 MMAs were smem-read bound).
#pragma unroll
                for (int p = 0; p < S; ++p) {
                    const uint64_t ad = make_desc_kb<KB>(a0 + p * A_SLICE);
#pragma unroll
                    for (int q0 = 0; q0 < S - p; q0 += 4) {
                        const int nq = (S - p - q0) < 4 ? (S - p - q0) : 4;
                        const uint64_t bd = make_desc_kb<KB>(b0 + q0 * B_SLICE);
                        const uint32_t acc = tmem_base + (uint32_t)((p + q0) * BN2);
                        const uint32_t idesc = make_idesc_i8(nq * BN2);
#pragma unroll
                        for (int k = 0; k < BKB2 / 32; ++k)
                            tc_mma<KIND_I8>(acc, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || p > 0 || k > 0) ? 1u : 0u);
                    }
                }
                tc_commit(&empty[s]);
            }
            tc_commit(tmem_full);
            if (P.pairs && blockIdx.x == 0 && blockIdx.y == 0) atomicAdd(P.pairs, (unsigned long long)oz2_pairs(S, na, nb));
        }
    } else {
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        oz2_epilogue<S>(tmem_base, warp & 3, (warp - 2) >> 2, lane, m0, n0, M, N, P, dense ? S : (na + nb - 1 < S ? na + nb - 1 : S), epi_role(P));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS));
    }
}

// --------------------------------------------------------------------------------------------------
// Cluster variant: CL CTAs with the same m-block and adjacent n-blocks form a thread-block cluster and SHARE the A stage.
// Each CTA fetches S/CL of the A slices and TMA-multicasts them into every CTA of the cluster; B stays private.  The single-CTA
// kernel moves 96 KB per stage from L2 and ncu/timing showed it bound there (~8 TB/s aggregate L2->SM); with CL = 2 the per-CTA
// L2 traffic drops to 64 KB, with CL = 4 to 48 KB.  A stage may be refilled only when ALL CTAs of the cluster have consumed it,
// so tcgen05.commit arrives (multicast) on the empty barrier of every CTA and those barriers count CL arrivals.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

template <int S, int CL>
__global__ void __launch_bounds__(OZ_THREADS, 1)
umma_ozaki2c_kernel(const __grid_constant__ CUtensorMap tmAmc, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmA1n, const __grid_constant__ CUtensorMap tmB1w,
                    const __grid_constant__ CUtensorMap tmA1x, const __grid_constant__ CUtensorMap tmB1x, const __grid_constant__ CUtensorMap tmA1nx,
                    const __grid_constant__ CUtensorMap tmB1wx, int M, int N, int K, OzParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int KB = 64, STAGES2 = stages_for(KB), A_SLICE = BM2 * KB, B_SLICE = BN2 * KB;
    constexpr int SL = (S + CL - 1) / CL;                         // A slices fetched by each CTA; the stage holds SL*CL slots: the
    constexpr int SA = SL * CL;                                   // last box may reach past slice S-1 (TMA zero-fills, counts bytes)
    constexpr int STAGE = SA * A_SLICE + S * B_SLICE;
    constexpr uint16_t MASK = (uint16_t)((1u << CL) - 1u);
    constexpr int MAXST = 8;                                      // barrier pairs: the trimmed path runs a deeper ring of smaller stages
    uint64_t *full = (uint64_t *)(smem + STAGES2 * STAGE);
    uint64_t *empty = full + MAXST;
    uint64_t *tmem_full = empty + MAXST;
    uint32_t *tmem_ptr = (uint32_t *)(tmem_full + 1);
    constexpr int TCOLS = S * BN2 <= 32 ? 32 : S * BN2 <= 64 ? 64 : S * BN2 <= 128 ? 128 : S * BN2 <= 256 ? 256 : 512;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    int m_tile, n_group;
    raster((int)(blockIdx.x + gridDim.x * (blockIdx.y / CL)), (int)gridDim.x, (int)(gridDim.y / CL), P.group_m, m_tile, n_group);
    const int m0 = m_tile * BM2, n0 = (n_group * CL + (int)crank) * BN2;
    const int kblocks = (K + KB - 1) / KB;
    const int na = planes_of(P.pa, S, P.bits), nb = planes_of(P.pb, S, P.bits);     // live planes: tmA1 / tmB1 load them one by one
    const bool dense = na == S && nb == S;
    if (P.pair_first && (na < nb ? na : nb) == 1) return;         // umma_ozaki2tp_kernel did this product (uniform over the grid)

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < MAXST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                           // every CTA's barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (CL == 2 && !dense) {
        // ---- trimmed operands (plane words): only `na` planes of op(A) and `nb` of op(B) are live.
        //  * ROLES: the operand with fewer live planes takes the A role (128-row tiles): one MMA then covers up to four planes of
        //    the other operand (merged N), where the opposite assignment would issue N = 64 MMAs at 2/3 of the pipe rate.  If that
        //    is op(B) the CTA computes a tile of C^T = op(B)^T op(A)^T and its epilogue addresses C transposed (both tilings have
        //    the same number of clusters because column tiles are padded to pairs; P.allow_swap = 0 switches the swap off).
        //  * SHARING: the two CTAs of the cluster multicast whichever role weighs more per K-block: the A role (same rows, adjacent
        //    column tiles - the dense geometry) or the B role (adjacent row tiles, same columns).
        //  * RING: a stage holds just the live planes, so the 184 KB of the two dense stages become up to 8 stages: the loads of a
        //    short K-block (a handful of MMAs) need that depth to cover the TMA round trip.
        //  * WIDE K-BLOCKS: a one-plane A role moves 128 bytes of K per stage (128-byte swizzle, the tm*x maps) - half as many
        //    boxes, barrier rounds and commits per byte (P.wide: 0 switches it off, timing experiments).
        const bool swap = P.allow_swap && nb < na;
        const int nar = swap ? nb : na, nbr = swap ? na : nb;
        const int Rr = swap ? N : M, Cr = swap ? M : N;
        const bool wide = nar == 1 && P.wide;
        const int kbw = wide ? 128 : KB, a_slice = BM2 * kbw, b_slice = BN2 * kbw, kblocks_t = (K + kbw - 1) / kbw;
        const CUtensorMap *mapA = wide ? (swap ? &tmB1wx : &tmA1x) : (swap ? &tmB1w : &tmA1);
        const CUtensorMap *mapB = wide ? (swap ? &tmA1nx : &tmB1x) : (swap ? &tmA1n : &tmB1);
        const int RT = (Rr + BM2 - 1) / BM2, CT = (Cr + BN2 - 1) / BN2;
        const bool share_b = nbr * B_SLICE > nar * A_SLICE && RT % 2 == 0 && CT % 2 == 0;
        const int cid = (int)(blockIdx.x + gridDim.x * (blockIdx.y / 2));
        int r_tile, c_tile;
        // (an odd number of column tiles is padded to whole clusters: the extra CTA loads zeros - TMA fills what lies outside the
        //  tensor - and its epilogue stores nothing)
        if (!share_b) { int rt, cg; raster(cid, RT, (CT + 1) / 2, P.group_m, rt, cg); r_tile = rt; c_tile = cg * 2 + (int)crank; }
        else { int rp, ct; raster(cid, RT / 2, CT, P.group_m, rp, ct); r_tile = rp * 2 + (int)crank; c_tile = ct; }
        const int r0 = r_tile * BM2, c0 = c_tile * BN2;
        const int stride = nar * a_slice + nbr * b_slice;
        int nst = (STAGES2 * STAGE) / stride;
        if (nst > MAXST) nst = MAXST;
        if (warp == 0) {
            if (lane == 0) {
                // my half of the shared role's planes goes to both CTAs, the other role's planes are mine alone
                const int nsh = share_b ? nbr : nar, half = (nsh + 1) / 2;
                const int sh0 = (int)crank * half, sh1 = sh0 + half < nsh ? sh0 + half : nsh;
                int st = 0; uint32_t ph = 0;
                for (int kb = 0; kb < kblocks_t; ++kb) {
                    mbar_wait(&empty[st], ph ^ 1);
                    uint8_t *base = smem + st * stride;
                    if (P.dbg & 1) { mbar_expect_tx(&full[st], 0); if (++st == nst) { st = 0; ph ^= 1; } continue; }
                    mbar_expect_tx(&full[st], stride);
                    if (!share_b) {
                        for (int p = sh0; p < sh1; ++p) tma_load_3d_mc(base + p * a_slice, mapA, &full[st], kb * kbw, r0, p, MASK);
                        for (int q = 0; q < nbr; ++q) tma_load_3d(base + nar * a_slice + q * b_slice, mapB, &full[st], kb * kbw, c0, q);
                    } else {
                        for (int p = 0; p < nar; ++p) tma_load_3d(base + p * a_slice, mapA, &full[st], kb * kbw, r0, p);
                        for (int q = sh0; q < sh1; ++q) tma_load_3d_mc(base + nar * a_slice + q * b_slice, mapB, &full[st], kb * kbw, c0, q, MASK);
                    }
                    if (++st == nst) { st = 0; ph ^= 1; }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                int st = 0; uint32_t ph = 0;
                // one live A-role plane (the common trimmed product): the whole K-block is two merged MMAs per k-step, planes 0..3
                // and 4..6 of the B role - issued from precomputed descriptors (the general loop's index arithmetic, ~100 dependent
                // instructions of this one thread per K-block, took longer than the MMAs themselves)
                const int c1 = nbr < 4 ? nbr : 4, c2 = nbr - c1;
                const uint32_t id1 = make_idesc_i8(c1 * BN2), id2 = make_idesc_i8(c2 * BN2);
                for (int kb = 0; kb < kblocks_t; ++kb) {
                    mbar_wait(&full[st], ph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(smem + st * stride);
                    if (wide) {
                        const uint64_t ad = make_smem_desc(a0), b1 = make_smem_desc(a0 + a_slice), b2 = make_smem_desc(a0 + a_slice + c1 * b_slice);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t accf = (kb > 0 || k > 0) ? 1u : 0u;
                            tc_mma<KIND_I8>(tmem_base, ad + 2 * k, b1 + 2 * k, id1, accf);
                            if (c2) tc_mma<KIND_I8>(tmem_base + 4 * BN2, ad + 2 * k, b2 + 2 * k, id2, accf);
                        }
                    } else if (nar == 1) {
                        const uint64_t ad = make_desc_kb<KB>(a0), b1 = make_desc_kb<KB>(a0 + A_SLICE), b2 = make_desc_kb<KB>(a0 + A_SLICE + c1 * B_SLICE);
#pragma unroll
                        for (int k = 0; k < KB / 32; ++k) {
                            const uint32_t accf = (kb > 0 || k > 0) ? 1u : 0u;
                            tc_mma<KIND_I8>(tmem_base, ad + 2 * k, b1 + 2 * k, id1, accf);
                            if (c2) tc_mma<KIND_I8>(tmem_base + 4 * BN2, ad + 2 * k, b2 + 2 * k, id2, accf);
                        }
                    } else oz2_issue_trimmed<S, KB>(a0, a0 + nar * A_SLICE, tmem_base, nar, nbr, kb == 0);
                    tc_commit_mc(&empty[st], MASK);
                    if (++st == nst) { st = 0; ph ^= 1; }
                }
                tc_commit(tmem_full);
                if (P.pairs && blockIdx.x == 0 && blockIdx.y == 0) atomicAdd(P.pairs, (unsigned long long)oz2_pairs(S, nar, nbr));
            }
        } else {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
            const int nd = nar + nbr - 1 < S ? nar + nbr - 1 : S;
            if (P.dbg & 2) {}
            else if (swap) oz2_epilogue<S, true>(tmem_base, warp & 3, (warp - 2) >> 2, lane, r0, c0, Rr, Cr, P, nd, epi_role(P, true));
            else oz2_epilogue<S, false>(tmem_base, warp & 3, (warp - 2) >> 2, lane, r0, c0, Rr, Cr, P, nd, epi_role(P, false));
        }
    } else if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES2;
                const uint32_t ph = (kb / STAGES2) & 1;
                mbar_wait(&empty[s], ph ^ 1);                     // all CL consumers released this stage (in every CTA)
                if (P.dbg & 1) { mbar_expect_tx(&full[s], 0); continue; }
                if (!dense) {                                     // my share of the live A planes (multicast), my live B planes
                    mbar_expect_tx(&full[s], na * A_SLICE + nb * B_SLICE);
                    const int p1 = ((int)crank + 1) * SL < na ? ((int)crank + 1) * SL : na;
                    for (int p = (int)crank * SL; p < p1; ++p) tma_load_3d_mc(smem + s * STAGE + p * A_SLICE, &tmA1, &full[s], kb * KB, m0, p, MASK);
                    for (int q = 0; q < nb; ++q) tma_load_3d(smem + s * STAGE + SA * A_SLICE + q * B_SLICE, &tmB1, &full[s], kb * KB, n0, q);
                    continue;
                }
                mbar_expect_tx(&full[s], STAGE);                  // bytes landing in MY smem: S A-slices (CL multicasts) + my B slices
                tma_load_3d_mc(smem + s * STAGE + crank * SL * A_SLICE, &tmAmc, &full[s], kb * KB, m0, crank * SL, MASK);
                tma_load_3d(smem + s * STAGE + SA * A_SLICE, &tmB, &full[s], kb * KB, n0, 0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES2;
                const uint32_t ph = (kb / STAGES2) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + s * STAGE), b0 = a0 + SA * A_SLICE;
                if (!dense) { oz2_issue_trimmed<S, KB>(a0, b0, tmem_base, na, nb, kb == 0); tc_commit_mc(&empty[s], MASK); continue; }
#pragma unroll
                for (int p = 0; p < S; ++p) {
                    const uint64_t ad = make_desc_kb<KB>(a0 + p * A_SLICE);
#pragma unroll
                    for (int q0 = 0; q0 < S - p; q0 += 4) {
                        const int nq = (S - p - q0) < 4 ? (S - p - q0) : 4;
                        const uint64_t bd = make_desc_kb<KB>(b0 + q0 * B_SLICE);
                        const uint32_t acc = tmem_base + (uint32_t)((p + q0) * BN2);
                        const uint32_t idesc = make_idesc_i8(nq * BN2);
#pragma unroll
                        for (int k = 0; k < KB / 32; ++k)
                            tc_mma<KIND_I8>(acc, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || p > 0 || k > 0) ? 1u : 0u);
                    }
                }
                tc_commit_mc(&empty[s], MASK);                    // arrives on empty[s] of every CTA in the cluster
            }
            tc_commit(tmem_full);
            if (P.pairs && blockIdx.x == 0 && blockIdx.y == 0) atomicAdd(P.pairs, (unsigned long long)oz2_pairs(S, na, nb));
        }
    } else {
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (!(P.dbg & 2)) oz2_epilogue<S>(tmem_base, warp & 3, (warp - 2) >> 2, lane, m0, n0, M, N, P, dense ? S : (na + nb - 1 < S ? na + nb - 1 : S), epi_role(P));
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                           // no CTA exits while a peer may still multicast into it
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS));
    }
}

// --------------------------------------------------------------------------------------------------
// PERSISTENT cluster variant: one 2-CTA cluster per SM pair for the whole launch, tiles dealt round-robin over the clusters in raster
// order.  TMEM, barriers and tensor-map fetches are set up once; the stage ring runs on across tile boundaries, so while the 16
// epilogue warps drain tile t the producer already fills both stages of tile t+1 (the accumulators - 448 of the 512 TMEM columns -
// cannot be double-buffered, so the MMA thread itself waits for `tmem_empty`).  Per tile this removes TMEM allocation, barrier
// initialisation, two cluster barriers and the exposed latency of the first loads: ~3 us of the ~87 us a 4096-deep tile takes, ~3 us
// of the ~30 us of the MLP config's 1024-deep tiles.
template <int S>
__global__ void __launch_bounds__(OZ_THREADS, 1)
umma_ozaki2cp_kernel(const __grid_constant__ CUtensorMap tmAmc, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmA1,
                     const __grid_constant__ CUtensorMap tmB1, int M, int N, int K, OzParams P, int m_tiles, int n_groups) {
    constexpr int CL = 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int KB = 64, STAGES2 = stages_for(KB), A_SLICE = BM2 * KB, B_SLICE = BN2 * KB;
    constexpr int SL = (S + CL - 1) / CL, SA = SL * CL;
    constexpr int STAGE = SA * A_SLICE + S * B_SLICE;
    constexpr uint16_t MASK = (uint16_t)((1u << CL) - 1u);
    uint64_t *full = (uint64_t *)(smem + STAGES2 * STAGE);
    uint64_t *empty = full + STAGES2;
    uint64_t *tmem_full = empty + STAGES2;
    uint64_t *tmem_empty = tmem_full + 1;
    uint32_t *tmem_ptr = (uint32_t *)(tmem_empty + 1);
    constexpr int TCOLS = 512;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const int cid = (int)(blockIdx.x / CL), nclusters = (int)(gridDim.x / CL);
    const int items = m_tiles * n_groups;
    const int kblocks = (K + KB - 1) / KB;
    const int na = planes_of(P.pa, S, P.bits), nb = planes_of(P.pb, S, P.bits);     // live planes: tmA1 / tmB1 load them one by one
    const bool dense = na == S && nb == S;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, EPI_CB * 4);                        // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t kc = 0;                                       // k-blocks issued so far, over all tiles: the ring never restarts
            for (int it = cid; it < items; it += nclusters) {
                int m_tile, n_group;
                raster(it, m_tiles, n_groups, P.group_m, m_tile, n_group);
                const int m0 = m_tile * BM2, n0 = (n_group * CL + (int)crank) * BN2;
                for (int kb = 0; kb < kblocks; ++kb, ++kc) {
                    const int s = (int)(kc % STAGES2);
                    const uint32_t ph = (kc / STAGES2) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (!dense) {
                        mbar_expect_tx(&full[s], na * A_SLICE + nb * B_SLICE);
                        const int p1 = ((int)crank + 1) * SL < na ? ((int)crank + 1) * SL : na;
                        for (int p = (int)crank * SL; p < p1; ++p) tma_load_3d_mc(smem + s * STAGE + p * A_SLICE, &tmA1, &full[s], kb * KB, m0, p, MASK);
                        for (int q = 0; q < nb; ++q) tma_load_3d(smem + s * STAGE + SA * A_SLICE + q * B_SLICE, &tmB1, &full[s], kb * KB, n0, q);
                        continue;
                    }
                    mbar_expect_tx(&full[s], STAGE);
                    tma_load_3d_mc(smem + s * STAGE + crank * SL * A_SLICE, &tmAmc, &full[s], kb * KB, m0, crank * SL, MASK);
                    tma_load_3d(smem + s * STAGE + SA * A_SLICE, &tmB, &full[s], kb * KB, n0, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t kc = 0, tile = 0;
            for (int it = cid; it < items; it += nclusters, ++tile) {
                if (tile > 0) { mbar_wait(tmem_empty, (tile - 1) & 1); tc_fence_after(); }   // the epilogue has read the previous tile out of TMEM
                for (int kb = 0; kb < kblocks; ++kb, ++kc) {
                    const int s = (int)(kc % STAGES2);
                    const uint32_t ph = (kc / STAGES2) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(smem + s * STAGE), b0 = a0 + SA * A_SLICE;
                    if (!dense) { oz2_issue_trimmed<S, KB>(a0, b0, tmem_base, na, nb, kb == 0); tc_commit_mc(&empty[s], MASK); continue; }
#pragma unroll
                    for (int p = 0; p < S; ++p) {
                        const uint64_t ad = make_desc_kb<KB>(a0 + p * A_SLICE);
#pragma unroll
                        for (int q0 = 0; q0 < S - p; q0 += 4) {
                            const int nq = (S - p - q0) < 4 ? (S - p - q0) : 4;
                            const uint64_t bd = make_desc_kb<KB>(b0 + q0 * B_SLICE);
                            const uint32_t acc = tmem_base + (uint32_t)((p + q0) * BN2);
                            const uint32_t idesc = make_idesc_i8(nq * BN2);
#pragma unroll
                            for (int k = 0; k < KB / 32; ++k)
                                tc_mma<KIND_I8>(acc, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || p > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    tc_commit_mc(&empty[s], MASK);
                }
                tc_commit(tmem_full);
            }
            if (P.pairs && blockIdx.x == 0) atomicAdd(P.pairs, (unsigned long long)oz2_pairs(S, na, nb));
        }
    } else {
        uint32_t tile = 0;
        for (int it = cid; it < items; it += nclusters, ++tile) {
            int m_tile, n_group;
            raster(it, m_tiles, n_groups, P.group_m, m_tile, n_group);
            const int m0 = m_tile * BM2, n0 = (n_group * CL + (int)crank) * BN2;
            mbar_wait(tmem_full, tile & 1);
            tc_fence_after();
            oz2_epilogue<S>(tmem_base, warp & 3, (warp - 2) >> 2, lane, m0, n0, M, N, P, dense ? S : (na + nb - 1 < S ? na + nb - 1 : S), epi_role(P));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive1(tmem_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS));
    }
}

// --------------------------------------------------------------------------------------------------
// CTA-PAIR variant (cta_group::2).  The cluster kernel above runs its main loop at ~79 % of the int8 pipe: per 64-byte K-block
// the MMAs read 192 KB of operands from shared memory and TMA writes 92 KB into it - 2.27 k clk of a 128 B/clk port against
// 1.79 k clk of tensor work (K-sweep in profiles/README.md: 0.27 ms per 1024 of K at 4096 x 4096).  Here two CTAs with ADJACENT
// m-blocks issue ONE tcgen05.mma with M = 256: each CTA keeps its own 128 rows of every A slice (no multicast, no zero slot) and
// supplies only HALF of the N rows of every B operand, which halves the B reads (112 -> 56 KB) and needs no duplicate of A.
//
// The two halves of a B operand must sit at the SAME shared-memory offset in both CTAs (one descriptor serves the pair), the
// leader's rows first in D.  With the merged-N trick (one MMA covers the B slices q0 .. q0+nq-1 of one A slice p, accumulators
// d = p+q contiguous in TMEM) that fixes what each CTA holds per group:
//      group            leader (rank 0)      peer (rank 1)        used for (p)            D
//      {0,1,2,3} N=256  s0 s1                s2 s3                0 1 2 3                 acc p   .. p+3
//      {4,5}     N=128  s4                   s5                   0 1                     acc p+4 .. p+5
//      {0,1}     N=128  s0 (second copy)     s1                   4 5                     acc p   .. p+1
//      {6} {4} {2} {0}  N=64: rows 0..31     rows 32..63          0 2 4 6 respectively    acc 6
// i.e. 6 slice-equivalents of B per CTA and stage (24 KB, 28 before) and 12 MMAs per 32-byte k-step (10 before):
// A reads 96 KB + B reads 56 KB + TMA writes 80 KB = 232 KB = 1.81 k clk per K-block - level with the 1.79 k clk of tensor work.
// Every one of the 28 slice pairs with p + q <= 6 is issued exactly once: nothing is wasted, no scratch accumulator is needed.
// TMA: both CTAs load with .cta_group::2 and signal the LEADER's full barrier (peer bit of the barrier address cleared); the
// leader's tcgen05.commit arrives on the empty barrier of both CTAs.  Only S = 7 x 8-bit digits (the default) is built this way.
__device__ __forceinline__ void tma_load_3d_2sm(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_mma_i8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit_2sm_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ constexpr uint32_t make_idesc_i8_pair(int n) {          // M = 256 across the pair
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
constexpr int PAIR_A_SLICE = BM2 * 64, PAIR_B_SLICE = BN2 * 64;                     // 8 KB, 4 KB
constexpr int PAIR_STAGE = 7 * PAIR_A_SLICE + 6 * PAIR_B_SLICE;                     // 56 KB + 24 KB
__host__ __device__ constexpr int smem_bytes_pair() { return 2 * PAIR_STAGE + 1024 + 256; }

__global__ void __launch_bounds__(OZ_THREADS, 1)
umma_ozaki2p_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmB1,
                    const __grid_constant__ CUtensorMap tmBh, int M, int N, int K, OzParams P) {
    constexpr int S = 7, KB = 64, STAGES2 = 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full = (uint64_t *)(smem + STAGES2 * PAIR_STAGE);
    uint64_t *empty = full + STAGES2;
    uint64_t *tmem_full = empty + STAGES2;
    uint32_t *tmem_ptr = (uint32_t *)(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();                     // 0 = leader: issues every MMA of the pair
    int pair_m, n_tile;
    raster((int)(blockIdx.x / 2 + (gridDim.x / 2) * blockIdx.y), (int)(gridDim.x / 2), (int)gridDim.y, P.group_m > 1 ? P.group_m / 2 : 1, pair_m, n_tile);
    const int m0 = (pair_m * 2 + (int)crank) * BM2, n0 = n_tile * BN2;
    const int kblocks = (K + KB - 1) / KB;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                              // one warp of EACH CTA, same smem slot for the address
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                           // both CTAs' barriers exist before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES2;
                const uint32_t ph = (kb / STAGES2) & 1;
                mbar_wait(&empty[s], ph ^ 1);                     // the leader's commit released this stage in both CTAs
                if (P.dbg & 1) { if (crank == 0) mbar_expect_tx(&full[s], 0); continue; }
                if (crank == 0) mbar_expect_tx(&full[s], 2 * PAIR_STAGE);     // the bytes of BOTH CTAs land on the leader's barrier
                uint8_t *st = smem + s * PAIR_STAGE, *b = st + 7 * PAIR_A_SLICE;
                const int k0 = kb * KB;
                tma_load_3d_2sm(st, &tmA, &full[s], k0, m0, 0);                                    // my 128 rows of all 7 A slices
                const int r = (int)crank;
                tma_load_3d_2sm(b, &tmB2, &full[s], k0, n0, 2 * r);                                // {0,1,2,3}: s0 s1 | s2 s3
                tma_load_3d_2sm(b + 2 * PAIR_B_SLICE, &tmB1, &full[s], k0, n0, 4 + r);             // {4,5}:     s4    | s5
                tma_load_3d_2sm(b + 3 * PAIR_B_SLICE, &tmB1, &full[s], k0, n0, r);                 // {0,1}:     s0    | s1
#pragma unroll
                for (int h = 0; h < 4; ++h)                                                          // {6} {4} {2} {0}: rows 0..31 | 32..63
                    tma_load_3d_2sm(b + 4 * PAIR_B_SLICE + h * (PAIR_B_SLICE / 2), &tmBh, &full[s], k0, n0 + 32 * r, 6 - 2 * h);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES2;
                const uint32_t ph = (kb / STAGES2) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + s * PAIR_STAGE), b0 = a0 + 7 * PAIR_A_SLICE;
                const uint32_t bG4 = b0, bG45 = b0 + 2 * PAIR_B_SLICE, bG01 = b0 + 3 * PAIR_B_SLICE, bH = b0 + 4 * PAIR_B_SLICE;
                auto acc = [&](int d) { return tmem_base + (uint32_t)(d * BN2); };
#pragma unroll
                for (int k = 0; k < KB / 32; ++k) {
                    const uint32_t first = (kb > 0 || k > 0) ? 1u : 0u;           // p = 0 touches every accumulator first
                    auto A = [&](int p) { return make_desc_kb<KB>(a0 + p * PAIR_A_SLICE) + 2 * k; };
                    auto B = [&](uint32_t addr) { return make_desc_kb<KB>(addr) + 2 * k; };
                    // p = 0
                    tc_mma_i8_2sm(acc(0), A(0), B(bG4), make_idesc_i8_pair(256), first);
                    tc_mma_i8_2sm(acc(4), A(0), B(bG45), make_idesc_i8_pair(128), first);
                    tc_mma_i8_2sm(acc(6), A(0), B(bH), make_idesc_i8_pair(64), first);
                    // p = 1
                    tc_mma_i8_2sm(acc(1), A(1), B(bG4), make_idesc_i8_pair(256), 1u);
                    tc_mma_i8_2sm(acc(5), A(1), B(bG45), make_idesc_i8_pair(128), 1u);
                    // p = 2
                    tc_mma_i8_2sm(acc(2), A(2), B(bG4), make_idesc_i8_pair(256), 1u);
                    tc_mma_i8_2sm(acc(6), A(2), B(bH + PAIR_B_SLICE / 2), make_idesc_i8_pair(64), 1u);
                    // p = 3
                    tc_mma_i8_2sm(acc(3), A(3), B(bG4), make_idesc_i8_pair(256), 1u);
                    // p = 4
                    tc_mma_i8_2sm(acc(4), A(4), B(bG01), make_idesc_i8_pair(128), 1u);
                    tc_mma_i8_2sm(acc(6), A(4), B(bH + 2 * (PAIR_B_SLICE / 2)), make_idesc_i8_pair(64), 1u);
                    // p = 5
                    tc_mma_i8_2sm(acc(5), A(5), B(bG01), make_idesc_i8_pair(128), 1u);
                    // p = 6
                    tc_mma_i8_2sm(acc(6), A(6), B(bH + 3 * (PAIR_B_SLICE / 2)), make_idesc_i8_pair(64), 1u);
                }
                tc_commit_2sm_mc(&empty[s], 0b11);                // frees the stage in both CTAs when these MMAs retire
            }
            tc_commit_2sm_mc(tmem_full, 0b11);                    // accumulators complete in both CTAs
        }
    } else {
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (!(P.dbg & 2)) oz2_epilogue<S>(tmem_base, warp & 3, (warp - 2) >> 2, lane, m0, n0, M, N, P, S, epi_role(P));
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                           // neither CTA frees TMEM / exits while the pair is still at work
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// --------------------------------------------------------------------------------------------------
// TRIMMED CTA-PAIR kernel (cta_group::2) for products in which one operand has ONE live digit plane (plane words: the adjoint of
// `sum`, constant fills, one-hot seeds).  ncu on the trimmed cluster path (profiles/prof_ozaki2c_r2b.txt) shows those launches
// bound by operand DELIVERY - 36 KB per 64-byte K-block arrive in every SM for 0.45 k clocks of tensor work, at the same
// ~12 TB/s of crossbar -> shared memory the dense kernel runs at - so multicast cannot help (every SM still receives every byte).
// A CTA pair can: the one-plane operand takes the A role (each CTA its own 128 rows, 8 KB), the other operand's planes are the
// B rows of M = 256 MMAs and each CTA supplies only HALF of them (2 KB per plane instead of 4): 22 KB per K-block for (1, 7)
// planes, and up to 8 stages of it in flight.  The B rows of one MMA are the 64-row planes of a chunk (planes 0..3 -> N = 256,
// planes 4..6 -> N = 192) stacked in plane order, so the leader holds the first 32 c rows of a c-plane chunk and the peer the
// rest: whole planes come as 64-row boxes, the odd half plane as a 32-row box.  Accumulator d = q sits at TMEM column 64 q in both
// CTAs, so the epilogue is the trimmed cluster path's.  tcgen05 instructions of one kernel must agree on the cta_group, hence a
// separate kernel; the host cannot know the plane counts (they are device data), so it launches this kernel in front of the
// cluster kernel and each of the two returns at once when the product is the other's (P.pair_first).
template <int S>
__global__ void __launch_bounds__(OZ_THREADS, 1)
umma_ozaki2tp_kernel(const __grid_constant__ CUtensorMap tmA128, const __grid_constant__ CUtensorMap tmA64, const __grid_constant__ CUtensorMap tmA32,
                     const __grid_constant__ CUtensorMap tmB128, const __grid_constant__ CUtensorMap tmB64, const __grid_constant__ CUtensorMap tmB32,
                     int M, int N, int K, OzParams P) {
    const int na = planes_of(P.pa, S, P.bits), nb = planes_of(P.pb, S, P.bits);
    if ((na < nb ? na : nb) != 1) return;                         // not a one-plane product: the cluster kernel behind this launch runs it
    constexpr int KB = 64, A_SLICE = BM2 * KB, HALF = (BN2 / 2) * KB, MAXST = 8, RING = 2 * (8 * A_SLICE + S * BN2 * KB);
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full = (uint64_t *)(smem + RING);
    uint64_t *empty = full + MAXST;
    uint64_t *tmem_full = empty + MAXST;
    uint32_t *tmem_ptr = (uint32_t *)(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();                     // 0 = leader: issues every MMA of the pair
    const bool swap = na != 1;                                     // op(B) is the one-plane operand: compute the tile of C^T
    const int nbr = swap ? na : nb;
    const int Rr = swap ? N : M, Cr = swap ? M : N;
    const CUtensorMap *mapA = swap ? &tmB128 : &tmA128, *mapB64 = swap ? &tmA64 : &tmB64, *mapB32 = swap ? &tmA32 : &tmB32;
    int rp, ct;
    raster((int)(blockIdx.x / 2), Rr / (2 * BM2), Cr / BN2, P.group_m > 1 ? P.group_m / 2 : 1, rp, ct);
    const int r0 = (rp * 2 + (int)crank) * BM2, c0 = ct * BN2;
    const int kblocks = (K + KB - 1) / KB;
    const int c1 = nbr < 4 ? nbr : 4, c2 = nbr - c1;              // planes per chunk: N = 64 c1 and 64 c2
    const int stride = A_SLICE + nbr * HALF;
    int nst = RING / stride;
    if (nst > MAXST) nst = MAXST;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < MAXST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty[st], ph ^ 1);                    // the leader's commit released this stage in both CTAs
                if (crank == 0) mbar_expect_tx(&full[st], (P.dbg & 1) ? 0 : 2 * stride);      // the bytes of BOTH CTAs land on the leader's barrier
                if (!(P.dbg & 1)) {
                    uint8_t *base = smem + st * stride;
                    const int k0 = kb * KB;
                    tma_load_3d_2sm(base, mapA, &full[st], k0, r0, 0);
                    uint8_t *dst = base + A_SLICE;
                    for (int chunk = 0; chunk < 2; ++chunk) {      // my half of each chunk's rows, in units of 32-row half planes
                        const int c = chunk ? c2 : c1, plane0 = chunk ? 4 : 0;
                        for (int h = (int)crank * c, h1 = h + c; h < h1;) {
                            if (!(h & 1) && h + 1 < h1) { tma_load_3d_2sm(dst, mapB64, &full[st], k0, c0, plane0 + (h >> 1)); dst += 2 * HALF; h += 2; }
                            else { tma_load_3d_2sm(dst, mapB32, &full[st], k0, c0 + 32 * (h & 1), plane0 + (h >> 1)); dst += HALF; h += 1; }
                        }
                    }
                }
                if (++st == nst) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) {
            int st = 0; uint32_t ph = 0;
            const uint32_t id1 = make_idesc_i8_pair(c1 * BN2), id2 = make_idesc_i8_pair(c2 * BN2);
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&full[st], ph);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + st * stride);
                const uint64_t ad = make_desc_kb<KB>(a0), b1 = make_desc_kb<KB>(a0 + A_SLICE), b2 = make_desc_kb<KB>(a0 + A_SLICE + c1 * HALF);
#pragma unroll
                for (int k = 0; k < KB / 32; ++k) {
                    const uint32_t accf = (kb > 0 || k > 0) ? 1u : 0u;
                    tc_mma_i8_2sm(tmem_base, ad + 2 * k, b1 + 2 * k, id1, accf);
                    if (c2) tc_mma_i8_2sm(tmem_base + 4 * BN2, ad + 2 * k, b2 + 2 * k, id2, accf);
                }
                tc_commit_2sm_mc(&empty[st], 0b11);               // frees the stage in both CTAs when these MMAs retire
                if (++st == nst) { st = 0; ph ^= 1; }
            }
            tc_commit_2sm_mc(tmem_full, 0b11);
            if (P.pairs && blockIdx.x == 0) atomicAdd(P.pairs, (unsigned long long)nbr);
        }
    } else {
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (P.dbg & 2) {}
        else if (swap) oz2_epilogue<S, true>(tmem_base, warp & 3, (warp - 2) >> 2, lane, r0, c0, Rr, Cr, P, nbr, epi_role(P, true));
        else oz2_epilogue<S, false>(tmem_base, warp & 3, (warp - 2) >> 2, lane, r0, c0, Rr, Cr, P, nbr, epi_role(P, false));
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                           // neither CTA frees TMEM / exits while the pair is still at work
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}
}  // namespace oz2

}  // namespace umma

bool umma_available() { return umma::encode_fn() != nullptr; }

// C(MxN) <- alpha*op(A)*op(B) + beta*C in FP32 via 3xTF32 on tcgen05.  Returns cudaErrorNotSupported when the problem is
// too small to be worth it (caller uses the SIMT kernel).
cudaError_t gemm_f32_umma(cudaStream_t st, bool ta, bool tb, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                          int64_t lda, const float *B, int64_t ldb, float beta, float *C, int64_t ldc) {
    using namespace umma;
    if (M < 128 || N < 128 || K < 32 || !umma_available()) return cudaErrorNotSupported;
    const int64_t Kp = (K + 3) & ~3LL;                       // 16-byte rows for TMA
    const size_t a_bytes = (size_t)2 * M * Kp * 4, b_bytes = (size_t)2 * N * Kp * 4;
    uint8_t *ws = (uint8_t *)workspace(((a_bytes + 255) & ~(size_t)255) + b_bytes);
    if (!ws) return cudaErrorMemoryAllocation;
    float *As = (float *)ws, *Bs = (float *)(ws + ((a_bytes + 255) & ~(size_t)255));
    // op(A) is MxK: ta => stored KxM (k contiguous per row m) => KMAJOR source.  op(B)^T rows are n: !tb => stored KxN => KMAJOR.
    dim3 ga((unsigned)((Kp + 31) / 32), (unsigned)((M + 31) / 32)), gb((unsigned)((Kp + 31) / 32), (unsigned)((N + 31) / 32));
    if (ta) k_split_tf32<true><<<ga, 256, 0, st>>>(As, A, lda, (int)M, (int)K, (int)Kp);
    else k_split_tf32<false><<<ga, 256, 0, st>>>(As, A, lda, (int)M, (int)K, (int)Kp);
    if (!tb) k_split_tf32<true><<<gb, 256, 0, st>>>(Bs, B, ldb, (int)N, (int)K, (int)Kp);
    else k_split_tf32<false><<<gb, 256, 0, st>>>(Bs, B, ldb, (int)N, (int)K, (int)Kp);
    CUtensorMap mA, mB;
    if (!make_map(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, As, Kp, M, 2, BM) ||
        !make_map(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Bs, Kp, N, 2, BN))
        return cudaErrorInvalidValue;
    Segs segs;
    segs.n = 3;                      // small terms first: lo*hi, hi*lo, then hi*hi
    segs.sa[0] = 1; segs.sb[0] = 0;
    segs.sa[1] = 0; segs.sb[1] = 1;
    segs.sa[2] = 0; segs.sb[2] = 0;
    EpiF32 epi{C, ldc, alpha, beta};
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(umma_gemm_kernel<KIND_TF32, EpiF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
    umma_gemm_kernel<KIND_TF32, EpiF32><<<grid, THREADS_F32, SMEM_BYTES, st>>>(mA, mB, (int)M, (int)N, (int)Kp, segs, epi);
    g_gemm_launches = 3;
    return cudaGetLastError();
}


// bits per int8 digit: 8 (default) uses the whole int8 range, so 7 slices carry the 56 fixed-point bits that took 8 slices of 7
// bits - 28 slice products instead of 36 for the same truncation bound; int32 accumulators stay exact for K*S < 2^17
int ozaki_bits() {
    static int b = -1;
    if (b < 0) { const char *e = getenv("RDB_OZAKI_BITS"); b = (e && atoi(e) == 7) ? 7 : 8; }
    return b;
}
int ozaki_default_slices() {
    static int n = -1;
    if (n < 0) {
        const char *e = getenv("RDB_OZAKI_SLICES");
        n = e ? atoi(e) : (ozaki_bits() == 8 ? 7 : 8);
        if (n < 2) n = 2;
        if (n > umma::OZ_MAX_SLICES) n = umma::OZ_MAX_SLICES;
    }
    return n;
}

// One K-chunk of the sliced product (K*nslices < 2^19 keeps every int32 accumulation chain exact).
static cudaError_t ozaki_chunk(cudaStream_t st, bool ta, bool tb, int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                               int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc, int nslices,
                               uint64_t tag_a, uint64_t tag_b, OperandWindow wa, OperandWindow wb);
static thread_local int g_chunk_launches = 0;

// C(MxN) <- alpha*op(A)*op(B) + beta*C in FP64 through exact int8 slices on tcgen05 (see the comment block above).
// Long contractions are cut into K-chunks (each sliced with its own row/column scales, accumulated into C with beta = 1): this
// keeps int32 exact and gives small-MxN / huge-K adjoints (dW = dZ * X', K = 65536 in the MLP config) enough tiles per wave.
cudaError_t gemm_f64_ozaki(cudaStream_t st, bool ta, bool tb, int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                           int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc, int nslices,
                           uint64_t tag_a, uint64_t tag_b, OperandWindow wa, OperandWindow wb) {
    if (M < 256 || N < 256 || K < 256 || !umma_available()) return cudaErrorNotSupported;
    const int64_t kmax = (((1 << (33 - 2 * ozaki_bits())) - 1) / nslices) & ~127LL;      // int32 exactness, see ozaki_chunk
    int64_t chunks = (K + kmax - 1) / kmax;
    const int64_t tiles = ((M + 127) / 128) * ((N + 63) / 64);
    (void)tiles;
    static const int split_more = [] { const char *e = getenv("RDB_OZAKI_SPLITK"); return e ? atoi(e) : 0; }();
    if (split_more) while (tiles < 592 && K / (chunks * 2) >= 8192) chunks *= 2;   // (experiment: more, shorter chunks; they run one after the other)
    chunks = 1 > chunks ? 1 : chunks;
    const int64_t kc = (((K + chunks - 1) / chunks) + 127) & ~127LL;
    int launches = 0;
    for (int64_t k0 = 0, c = 0; k0 < K; k0 += kc, ++c) {
        const int64_t kk = K - k0 < kc ? K - k0 : kc;
        const double *Ac = ta ? A + k0 : A + k0 * lda;                   // op(A)(:, k0:) : stored KxM (ta) or MxK
        const double *Bc = tb ? B + k0 * ldb : B + k0;                   // op(B)(k0:, :) : stored NxK (tb) or KxN
        // (windows are honoured for single-chunk products only; with several K-chunks a windowed operand is sliced per panel,
        //  untagged, because its tag names the whole operand)
        const bool one = chunks == 1;
        cudaError_t e = ozaki_chunk(st, ta, tb, M, N, kk, alpha, Ac, lda, Bc, ldb, c == 0 ? beta : 1.0, C, ldc, nslices,
                                    (!one && wa.full) ? 0 : tag_a, (!one && wb.full) ? 0 : tag_b, one ? wa : OperandWindow(), one ? wb : OperandWindow());
        if (e != cudaSuccess) return e;
        launches += g_chunk_launches;      // per uncached operand: absmax, exponent, slice; + 1 tcgen05 kernel
    }
    g_gemm_launches = launches;
    return cudaSuccess;
}

cudaError_t guard_fold(cudaStream_t st, GemmState &G) {
    if (!G.guard || G.guard_used == 0) return cudaSuccess;
    umma::k_guard_fold<<<(G.guard_used + 255) / 256, 256, 0, st>>>(G.guard, G.guard_used, G.guard_sum);
    G.guard_used = 0;
    G.guard_pending = true;
    return cudaGetLastError();
}
cudaError_t guard_read(cudaStream_t st, GemmState &G, GuardSummary *out) {
    *out = GuardSummary{0, 0, 0.0, 0ull};
    if (!G.guard) return cudaSuccess;
    cudaError_t e = guard_fold(st, G);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(G.guard_host, G.guard_sum, sizeof(GuardSummary), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(G.guard_sum, 0, sizeof(GuardSummary), st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    *out = *G.guard_host;
    G.guard_pending = false;
    return cudaSuccess;
}

// ---- sliced-operand cache: CONST matrices (frozen by `capture`) and value buffers are sliced once per content version and
// reused by every `*` instruction / adjoint / seed block that reads them (the jacobian! config re-reads A and B for every block).
static size_t cache_cap() {
    static size_t cap = 0;
    if (!cap) { const char *e = getenv("RDB_SLICE_CACHE_MB"); cap = (size_t)(e ? atoll(e) : 8192) << 20; }
    return cap;
}

int ozaki_bits();
static int slice_operand(cudaStream_t st, const double *src, int64_t ld, int64_t rows, int64_t K, int64_t Kp, bool kmajor, int ns,
                         int8_t *q, int *e, double *sc) {      // returns the number of kernels launched
    const int bits = ozaki_bits();
    using namespace umma;
    static const bool uniform_on = [] { const char *e = getenv("RDB_SLICE_UNIFORM"); return !(e && e[0] == '0'); }();
    if (uniform_on && gemm_state().uniform_has(src) && Kp % 16 == 0 && ((uintptr_t)q & 15) == 0) {
        // every element of this operand equals src[0] (the engine's word for it): planes as fills, no read of the matrix
        k_slice_uniform<<<148 * 4, 256, 0, st>>>(q, src, (int)rows, (int)K, (int)Kp, ns, bits, sc);
        return 1;
    }
    // k-contiguous rows that fit shared memory and are 16-byte aligned: the one-pass kernel
    static const bool fused_on = [] { const char *e = getenv("RDB_SLICE_FUSED"); return !(e && e[0] == '0'); }();
    // (measured: a win while four rows fit 64 KB so several CTAs share an SM - the MLP config's K = 1024 operands, 12.04 -> 11.85 ms;
    // at K = 4096, one 128 KB CTA per SM, it loses to the two-pass kernels)
    if (kmajor && fused_on && K % 2 == 0 && ld % 2 == 0 && (((uintptr_t)src) & 15) == 0 && Kp <= 2048) {
        int wpc = 4;
        const size_t smem = (size_t)wpc * Kp * 8;
        static size_t attr_smem = 0;
        if (smem > attr_smem) {
            if (cudaFuncSetAttribute(k_slice_fused_kmajor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)) == cudaSuccess) attr_smem = 200 * 1024;
        }
        if (smem <= attr_smem || smem <= 48 * 1024) {
            cudaMemsetAsync(sc + 3 * rows, 0, 8, st);                        // plane word
            k_slice_fused_kmajor<<<(unsigned)((rows + wpc - 1) / wpc), 32 * wpc, smem, st>>>(q, src, ld, (int)rows, (int)K, (int)Kp, ns, bits, sc);
            return 1;
        }
    }
    unsigned long long *mx = (unsigned long long *)sc;                  // scale array doubles as max scratch
    if (kmajor) k_row_absmax<true><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(src, ld, (int)rows, (int)K, mx);
    else {
        cudaMemsetAsync(mx, 0, rows * 24, st);                           // max | 1-norm | nnz
        k_row_absmax<false><<<dim3((unsigned)((rows + 255) / 256), (unsigned)((K + 63) / 64)), 256, 0, st>>>(src, ld, (int)rows, (int)K, mx);
    }
    k_row_exponent<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(mx, (int)rows, bits, e, sc);
    dim3 g((unsigned)((Kp + 127) / 128), (unsigned)((rows + 31) / 32));
    unsigned long long *pw = reinterpret_cast<unsigned long long *>(sc) + 3 * rows;    // zeroed by k_row_exponent
    const bool s7 = bits == 8 && ns == 7;                              // the default digits: specialised kernels
    if (kmajor) {
        if (s7) k_slice_i8<true, 7><<<g, 256, 0, st>>>(q, src, ld, (int)rows, (int)K, (int)Kp, ns, bits, e, pw);
        else k_slice_i8<true><<<g, 256, 0, st>>>(q, src, ld, (int)rows, (int)K, (int)Kp, ns, bits, e, pw);
    } else {
        dim3 gc((unsigned)((Kp + 31) / 32), (unsigned)((rows + 127) / 128));
        if (gc.y <= 65535) {
            if (s7) k_slice_i8_cm<7><<<gc, 256, 0, st>>>(q, src, ld, (int)rows, (int)K, (int)Kp, ns, bits, e, pw);
            else k_slice_i8_cm<><<<gc, 256, 0, st>>>(q, src, ld, (int)rows, (int)K, (int)Kp, ns, bits, e, pw);
        } else k_slice_i8<false><<<g, 256, 0, st>>>(q, src, ld, (int)rows, (int)K, (int)Kp, ns, bits, e, pw);
    }
    return 3;
}

// returns the slices/scales of one operand; launches = kernels issued (0 on a cache hit)
static cudaError_t prepare_operand(cudaStream_t st, uint64_t tag, const double *src, int64_t ld, int64_t rows, int64_t K, int64_t Kp,
                                   bool kmajor, int ns, int8_t *ws_q, int *ws_e, double *ws_sc, int8_t **q, double **sc, int *launches) {
    GemmState &G = gemm_state();
    auto &g_cache = G.cache;
    uint64_t &g_use = G.use;
    size_t &g_cache_bytes = G.cache_bytes;
    if (!tag || G.no_cache) {
        if (!ws_q) { *q = nullptr; *sc = nullptr; *launches = 0; return cudaSuccess; }     // slicing ahead: nothing to do
        *launches = slice_operand(st, src, ld, rows, K, Kp, kmajor, ns, ws_q, ws_e, ws_sc);
        *q = ws_q; *sc = ws_sc;
        return cudaSuccess;
    }
    SliceEntry *slot = nullptr;
    for (auto &en : g_cache)
        if (en.ptr == src && en.ld == ld && en.rows == rows && en.K == K && en.kmajor == kmajor && en.ns == ns && en.bits == ozaki_bits()) { slot = &en; break; }
    if (slot && slot->tag == tag) {
        slot->last_use = ++g_use;
        if (slot->filled_on != st && slot->ready) cudaStreamWaitEvent(st, slot->ready, 0);   // sliced ahead of time on the side stream
        *q = slot->q; *sc = slot->sc; *launches = 0;
        return cudaSuccess;
    }
    if (!slot) {
        const size_t bytes = (size_t)ns * rows * Kp + (size_t)rows * 24 + 8 + 512; // slices + (scale | 1-norm | nnz) per row + plane word
        while (!g_cache.empty() && g_cache_bytes + bytes > cache_cap()) {          // evict least recently used
            size_t lru = 0;
            for (size_t i = 1; i < g_cache.size(); ++i) if (g_cache[i].last_use < g_cache[lru].last_use) lru = i;
            cudaFree(g_cache[lru].q);
            if (g_cache[lru].ready) cudaEventDestroy(g_cache[lru].ready);
            g_cache_bytes -= g_cache[lru].bytes;
            g_cache.erase(g_cache.begin() + lru);
        }
        if (bytes > cache_cap()) {                                                  // too big to cache
            if (!ws_q) { *q = nullptr; *sc = nullptr; *launches = 0; return cudaSuccess; }     // slicing ahead: nothing to do
            *launches = slice_operand(st, src, ld, rows, K, Kp, kmajor, ns, ws_q, ws_e, ws_sc);
            *q = ws_q; *sc = ws_sc;
            return cudaSuccess;
        }
        SliceEntry en{};
        if (cudaMalloc((void **)&en.q, bytes) != cudaSuccess) {
            cudaGetLastError();
            if (!ws_q) { *q = nullptr; *sc = nullptr; *launches = 0; return cudaSuccess; }
            *launches = slice_operand(st, src, ld, rows, K, Kp, kmajor, ns, ws_q, ws_e, ws_sc); *q = ws_q; *sc = ws_sc; return cudaSuccess;
        }
        cudaEventCreateWithFlags(&en.ready, cudaEventDisableTiming);
        en.sc = (double *)(en.q + (((size_t)ns * rows * Kp + 255) & ~(size_t)255));
        en.ptr = src; en.ld = ld; en.rows = rows; en.K = K; en.kmajor = kmajor; en.ns = ns; en.bits = ozaki_bits(); en.bytes = bytes;
        g_cache_bytes += bytes;
        g_cache.push_back(en);
        slot = &g_cache.back();
    }
    const int nl_fill = slice_operand(st, src, ld, rows, K, Kp, kmajor, ns, slot->q, ws_e, slot->sc);    // (re)fill in place: same buffers, new version
    if (slot->ready) cudaEventRecord(slot->ready, st);
    slot->filled_on = st;
    slot->tag = tag;
    slot->last_use = ++g_use;
    *q = slot->q; *sc = slot->sc; *launches = nl_fill;
    return cudaSuccess;
}

// Slice an operand AHEAD of the GEMM that will read it, on another stream (the slicing passes are HBM-bound, the GEMM kernels
// tensor-bound and 1 CTA/SM: the two overlap).  The entry lands in the cache under `tag`; the GEMM's own prepare_operand then
// finds it and makes its stream wait for the entry's event.  kmajor: the operand's k index is the contiguous one
// (A with ta, B without tb).  Only what gemm_f64_ozaki would run as ONE exact-int32 chunk is sliced ahead.
cudaError_t ozaki_preslice(cudaStream_t side, const double *src, int64_t ld, int64_t rows, int64_t K, bool kmajor, int nslices, uint64_t tag,
                           int *launched) {
    if (launched) *launched = 0;
    if (!tag || !src || rows < 256 || K < 256 || !umma_available()) return cudaSuccess;
    if (nslices <= 0) nslices = ozaki_default_slices();
    const int64_t kmax = (((1 << (33 - 2 * ozaki_bits())) - 1) / nslices) & ~127LL;
    if (K > kmax) return cudaSuccess;
    GemmState &G = gemm_state();
    if (G.no_cache) return cudaSuccess;
    int *&side_e = G.side_e;
    int64_t &side_e_n = G.side_e_n;
    if (rows > side_e_n) {
        if (side_e) cudaFree(side_e);
        side_e = nullptr; side_e_n = 0;
        if (cudaMalloc((void **)&side_e, (size_t)rows * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return cudaSuccess; }
        side_e_n = rows;
    }
    const int64_t Kp = (K + 15) & ~15LL;
    int8_t *q; double *sc; int launches = 0;
    const cudaError_t e = prepare_operand(side, tag, src, ld, rows, K, Kp, kmajor, nslices, nullptr, side_e, nullptr, &q, &sc, &launches);
    if (launched) *launched = launches;
    return e;
}

static cudaError_t ozaki_chunk(cudaStream_t st, bool ta, bool tb, int64_t M, int64_t N, int64_t K, double alpha, const double *A,
                               int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc, int nslices,
                               uint64_t tag_a, uint64_t tag_b, OperandWindow wa, OperandWindow wb) {
    using namespace umma;
    // int32 exactness: (d+1)*K*2^(2(bits-1)) < 2^31 for d <= nslices-1
    if (K * (int64_t)nslices > (1 << (33 - 2 * ozaki_bits())) - 1) return cudaErrorInvalidValue;
    const int64_t Kp = (K + 15) & ~15LL;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    GemmState &G = gemm_state();
    if (G.no_cache) tag_a = tag_b = 0;         // CUDA-graph tapes: slices always land in the workspace, inside the graph
    if (!tag_a) wa = OperandWindow();
    if (!tag_b) wb = OperandWindow();
    // an operand that is a row window of a larger, tagged operand: slice (or find) the WHOLE operand in the cache and point the
    // tensor map at the window - the panels of a panelled GEMM share one slicing
    const size_t a_bytes = (tag_a && !wa.full) ? 0 : al((size_t)nslices * M * Kp), b_bytes = (tag_b && !wb.full) ? 0 : al((size_t)nslices * N * Kp);
    const size_t ea_bytes = al(std::max<int64_t>(M, wa.rows_full) * 4), eb_bytes = al(std::max<int64_t>(N, wb.rows_full) * 4), sa_bytes = al(M * 24 + 8), sb_bytes = al(N * 24 + 8);
    uint8_t *ws = (uint8_t *)workspace(a_bytes + b_bytes + ea_bytes + eb_bytes + sa_bytes + sb_bytes);
    if (!ws) return cudaErrorMemoryAllocation;
    int8_t *ws_a = (int8_t *)ws, *ws_b = (int8_t *)(ws + a_bytes);
    int *ea = (int *)(ws + a_bytes + b_bytes), *eb = (int *)(ws + a_bytes + b_bytes + ea_bytes);
    double *ws_sa = (double *)(ws + a_bytes + b_bytes + ea_bytes + eb_bytes);
    double *ws_sb = (double *)(ws + a_bytes + b_bytes + ea_bytes + eb_bytes + sa_bytes);
    // rows of op(A) are m: ta => stored KxM, k contiguous per m => KMAJOR.  rows of op(B)^T are n: !tb => stored KxN => KMAJOR.
    int8_t *As, *Bs;
    double *sa, *sb;
    const unsigned long long *pwa = nullptr, *pwb = nullptr;       // plane words: behind the statistics of the operand that was sliced
    int la = 0, lb = 0;
    int64_t a_stride = 0, b_stride = 0;                 // rows between slice planes when the operand is a window
    cudaError_t pe = cudaSuccess;
    As = nullptr; Bs = nullptr;
    if (wa.full) {
        pe = prepare_operand(st, tag_a, (const double *)wa.full, lda, wa.rows_full, K, Kp, ta, nslices, nullptr, ea, nullptr, &As, &sa, &la);
        if (pe != cudaSuccess) return pe;
        if (As) { pwa = reinterpret_cast<const unsigned long long *>(sa) + 3 * wa.rows_full; As += wa.row0 * Kp; sa += wa.row0; a_stride = wa.rows_full; }
    }
    if (!As) {                                         // no window, or the whole operand could not be cached: this panel on its own
        pe = prepare_operand(st, wa.full ? 0 : tag_a, A, lda, M, K, Kp, ta, nslices, ws_a, ea, ws_sa, &As, &sa, &la);
        if (pe != cudaSuccess) return pe;
        pwa = reinterpret_cast<const unsigned long long *>(sa) + 3 * M;
    }
    if (wb.full) {
        pe = prepare_operand(st, tag_b, (const double *)wb.full, ldb, wb.rows_full, K, Kp, !tb, nslices, nullptr, eb, nullptr, &Bs, &sb, &lb);
        if (pe != cudaSuccess) return pe;
        if (Bs) { pwb = reinterpret_cast<const unsigned long long *>(sb) + 3 * wb.rows_full; Bs += wb.row0 * Kp; sb += wb.row0; b_stride = wb.rows_full; }
    }
    if (!Bs) {
        pe = prepare_operand(st, wb.full ? 0 : tag_b, B, ldb, N, K, Kp, !tb, nslices, ws_b, eb, ws_sb, &Bs, &sb, &lb);
        if (pe != cudaSuccess) return pe;
        pwb = reinterpret_cast<const unsigned long long *>(sb) + 3 * N;
    }
    g_chunk_launches = la + lb + 1;
    CUtensorMap mA, mB;
    static int group_m = -1;
    if (group_m < 0) { const char *e = getenv("RDB_OZAKI_GROUP_M"); group_m = e ? atoi(e) : 8; if (group_m < 1) group_m = 1 << 20; }
    // accuracy certificate: one slot per launch; a full table is folded (stream-ordered) and reused
    if (!G.guard) {
        if (cudaMalloc((void **)&G.guard, kGuardSlots * 2 * sizeof(unsigned long long)) != cudaSuccess ||
            cudaMalloc((void **)&G.guard_sum, sizeof(GuardSummary)) != cudaSuccess ||
            cudaMallocHost((void **)&G.guard_host, sizeof(GuardSummary)) != cudaSuccess) return cudaErrorMemoryAllocation;
        cudaMemsetAsync(G.guard, 0, kGuardSlots * 2 * sizeof(unsigned long long), st);
        cudaMemsetAsync(G.guard_sum, 0, sizeof(GuardSummary), st);
        G.guard_used = 0;
    }
    if (G.guard_used >= kGuardSlots) { cudaError_t ge = guard_fold(st, G); if (ge != cudaSuccess) return ge; }
    unsigned long long *gslot = G.guard + 2 * (G.guard_used++);
    double g_rho, g_h, g_drop;
    guard_coeffs(ozaki_bits(), nslices, g_rho, g_h, g_drop);
    // the row statistics sit `rows` apart behind the scales (the rows of the whole operand when this is a window of it)
    const int64_t ra = a_stride ? a_stride : M, rb = b_stride ? b_stride : N;
    static const int dbg = [] { const char *e = getenv("RDB_OZAKI_DBG"); return e ? atoi(e) : 0; }();
    static const bool trim = [] { const char *e = getenv("RDB_OZAKI_TRIM"); return !(e && e[0] == '0'); }();    // 0: every plane of every operand is multiplied
    OzParams P{C, ldc, alpha, beta, sa, sb, nslices, group_m, ozaki_bits(), gslot, sa + ra, sa + 2 * ra, sb + rb, sb + 2 * rb, g_rho, g_h, g_drop, dbg,
               trim ? pwa : nullptr, trim ? pwb : nullptr, &G.guard_sum->pairs,
               getenv("RDB_OZAKI_NOSWAP") ? 0 : 1, 0,
               getenv("RDB_OZAKI_WIDE") ? atoi(getenv("RDB_OZAKI_WIDE")) : 1};
    static int gen = -1;
    if (gen < 0) { const char *e = getenv("RDB_OZAKI_GEN"); gen = e ? atoi(e) : 2; }
    if (gen == 2 && nslices <= 8) {
        using namespace oz2;
        static int kbytes = -1;
        if (kbytes < 0) { const char *e = getenv("RDB_OZAKI_KB"); kbytes = (e && atoi(e) == 32) ? 32 : 64; }
        if (!make_map(&mA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, As, Kp, M, nslices, BM2, kbytes, nslices, a_stride) ||
            !make_map(&mB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Bs, Kp, N, nslices, BN2, kbytes, nslices, b_stride))
            return cudaErrorInvalidValue;
        // one-plane boxes of the same operands: what the kernels load when the plane words say that only some planes are live
        CUtensorMap mA1, mB1;
        if (!make_map(&mA1, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, As, Kp, M, nslices, BM2, kbytes, 1, a_stride) ||
            !make_map(&mB1, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Bs, Kp, N, nslices, BN2, kbytes, 1, b_stride))
            return cudaErrorInvalidValue;
        CUtensorMap mA1n, mB1w;       // the same planes with the box heights exchanged, for the role swap of the trimmed path
        if (!make_map(&mA1n, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, As, Kp, M, nslices, BN2, kbytes, 1, a_stride) ||
            !make_map(&mB1w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Bs, Kp, N, nslices, BM2, kbytes, 1, b_stride))
            return cudaErrorInvalidValue;
        CUtensorMap mA1x, mB1x, mA1nx, mB1wx;   // one-plane boxes of 128 bytes of K (128-byte swizzle): wide K-blocks of the trimmed path
        if (!make_map(&mA1x, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, As, Kp, M, nslices, BM2, 128, 1, a_stride) ||
            !make_map(&mB1x, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Bs, Kp, N, nslices, BN2, 128, 1, b_stride) ||
            !make_map(&mA1nx, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, As, Kp, M, nslices, BN2, 128, 1, a_stride) ||
            !make_map(&mB1wx, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Bs, Kp, N, nslices, BM2, 128, 1, b_stride))
            return cudaErrorInvalidValue;
        dim3 grid2((unsigned)((M + BM2 - 1) / BM2), (unsigned)((N + BN2 - 1) / BN2));
        static int pair = -1;
        if (pair < 0) { const char *e = getenv("RDB_OZAKI_PAIR"); pair = e ? atoi(e) : 0; }   // measured: 1.26 ms per 4096^3 against 1.21 ms for the multicast cluster kernel
        if (pair && nslices == 7 && ozaki_bits() == 8 && kbytes == 64 && !getenv("RDB_OZAKI_CLUSTER")) {
            // CTA pairs along m (cta_group::2): see umma_ozaki2p_kernel
            CUtensorMap mAp, mB2, mB1, mBh;
            if (!make_map(&mAp, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, As, Kp, M, nslices, BM2, 64, nslices, a_stride) ||
                !make_map(&mB2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Bs, Kp, N, nslices, BN2, 64, 2, b_stride) ||
                !make_map(&mB1, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Bs, Kp, N, nslices, BN2, 64, 1, b_stride) ||
                !make_map(&mBh, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Bs, Kp, N, nslices, BN2 / 2, 64, 1, b_stride))
                return cudaErrorInvalidValue;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2u * (unsigned)((M + 2 * BM2 - 1) / (2 * BM2)), (unsigned)((N + BN2 - 1) / BN2));
            cfg.blockDim = dim3(OZ_THREADS); cfg.dynamicSmemBytes = smem_bytes_pair(); cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            static bool attrp = false;
            if (!attrp) {
                cudaError_t e = cudaFuncSetAttribute(umma_ozaki2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_pair());
                if (e != cudaSuccess) return e;
                attrp = true;
            }
            int Mi = (int)M, Ni = (int)N, Ki = (int)Kp;
            return cudaLaunchKernelEx(&cfg, umma_ozaki2p_kernel, mAp, mB2, mB1, mBh, Mi, Ni, Ki, P);
        }
        static int cl = -1;
        if (cl < 0) { const char *e = getenv("RDB_OZAKI_CLUSTER"); cl = e ? atoi(e) : 2; if (cl != 1 && cl != 2 && cl != 4) cl = 1; }
        // 2-CTA clusters take any number of column tiles: an odd count is padded with one idle tile column (its loads lie outside the
        // tensor and arrive as zeros, its stores are masked), so panels and ragged shapes run the cluster kernel - and its trimmed
        // path - too
        if (cl == 2 && (nslices == 8 || nslices == 7) && kbytes == 64 && (grid2.y & 1) && !getenv("RDB_OZAKI_NOPAD")) grid2.y += 1;
        if ((nslices == 8 || nslices == 7) && kbytes == 64 && cl > 1 && grid2.y % cl == 0) {
            // cluster of `cl` CTAs along n sharing the A stage through TMA multicast
            CUtensorMap mAmc;
            if (!make_map(&mAmc, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, As, Kp, M, nslices, BM2, kbytes, (nslices + cl - 1) / cl, a_stride)) return cudaErrorInvalidValue;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid2; cfg.blockDim = dim3(OZ_THREADS); cfg.dynamicSmemBytes = smem_bytes_cluster(nslices, cl); cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = (unsigned)cl; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            static bool attrc = false;
            if (!attrc) {
                cudaError_t e = cudaFuncSetAttribute(umma_ozaki2c_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_cluster(8, 2));
                if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_ozaki2c_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_cluster(8, 4));
                if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_ozaki2c_kernel<7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_cluster(7, 2));
                if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_ozaki2c_kernel<7, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_cluster(7, 4));
                if (e != cudaSuccess) return e;
                attrc = true;
            }
            int Mi = (int)M, Ni = (int)N, Ki = (int)Kp;
            static int persistent = -1;
            if (persistent < 0) { const char *e = getenv("RDB_OZAKI_PERSISTENT"); persistent = e ? atoi(e) : 1; }
            // measured (4096 x 4096 outputs): +4 % at K = 1024, nothing at K = 4096, -0.5 % at K = 8192 (the static deal over the resident
            // clusters loses what the hardware's dynamic CTA scheduling evens out) => short contractions only; 2 = always (tests)
            if ((persistent == 2 || (persistent == 1 && K <= 2048)) && nslices == 7 && cl == 2) {
                static int max_clusters = -1;
                static bool attrcp = false;
                if (!attrcp) {
                    cudaError_t e = cudaFuncSetAttribute(umma_ozaki2cp_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_cluster(7, 2));
                    if (e != cudaSuccess) return e;
                    attrcp = true;
                }
                if (max_clusters < 0) {
                    cudaLaunchConfig_t q = {};
                    q.gridDim = dim3(2 * 148); q.blockDim = dim3(OZ_THREADS); q.dynamicSmemBytes = smem_bytes_cluster(7, 2);
                    cudaLaunchAttribute qa[1];
                    qa[0].id = cudaLaunchAttributeClusterDimension;
                    qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
                    q.attrs = qa; q.numAttrs = 1;
                    int nc = 0;
                    if (cudaOccupancyMaxActiveClusters(&nc, umma_ozaki2cp_kernel<7>, &q) != cudaSuccess || nc < 1) { cudaGetLastError(); nc = 0; }
                    max_clusters = nc;
                }
                const int m_tiles = (int)grid2.x, n_groups = (int)(grid2.y / 2);
                const int items = m_tiles * n_groups;
                if (max_clusters > 0 && items >= max_clusters) {       // (fewer tiles than clusters: nothing to keep resident for)
                    cudaLaunchConfig_t pc = {};
                    pc.gridDim = dim3(2u * (unsigned)max_clusters); pc.blockDim = dim3(OZ_THREADS); pc.dynamicSmemBytes = smem_bytes_cluster(7, 2); pc.stream = st;
                    cudaLaunchAttribute pa[1];
                    pa[0].id = cudaLaunchAttributeClusterDimension;
                    pa[0].val.clusterDim.x = 2; pa[0].val.clusterDim.y = 1; pa[0].val.clusterDim.z = 1;
                    pc.attrs = pa; pc.numAttrs = 1;
                    return cudaLaunchKernelEx(&pc, umma_ozaki2cp_kernel<7>, mAmc, mB, mA1, mB1, Mi, Ni, Ki, P, m_tiles, n_groups);
                }
            }
            // one-plane products go to the trimmed CTA-pair kernel, launched in front (it returns at once for every other product)
            // (measured, profiles/trim_dbg.sh at 4096^3, (1, 7) planes: 0.85 ms per call against 0.72 for the trimmed cluster path - its
            //  MMA skeleton is 0.08 ms faster but the cta_group::2 loads that complete on the leader's barrier cost 0.27 ms instead
            //  of 0.06 => OFF unless RDB_OZAKI_TPAIR=1; kept for the tests and the next attempt)
            static const bool tpair = [] { const char *e = getenv("RDB_OZAKI_TPAIR"); return e && e[0] == '1'; }();
            if (tpair && trim && nslices == 7 && ozaki_bits() == 8 && cl == 2 && M % 256 == 0 && N % 256 == 0) {
                CUtensorMap mA32, mB32;
                if (!make_map(&mA32, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, As, Kp, M, nslices, BN2 / 2, kbytes, 1, a_stride) ||
                    !make_map(&mB32, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Bs, Kp, N, nslices, BN2 / 2, kbytes, 1, b_stride))
                    return cudaErrorInvalidValue;
                static bool attrtp = false;
                if (!attrtp) {
                    cudaError_t e = cudaFuncSetAttribute(umma_ozaki2tp_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_cluster(7, 2));
                    if (e != cudaSuccess) return e;
                    attrtp = true;
                }
                cudaLaunchConfig_t tc = {};
                tc.gridDim = dim3((unsigned)(2 * (M / 256) * (N / 64))); tc.blockDim = dim3(OZ_THREADS);
                tc.dynamicSmemBytes = smem_bytes_cluster(7, 2); tc.stream = st;
                cudaLaunchAttribute ta_[1];
                ta_[0].id = cudaLaunchAttributeClusterDimension;
                ta_[0].val.clusterDim.x = 2; ta_[0].val.clusterDim.y = 1; ta_[0].val.clusterDim.z = 1;
                tc.attrs = ta_; tc.numAttrs = 1;
                cudaError_t e = cudaLaunchKernelEx(&tc, umma_ozaki2tp_kernel<7>, mA1, mA1n, mA32, mB1w, mB1, mB32, Mi, Ni, Ki, P);
                if (e != cudaSuccess) return e;
                P.pair_first = 1;
                g_chunk_launches += 1;
            }
            if (nslices == 8) {
                if (cl == 2) return cudaLaunchKernelEx(&cfg, umma_ozaki2c_kernel<8, 2>, mAmc, mB, mA1, mB1, mA1n, mB1w, mA1x, mB1x, mA1nx, mB1wx, Mi, Ni, Ki, P);
                return cudaLaunchKernelEx(&cfg, umma_ozaki2c_kernel<8, 4>, mAmc, mB, mA1, mB1, mA1n, mB1w, mA1x, mB1x, mA1nx, mB1wx, Mi, Ni, Ki, P);
            }
            if (cl == 2) return cudaLaunchKernelEx(&cfg, umma_ozaki2c_kernel<7, 2>, mAmc, mB, mA1, mB1, mA1n, mB1w, mA1x, mB1x, mA1nx, mB1wx, Mi, Ni, Ki, P);
            return cudaLaunchKernelEx(&cfg, umma_ozaki2c_kernel<7, 4>, mAmc, mB, mA1, mB1, mA1n, mB1w, mA1x, mB1x, mA1nx, mB1wx, Mi, Ni, Ki, P);
        }
#define RDB_OZ2(SS)                                                                                                        \
    case SS: {                                                                                                             \
        static bool attr2 = false;                                                                                         \
        if (!attr2) {                                                                                                      \
            cudaError_t e = cudaFuncSetAttribute(umma_ozaki2_kernel<SS, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(SS, 64)); \
            if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_ozaki2_kernel<SS, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(SS, 32)); \
            if (e != cudaSuccess) return e;                                                                                \
            attr2 = true;                                                                                                  \
        }                                                                                                                  \
        if (kbytes == 64) umma_ozaki2_kernel<SS, 64><<<grid2, OZ_THREADS, smem_bytes(SS, 64), st>>>(mA, mB, mA1, mB1, (int)M, (int)N, (int)Kp, P); \
        else umma_ozaki2_kernel<SS, 32><<<grid2, OZ_THREADS, smem_bytes(SS, 32), st>>>(mA, mB, mA1, mB1, (int)M, (int)N, (int)Kp, P);   \
    } break;
        switch (nslices) { RDB_OZ2(2) RDB_OZ2(3) RDB_OZ2(4) RDB_OZ2(5) RDB_OZ2(6) RDB_OZ2(7) RDB_OZ2(8) default: return cudaErrorInvalidValue; }
#undef RDB_OZ2
        return cudaGetLastError();
    }
    if (!make_map(&mA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, As, Kp, M, nslices, BM, BKB, 1, a_stride) ||
        !make_map(&mB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, Bs, Kp, N, nslices, BN, BKB, 1, b_stride))
        return cudaErrorInvalidValue;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(umma_ozaki_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
    umma_ozaki_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(mA, mB, (int)M, (int)N, (int)Kp, P);
    return cudaGetLastError();
}

}  // namespace rdb
