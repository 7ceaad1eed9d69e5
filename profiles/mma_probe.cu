// Microbenchmark: cycles per tcgen05.mma kind::i8 as a function of N, accumulator reuse and cta_group, operands in shared memory
// (SWIZZLE_64B K-major tiles like the int8-slice GEMM's).  One issuing thread per CTA, grid = one CTA (or CTA pair) per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/mma_probe.bin profiles/mma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ uint64_t make_desc64(uint32_t saddr) {          // K-major, 64-byte rows, SWIZZLE_64B
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((8 * 64) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc_i8(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (CG == 1)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t *bar) {
    if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

struct Op { int n, acc, a_off, b_off; };
constexpr int SLA = 8192, SLB = 4096, B0 = 7 * SLA;      // slice sizes: 128 x 64 B, 64 x 64 B; B slices behind 7 A slices
constexpr int MAXOPS = 32;
struct Pattern { int count; Op ops[MAXOPS]; const char *name; };

constexpr Pattern rep8(const char *name, int n, bool rot, bool same_ab) {
    Pattern P{}; P.count = 8; P.name = name;
    for (int i = 0; i < 8; ++i) P.ops[i] = Op{n, rot ? (i % (512 / n)) * n : 0, same_ab ? 0 : (i % 7) * SLA, B0 + (same_ab ? 0 : (i % 2) * 2048)};
    return P;
}
constexpr Pattern cluster_step() {       // the cluster kernel's k-step: for A slice p, partners q0.. in groups of <= 4 (accumulator p+q0)
    Pattern P{}; P.name = "cluster-kernel k-step (10 MMAs, sum N = 1792)"; int c = 0;
    for (int p = 0; p < 7; ++p)
        for (int q0 = 0; q0 < 7 - p; q0 += 4) { int nq = 7 - p - q0 < 4 ? 7 - p - q0 : 4; P.ops[c++] = Op{nq * 64, (p + q0) * 64, p * SLA, B0 + q0 * SLB}; }
    P.count = c; return P;
}
constexpr Pattern pair_step() {          // the CTA-pair kernel's k-step (12 MMAs)
    Pattern P{}; P.name = "pair-kernel k-step (12 MMAs, sum N = 1792)";
    const int G4 = B0, G45 = B0 + 2 * SLB, G01 = B0 + 3 * SLB, H = B0 + 4 * SLB;
    const Op o[12] = {{256, 0, 0, G4}, {128, 256, 0, G45}, {64, 384, 0, H}, {256, 64, SLA, G4}, {128, 320, SLA, G45}, {256, 128, 2 * SLA, G4},
                      {64, 384, 2 * SLA, H + 2048}, {256, 192, 3 * SLA, G4}, {128, 256, 4 * SLA, G01}, {64, 384, 4 * SLA, H + 4096},
                      {128, 320, 5 * SLA, G01}, {64, 384, 6 * SLA, H + 6144}};
    for (int i = 0; i < 12; ++i) P.ops[i] = o[i];
    P.count = 12; return P;
}
constexpr Pattern all28() {
    Pattern P{}; P.name = "28 x N=64 (one per slice pair)"; int c = 0;
    for (int p = 0; p < 7; ++p) for (int q = 0; q < 7 - p; ++q) P.ops[c++] = Op{64, (p + q) * 64, p * SLA, B0 + q * SLB};
    P.count = c; return P;
}
constexpr Pattern seven256() {
    Pattern P{}; P.name = "7 x N=256 (hypothetical)";
    for (int p = 0; p < 7; ++p) P.ops[p] = Op{256, (p % 2) * 256, p * SLA, B0};
    P.count = 7; return P;
}
constexpr Pattern mix(const char *name, int n1, int n2) {
    Pattern P{}; P.name = name; P.count = 8;
    for (int i = 0; i < 8; ++i) P.ops[i] = Op{i % 2 ? n2 : n1, (i % 2) * 256, (i % 7) * SLA, B0};
    return P;
}
constexpr int NPAT = 14;
__device__ constexpr Pattern PATS[NPAT] = {
    rep8("8 x N=256 same acc same operands", 256, false, true), rep8("8 x N=256 rotating acc, rotating A", 256, true, false),
    rep8("8 x N=192 rotating", 192, true, false), rep8("8 x N=128 rotating", 128, true, false), rep8("8 x N=64 rotating", 64, true, false),
    rep8("8 x N=64 same acc", 64, false, false), rep8("8 x N=32 rotating", 32, true, false), rep8("8 x N=16 rotating", 16, true, false),
    cluster_step(), pair_step(), all28(), seven256(), mix("4 x (N=256, N=64) alternating", 256, 64), mix("4 x (N=256, N=128) alternating", 256, 128)};
static const Pattern HPATS[NPAT] = {
    rep8("8 x N=256 same acc same operands", 256, false, true), rep8("8 x N=256 rotating acc, rotating A", 256, true, false),
    rep8("8 x N=192 rotating", 192, true, false), rep8("8 x N=128 rotating", 128, true, false), rep8("8 x N=64 rotating", 64, true, false),
    rep8("8 x N=64 same acc", 64, false, false), rep8("8 x N=32 rotating", 32, true, false), rep8("8 x N=16 rotating", 16, true, false),
    cluster_step(), pair_step(), all28(), seven256(), mix("4 x (N=256, N=64) alternating", 256, 64), mix("4 x (N=256, N=128) alternating", 256, 128)};

constexpr int SMEM = 200 * 1024;

template <int CG, int PID>
__device__ __forceinline__ void issue(uint32_t tb, uint64_t d0, int reps) {
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < PATS[PID].count; ++i) {
            constexpr Pattern P = PATS[PID];
            mma_i8<CG>(tb + P.ops[i].acc, d0 + (uint64_t)(P.ops[i].a_off >> 4), d0 + (uint64_t)(P.ops[i].b_off >> 4), idesc_i8(CG == 2 ? 256 : 128, P.ops[i].n), 1u);
        }
    }
}

template <int CG, int PID>
__device__ __forceinline__ void run_one(uint32_t tb, uint64_t d0, int reps, uint32_t crank, uint64_t *bar, uint32_t &phase, long long *out) {
    long long t0 = 0, t1 = 0;
    if (CG == 1 || crank == 0) {
        for (int pass = 0; pass < 2; ++pass) {            // pass 0 warms up
            t0 = clock64();
            issue<CG, PID>(tb, d0, reps);
            commit<CG>(bar);
            mbar_wait(bar, phase); phase ^= 1;
            t1 = clock64();
        }
    } else {
        for (int pass = 0; pass < 2; ++pass) { mbar_wait(bar, phase); phase ^= 1; }
    }
    if (blockIdx.x == 0) out[PID] = t1 - t0;
}

template <int CG>
__global__ void __launch_bounds__(128, 1) probe(int reps, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < (SMEM - 2048) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = (uint32_t)(i * 2654435761u) ^ (blockIdx.x * 40503u);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_ptr;
    uint32_t crank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    uint32_t phase = 0;
    if (threadIdx.x == 0) {
        const uint64_t d0 = make_desc64(smem_u32(smem));
        run_one<CG, 0>(tb, d0, reps, crank, &bar, phase, out);  run_one<CG, 1>(tb, d0, reps, crank, &bar, phase, out);
        run_one<CG, 2>(tb, d0, reps, crank, &bar, phase, out);  run_one<CG, 3>(tb, d0, reps, crank, &bar, phase, out);
        run_one<CG, 4>(tb, d0, reps, crank, &bar, phase, out);  run_one<CG, 5>(tb, d0, reps, crank, &bar, phase, out);
        run_one<CG, 6>(tb, d0, reps, crank, &bar, phase, out);  run_one<CG, 7>(tb, d0, reps, crank, &bar, phase, out);
        run_one<CG, 8>(tb, d0, reps, crank, &bar, phase, out);  run_one<CG, 9>(tb, d0, reps, crank, &bar, phase, out);
        run_one<CG, 10>(tb, d0, reps, crank, &bar, phase, out); run_one<CG, 11>(tb, d0, reps, crank, &bar, phase, out);
        run_one<CG, 12>(tb, d0, reps, crank, &bar, phase, out); run_one<CG, 13>(tb, d0, reps, crank, &bar, phase, out);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(512));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(512));
    }
}

int main(int argc, char **argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 200;
    long long *d_o;
    cudaMalloc(&d_o, NPAT * sizeof(long long));
    std::vector<long long> out(NPAT);
    for (int cg = 1; cg <= 2; ++cg) {
        if (cg == 1) {
            cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
            probe<1><<<148, 128, SMEM>>>(reps, d_o);
        } else {
            cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = SMEM;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int rr = reps;
            cudaLaunchKernelEx(&cfg, probe<2>, rr, d_o);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cta_group::%d failed: %s\n", cg, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), d_o, NPAT * sizeof(long long), cudaMemcpyDeviceToHost);
        printf("== cta_group::%d (M = %d), %d repetitions per pattern, all SMs busy\n", cg, cg == 2 ? 256 : 128, reps);
        for (int i = 0; i < NPAT; ++i) {
            int ideal = 0;
            for (int k = 0; k < HPATS[i].count; ++k) ideal += HPATS[i].ops[k].n / 2;
            const double per = (double)out[i] / reps;
            printf("%-52s %8.1f clk per pattern, %6.1f per MMA, ideal %5d => %.1f %%\n", HPATS[i].name, per, per / HPATS[i].count, ideal, 100.0 * ideal / per);
        }
    }
    return 0;
}
