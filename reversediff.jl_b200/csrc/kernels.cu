// Elementwise / reduction / gather-scatter kernels of the replay engine: kernel families (a) and (b) of
// BASELINE.json's north star.  All of them are HBM-bound streams: coalesced (128-bit where alignment allows),
// grid sized in multiples of the SM count, two-pass deterministic reductions (per-block partial sums with
// warp shuffles, then one finishing block) instead of atomics wherever the access pattern is known.
//
// Compiled with -fmad=false: ForwardDiff's Dual arithmetic on the CPU is not FMA-contracted, and parity with the
// reference's partials (e.g. tanh' = 1 - t*t) is checked to 1e-12 of max|.|.
// The scalar-expression interpreter in this file is the fallback / second-order path; first-order elementwise forward sweeps
// normally run the NVRTC-specialised kernels of jit.cu.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include "kernels.h"
#include "expr_eval.cuh"

namespace rdb {

constexpr int kSMs = 148;
constexpr int kThreads = 256;

static inline int grid_for(int64_t n, int per_thread = 1) {
    int64_t b = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
    int64_t cap = (int64_t)kSMs * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ------------------------------------------------------------------------------------ block reduction
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double block_sum(double v) {   // result valid in thread 0
    __shared__ double ws[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) ws[w] = v;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? ws[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

template <typename T> struct Vec;
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };
template <> struct Vec<float> { using type = float4; static constexpr int N = 4; };
template <typename T> __device__ __forceinline__ void vec_unpack(const typename Vec<T>::type &v, T *o);
template <> __device__ __forceinline__ void vec_unpack<double>(const double2 &v, double *o) { o[0] = v.x; o[1] = v.y; }
template <> __device__ __forceinline__ void vec_unpack<float>(const float4 &v, float *o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
template <typename T> __device__ __forceinline__ typename Vec<T>::type vec_pack(const T *o);
template <> __device__ __forceinline__ double2 vec_pack<double>(const double *o) { return make_double2(o[0], o[1]); }
template <> __device__ __forceinline__ float4 vec_pack<float>(const float *o) { return make_float4(o[0], o[1], o[2], o[3]); }

static inline bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

// ------------------------------------------------------------------------------------------------ fill
template <typename T>
__global__ void k_fill(T *p, int64_t n, T v) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) p[i] = v;
}
template <typename T>
cudaError_t launch_fill(cudaStream_t s, T *p, int64_t n, T v) {
    if (n <= 0) return cudaSuccess;
    k_fill<T><<<grid_for(n, 4), kThreads, 0, s>>>(p, n, v);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------- add (+ fused block sums)
template <typename T, bool VEC, bool REDUCE>
__global__ void __launch_bounds__(kThreads)
k_add(T *__restrict__ out, const T *__restrict__ a, const T *__restrict__ b, T sign, int64_t n, double *block_sums) {
    constexpr int W = Vec<T>::N;
    using V = typename Vec<T>::type;
    double local = 0.0;
    int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        int64_t nv = n / W;
        const V *av = reinterpret_cast<const V *>(a), *bv = reinterpret_cast<const V *>(b);
        V *ov = reinterpret_cast<V *>(out);
        for (int64_t i = tid; i < nv; i += st) {
            T x[W], y[W], z[W];
            vec_unpack<T>(av[i], x);
            vec_unpack<T>(bv[i], y);
#pragma unroll
            for (int j = 0; j < W; ++j) { z[j] = x[j] + sign * y[j]; if (REDUCE) local += (double)z[j]; }
            ov[i] = vec_pack<T>(z);
        }
        for (int64_t i = nv * W + tid; i < n; i += st) {
            T z = a[i] + sign * b[i];
            out[i] = z;
            if (REDUCE) local += (double)z;
        }
    } else {
        for (int64_t i = tid; i < n; i += st) {
            T z = a[i] + sign * b[i];
            out[i] = z;
            if (REDUCE) local += (double)z;
        }
    }
    if (REDUCE) {
        double s = block_sum(local);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
    }
}
template <typename T>
cudaError_t launch_add(cudaStream_t s, T *out, const T *a, const T *b, T sign, int64_t n, T *block_sums_) {
    double *bs = reinterpret_cast<double *>(block_sums_);
    bool vec = aligned16(out) && aligned16(a) && aligned16(b);
    int grid = bs ? kReduceBlocks : grid_for(n, Vec<T>::N * 2);
    if (vec) {
        if (bs) k_add<T, true, true><<<grid, kThreads, 0, s>>>(out, a, b, sign, n, bs);
        else k_add<T, true, false><<<grid, kThreads, 0, s>>>(out, a, b, sign, n, bs);
    } else {
        if (bs) k_add<T, false, true><<<grid, kThreads, 0, s>>>(out, a, b, sign, n, bs);
        else k_add<T, false, false><<<grid, kThreads, 0, s>>>(out, a, b, sign, n, bs);
    }
    return cudaGetLastError();
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(kThreads) k_block_sums(const T *__restrict__ x, int64_t n, double *block_sums) {
    constexpr int W = Vec<T>::N;
    using V = typename Vec<T>::type;
    double local = 0.0;
    int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        int64_t nv = n / W;
        const V *xv = reinterpret_cast<const V *>(x);
        for (int64_t i = tid; i < nv; i += st) {
            T e[W];
            vec_unpack<T>(xv[i], e);
#pragma unroll
            for (int j = 0; j < W; ++j) local += (double)e[j];
        }
        for (int64_t i = nv * W + tid; i < n; i += st) local += (double)x[i];
    } else {
        for (int64_t i = tid; i < n; i += st) local += (double)x[i];
    }
    double s = block_sum(local);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
}
template <typename T>
cudaError_t launch_block_sums(cudaStream_t s, const T *x, int64_t n, T *block_sums_) {
    double *bs = reinterpret_cast<double *>(block_sums_);
    if (aligned16(x)) k_block_sums<T, true><<<kReduceBlocks, kThreads, 0, s>>>(x, n, bs);
    else k_block_sums<T, false><<<kReduceBlocks, kThreads, 0, s>>>(x, n, bs);
    return cudaGetLastError();
}

template <typename T>
__global__ void __launch_bounds__(1024) k_finish_sum(const double *block_sums, T *out, double scale, int accumulate) {
    double v = 0.0;
    for (int i = threadIdx.x; i < kReduceBlocks; i += blockDim.x) v += block_sums[i];
    v = block_sum(v);
    if (threadIdx.x == 0) out[0] = accumulate ? (T)((double)out[0] + scale * v) : (T)(scale * v);
}
template <typename T>
cudaError_t launch_finish_sum(cudaStream_t s, const T *block_sums_, T *out, T scale, bool accumulate) {
    k_finish_sum<T><<<1, 1024, 0, s>>>(reinterpret_cast<const double *>(block_sums_), out, (double)scale, accumulate ? 1 : 0);
    return cudaGetLastError();
}

// block_sums <- per-block sums of a[i]*b[i]   (dot forward, reductions.jl:106-112; scalar-argument adjoints, propagation.jl:98-108)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_dot_block_sums(const T *__restrict__ a, const T *__restrict__ b, int64_t n, double *block_sums) {
    double local = 0.0;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) local += (double)(a[i] * b[i]);
    double s = block_sum(local);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
}
template <typename T>
cudaError_t launch_dot_block_sums(cudaStream_t s, const T *a, const T *b, int64_t n, T *block_sums_) {
    k_dot_block_sums<T><<<kReduceBlocks, kThreads, 0, s>>>(a, b, n, reinterpret_cast<double *>(block_sums_));
    return cudaGetLastError();
}

// dst[i + r*n] += delta[r]*src[i]     (dot reverse: lmul!(output_deriv, tmp); increment_deriv!, reductions.jl:114-131)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_scal_acc(T *__restrict__ dst, const T *__restrict__ src, const T *__restrict__ delta, int64_t n, int64_t R, int overwrite) {
    int64_t tot = n * R;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < tot; i += st) { int64_t r = i / n; T v = delta[r] * src[i - r * n]; dst[i] = overwrite ? v : dst[i] + v; }
}
template <typename T>
cudaError_t launch_scal_acc(cudaStream_t s, T *dst, const T *src, const T *delta, int64_t n, int64_t R, bool overwrite) {
    if (n * R <= 0) return cudaSuccess;
    k_scal_acc<T><<<grid_for(n * R, 4), kThreads, 0, s>>>(dst, src, delta, n, R, overwrite ? 1 : 0);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ axpy
template <typename T, bool VEC, bool OVERWRITE>
__global__ void __launch_bounds__(kThreads) k_axpy(T *__restrict__ dst, const T *__restrict__ src, T sign, int64_t n) {
    constexpr int W = Vec<T>::N;
    using V = typename Vec<T>::type;
    int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        int64_t nv = n / W;
        const V *sv = reinterpret_cast<const V *>(src);
        V *dv = reinterpret_cast<V *>(dst);
        for (int64_t i = tid; i < nv; i += st) {
            T x[W], y[W];
            vec_unpack<T>(sv[i], x);
            if (OVERWRITE) {
#pragma unroll
                for (int j = 0; j < W; ++j) y[j] = sign * x[j];
            } else {
                vec_unpack<T>(dv[i], y);
#pragma unroll
                for (int j = 0; j < W; ++j) y[j] = y[j] + sign * x[j];
            }
            dv[i] = vec_pack<T>(y);
        }
        for (int64_t i = nv * W + tid; i < n; i += st) dst[i] = OVERWRITE ? sign * src[i] : dst[i] + sign * src[i];
    } else {
        for (int64_t i = tid; i < n; i += st) dst[i] = OVERWRITE ? sign * src[i] : dst[i] + sign * src[i];
    }
}
template <typename T>
cudaError_t launch_axpy(cudaStream_t s, T *dst, const T *src, T sign, int64_t n, bool overwrite) {
    if (n <= 0) return cudaSuccess;
    bool vec = aligned16(dst) && aligned16(src);
    int grid = grid_for(n, Vec<T>::N * 2);
    if (vec) {
        if (overwrite) k_axpy<T, true, true><<<grid, kThreads, 0, s>>>(dst, src, sign, n);
        else k_axpy<T, true, false><<<grid, kThreads, 0, s>>>(dst, src, sign, n);
    } else {
        if (overwrite) k_axpy<T, false, true><<<grid, kThreads, 0, s>>>(dst, src, sign, n);
        else k_axpy<T, false, false><<<grid, kThreads, 0, s>>>(dst, src, sign, n);
    }
    return cudaGetLastError();
}

// blockIdx.y = seed r; 128-bit stores when a seed's row is 16-byte aligned (no per-element division, one broadcast load)
template <typename T, bool OVERWRITE, bool VEC>
__global__ void __launch_bounds__(kThreads)
k_acc_scalar(T *__restrict__ dst, int64_t n, const T *__restrict__ delta, T scale) {
    const int64_t r = blockIdx.y;
    const T d = scale * delta[r];
    T *p = dst + r * n;
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (VEC) {
        constexpr int NV = Vec<T>::N;
        using V = typename Vec<T>::type;
        V *pv = reinterpret_cast<V *>(p);
        T dv[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) dv[k] = d;
        const V fillv = vec_pack<T>(dv);
        for (const int64_t nv = n / NV; i < nv; i += st) {
            if (OVERWRITE) pv[i] = fillv;
            else {
                T o[NV];
                vec_unpack<T>(pv[i], o);
#pragma unroll
                for (int k = 0; k < NV; ++k) o[k] = o[k] + d;
                pv[i] = vec_pack<T>(o);
            }
        }
    } else {
        for (; i < n; i += st) p[i] = OVERWRITE ? d : p[i] + d;
    }
}
template <typename T>
cudaError_t launch_acc_scalar(cudaStream_t s, T *dst, int64_t n, int64_t R, const T *delta, T scale, bool overwrite) {
    if (n * R <= 0) return cudaSuccess;
    if (R > 65535) return cudaErrorInvalidValue;
    constexpr int NV = Vec<T>::N;
    const bool vec = n % NV == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    const int64_t work = vec ? n / NV : n;
    int64_t gx = (work + kThreads * 4 - 1) / (kThreads * 4);
    const int64_t capx = std::max<int64_t>(1, ((int64_t)kSMs * 16 + R - 1) / R);
    if (gx > capx) gx = capx;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)R);
    if (overwrite) { if (vec) k_acc_scalar<T, true, true><<<grid, kThreads, 0, s>>>(dst, n, delta, scale); else k_acc_scalar<T, true, false><<<grid, kThreads, 0, s>>>(dst, n, delta, scale); }
    else { if (vec) k_acc_scalar<T, false, true><<<grid, kThreads, 0, s>>>(dst, n, delta, scale); else k_acc_scalar<T, false, false><<<grid, kThreads, 0, s>>>(dst, n, delta, scale); }
    return cudaGetLastError();
}

// ======================================================================================================
// Scalar-expression interpreter: one Dual{N} (or hyper-dual) evaluation per element
// ======================================================================================================
template <int N>
struct ExprDev {
    const void *arg[N];
    int64_t stride[N][kMaxDims];
    int same[N];
    void *partial[N];
    void *second[Tri<N>::NH];
    int64_t dims[kMaxDims];
    int64_t n;
    void *out;
    const int4 *ops;
    const double *fconst;
    double *block_sums;
    int n_ops;
};

template <typename T, int N, int ORDER, bool REDUCE>
__global__ void __launch_bounds__(kThreads) k_expr_forward(const ExprDev<N> P) {
    __shared__ int4 s_ops[kMaxExprRegs];
    for (int i = threadIdx.x; i < P.n_ops; i += blockDim.x) s_ops[i] = P.ops[i];
    __syncthreads();
    double local = 0.0;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < P.n; i += st) {
        T x[N];
        bool need_idx = false;
#pragma unroll
        for (int p = 0; p < N; ++p) need_idx |= !P.same[p];
        int64_t idx[kMaxDims] = {0, 0, 0, 0};
        if (need_idx) {
            int64_t rem = i;
#pragma unroll
            for (int d = 0; d < kMaxDims; ++d) {
                if (P.n <= 0x7fffffffLL) { unsigned r32 = (unsigned)rem, dd = (unsigned)P.dims[d]; idx[d] = r32 % dd; rem = r32 / dd; }
                else { idx[d] = rem % P.dims[d]; rem = rem / P.dims[d]; }
            }
        }
#pragma unroll
        for (int p = 0; p < N; ++p) {
            int64_t off = i;
            if (!P.same[p]) {
                off = 0;
#pragma unroll
                for (int d = 0; d < kMaxDims; ++d) off += idx[d] * P.stride[p][d];
            }
            x[p] = reinterpret_cast<const T *>(P.arg[p])[off];
        }
        T val, dd[N], hh[Tri<N>::NH];
        eval_expr<T, N, ORDER>(s_ops, P.n_ops, P.fconst, x, val, dd, hh);
        reinterpret_cast<T *>(P.out)[i] = val;
#pragma unroll
        for (int p = 0; p < N; ++p)
            if (P.partial[p]) reinterpret_cast<T *>(P.partial[p])[i] = dd[p];
        if (ORDER == 2) {
#pragma unroll
            for (int h = 0; h < Tri<N>::NH; ++h)
                if (P.second[h]) reinterpret_cast<T *>(P.second[h])[i] = hh[h];
        }
        if (REDUCE) local += (double)val;
    }
    if (REDUCE) {
        double s = block_sum(local);
        if (threadIdx.x == 0) P.block_sums[blockIdx.x] = s;
    }
}

template <typename T, int N>
static cudaError_t expr_forward_n(cudaStream_t s, const ExprParams &p) {
    ExprDev<N> d;
    for (int i = 0; i < N; ++i) {
        d.arg[i] = p.arg[i].ptr;
        d.same[i] = p.arg[i].same;
        d.partial[i] = p.partial[i];
        for (int k = 0; k < kMaxDims; ++k) d.stride[i][k] = p.arg[i].stride[k];
    }
    for (int h = 0; h < Tri<N>::NH; ++h) d.second[h] = p.order == 2 ? p.second[h] : nullptr;
    for (int k = 0; k < kMaxDims; ++k) d.dims[k] = p.dims[k];
    d.n = p.n; d.out = p.out; d.ops = reinterpret_cast<const int4 *>(p.ops); d.fconst = p.fconst;
    d.block_sums = reinterpret_cast<double *>(p.block_sums); d.n_ops = p.n_ops;
    int grid = p.reduce ? kReduceBlocks : grid_for(p.n, 2);
    if (p.order == 2) {
        if constexpr (N <= 4) {                 // second partials (hessian!) are built for closures of up to four arguments
            if (p.reduce) k_expr_forward<T, N, 2, true><<<grid, kThreads, 0, s>>>(d);
            else k_expr_forward<T, N, 2, false><<<grid, kThreads, 0, s>>>(d);
        } else {
            return cudaErrorInvalidValue;
        }
    } else {
        if (p.reduce) k_expr_forward<T, N, 1, true><<<grid, kThreads, 0, s>>>(d);
        else k_expr_forward<T, N, 1, false><<<grid, kThreads, 0, s>>>(d);
    }
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_expr_forward(cudaStream_t s, const ExprParams &p) {
    if (p.n <= 0) return cudaSuccess;
    switch (p.n_args) {
        case 1: return expr_forward_n<T, 1>(s, p);
        case 2: return expr_forward_n<T, 2>(s, p);
        case 3: return expr_forward_n<T, 3>(s, p);
        case 4: return expr_forward_n<T, 4>(s, p);
        case 5: return expr_forward_n<T, 5>(s, p);
        case 6: return expr_forward_n<T, 6>(s, p);
        case 7: return expr_forward_n<T, 7>(s, p);
        case 8: return expr_forward_n<T, 8>(s, p);
    }
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------- reverse: delta * partial accumulate
template <typename T, bool OVERWRITE>
__global__ void __launch_bounds__(kThreads)
k_mul_acc(T *__restrict__ dst, const T *__restrict__ delta, const T *__restrict__ partial, int64_t n, int64_t R) {
    int64_t tot = n * R;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    if (R == 1) {
        for (; i < tot; i += st) {
            T v = delta[i] * partial[i];
            dst[i] = OVERWRITE ? v : dst[i] + v;
        }
    } else {
        for (; i < tot; i += st) {
            T v = delta[i] * partial[i % n];
            dst[i] = OVERWRITE ? v : dst[i] + v;
        }
    }
}
// the same with delta[e + r n] = scale * dscalar[r] for every e: what `sum`/`mean` reverse would have written into deriv(output)
// first (reductions.jl:15-21) - the fill and its re-read are skipped, the product is formed from the same two factors
template <typename T, bool OVERWRITE>
__global__ void __launch_bounds__(kThreads)
k_mul_acc_sdelta(T *__restrict__ dst, const T *__restrict__ dscalar, T scale, const T *__restrict__ partial, int64_t n, int64_t R) {
    int64_t tot = n * R;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < tot; i += st) {
        const int64_t r = i / n, e = i - r * n;
        const T d = scale * dscalar[r];
        T v = d * partial[e];
        dst[i] = OVERWRITE ? v : dst[i] + v;
    }
}
template <typename T>
cudaError_t launch_mul_acc_sdelta(cudaStream_t s, T *dst, const T *dscalar, T scale, const T *partial, int64_t n, int64_t R, bool overwrite) {
    if (n * R <= 0) return cudaSuccess;
    int grid = grid_for(n * R, 4);
    if (overwrite) k_mul_acc_sdelta<T, true><<<grid, kThreads, 0, s>>>(dst, dscalar, scale, partial, n, R);
    else k_mul_acc_sdelta<T, false><<<grid, kThreads, 0, s>>>(dst, dscalar, scale, partial, n, R);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_mul_acc(cudaStream_t s, T *dst, const T *delta, const T *partial, int64_t n, int64_t R, bool overwrite) {
    if (n * R <= 0) return cudaSuccess;
    int grid = grid_for(n * R, 4);
    if (overwrite) k_mul_acc<T, true><<<grid, kThreads, 0, s>>>(dst, delta, partial, n, R);
    else k_mul_acc<T, false><<<grid, kThreads, 0, s>>>(dst, delta, partial, n, R);
    return cudaGetLastError();
}

// column-vector clamp: pass 1 -> scratch[(c*R + r)*d0 + i] = sum over the chunk's columns
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_colvec_partial(double *__restrict__ scratch, const T *__restrict__ delta, const T *__restrict__ partial, int64_t d0,
                 int64_t d1, int64_t R, int64_t cols_per_chunk) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t c = blockIdx.y, r = blockIdx.z;
    if (i >= d0) return;
    int64_t j0 = c * cols_per_chunk, j1 = j0 + cols_per_chunk;
    if (j1 > d1) j1 = d1;
    const T *dl = delta + r * d0 * d1;
    double acc = 0.0;
#pragma unroll 4
    if (partial) { for (int64_t j = j0; j < j1; ++j) acc += (double)(dl[i + j * d0] * partial[i + j * d0]); }
    else { for (int64_t j = j0; j < j1; ++j) acc += (double)dl[i + j * d0]; }
    scratch[(c * R + r) * d0 + i] = acc;
}
// pass 1 fused with the same-shape argument's accumulation of the SAME instruction: delta is read once for both
// (dx[e] (+)= delta[e] * px[e] exactly as k_mul_acc does it; the column sums exactly as k_colvec_partial)
template <typename T, bool OVERWRITE, bool SDELTA>
__global__ void __launch_bounds__(kThreads)
k_colvec_partial_fused(double *__restrict__ scratch, const T *__restrict__ delta, const T *__restrict__ partial_col,
                       T *__restrict__ dx, const T *__restrict__ partial_same, int64_t d0, int64_t d1, int64_t R, int64_t cols_per_chunk, T scale) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t c = blockIdx.y, r = blockIdx.z;
    if (i >= d0) return;
    int64_t j0 = c * cols_per_chunk, j1 = j0 + cols_per_chunk;
    if (j1 > d1) j1 = d1;
    const T *dl = SDELTA ? delta : delta + r * d0 * d1;       // SDELTA: delta holds R scalars (the adjoint of a `sum`/`mean` output)
    const T ds = SDELTA ? scale * delta[r] : (T)0;
    T *dxr = dx + r * d0 * d1;
    double acc = 0.0;
#pragma unroll 4
    for (int64_t j = j0; j < j1; ++j) {
        const T d = SDELTA ? ds : dl[i + j * d0];
        const T v = d * partial_same[i + j * d0];
        dxr[i + j * d0] = OVERWRITE ? v : dxr[i + j * d0] + v;
        acc += (double)(d * partial_col[i + j * d0]);
    }
    scratch[(c * R + r) * d0 + i] = acc;
}
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_colvec_finish(T *__restrict__ dst, const double *__restrict__ scratch, int64_t d0, int64_t R, int64_t chunks, int overwrite) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= d0 * R) return;
    double acc = 0.0;
    for (int64_t c = 0; c < chunks; ++c) acc += scratch[c * R * d0 + i];
    dst[i] = overwrite ? (T)acc : dst[i] + (T)acc;
}
template <typename T>
cudaError_t launch_mul_acc_colvec(cudaStream_t s, T *dst, const T *delta, const T *partial, int64_t d0, int64_t d1, int64_t R,
                                  T *scratch_, int64_t scratch_elems, bool overwrite) {
    if (d0 * d1 * R <= 0) return cudaSuccess;
    double *scratch = reinterpret_cast<double *>(scratch_);
    int64_t bx = (d0 + kThreads - 1) / kThreads;
    // enough column chunks to fill the machine (~8 CTAs per SM), bounded by the scratch size
    int64_t want = (kSMs * 8 + bx * R - 1) / (bx * R);
    if (want < 1) want = 1;
    if (want > d1) want = d1;
    int64_t maxc = scratch_elems / (d0 * R);
    if (maxc < 1) return cudaErrorMemoryAllocation;
    if (want > maxc) want = maxc;
    if (want > 65535) want = 65535;
    int64_t cols = (d1 + want - 1) / want;
    int64_t chunks = (d1 + cols - 1) / cols;
    if (R > 65535) return cudaErrorInvalidValue;
    dim3 grid((unsigned)bx, (unsigned)chunks, (unsigned)R);
    k_colvec_partial<T><<<grid, kThreads, 0, s>>>(scratch, delta, partial, d0, d1, R, cols);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_colvec_finish<T><<<(unsigned)((d0 * R + kThreads - 1) / kThreads), kThreads, 0, s>>>(dst, scratch, d0, R, chunks, overwrite ? 1 : 0);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_mul_acc_colvec_fused(cudaStream_t s, T *dst_col, const T *delta, const T *partial_col, int64_t d0, int64_t d1, int64_t R,
                                        T *scratch_, int64_t scratch_elems, bool overwrite_col, T *dst_same, const T *partial_same,
                                        bool overwrite_same, bool sdelta, T scale) {
    if (d0 * d1 * R <= 0) return cudaSuccess;
    double *scratch = reinterpret_cast<double *>(scratch_);
    int64_t bx = (d0 + kThreads - 1) / kThreads;
    int64_t want = (kSMs * 8 + bx * R - 1) / (bx * R);
    if (want < 1) want = 1;
    if (want > d1) want = d1;
    int64_t maxc = scratch_elems / (d0 * R);
    if (maxc < 1) return cudaErrorMemoryAllocation;
    if (want > maxc) want = maxc;
    if (want > 65535) want = 65535;
    int64_t cols = (d1 + want - 1) / want;
    int64_t chunks = (d1 + cols - 1) / cols;
    if (R > 65535) return cudaErrorInvalidValue;
    dim3 grid((unsigned)bx, (unsigned)chunks, (unsigned)R);
    if (sdelta) {
        if (overwrite_same) k_colvec_partial_fused<T, true, true><<<grid, kThreads, 0, s>>>(scratch, delta, partial_col, dst_same, partial_same, d0, d1, R, cols, scale);
        else k_colvec_partial_fused<T, false, true><<<grid, kThreads, 0, s>>>(scratch, delta, partial_col, dst_same, partial_same, d0, d1, R, cols, scale);
    } else {
        if (overwrite_same) k_colvec_partial_fused<T, true, false><<<grid, kThreads, 0, s>>>(scratch, delta, partial_col, dst_same, partial_same, d0, d1, R, cols, scale);
        else k_colvec_partial_fused<T, false, false><<<grid, kThreads, 0, s>>>(scratch, delta, partial_col, dst_same, partial_same, d0, d1, R, cols, scale);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_colvec_finish<T><<<(unsigned)((d0 * R + kThreads - 1) / kThreads), kThreads, 0, s>>>(dst_col, scratch, d0, R, chunks, overwrite_col ? 1 : 0);
    return cudaGetLastError();
}

struct GenericIdx {
    int64_t dims[kMaxDims];
    int64_t dst_stride[kMaxDims];
};
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_mul_acc_generic(T *dst, int64_t dst_n, GenericIdx g, const T *__restrict__ delta, const T *__restrict__ partial,
                  int64_t n, int64_t R) {
    int64_t tot = n * R;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < tot; i += st) {
        int64_t r = i / n, e = i - r * n, rem = e, off = 0;
#pragma unroll
        for (int d = 0; d < kMaxDims; ++d) {
            off += (rem % g.dims[d]) * g.dst_stride[d];
            rem /= g.dims[d];
        }
        atomicAdd(dst + off + r * dst_n, delta[i] * partial[e]);
    }
}
template <typename T>
cudaError_t launch_mul_acc_generic(cudaStream_t s, T *dst, int64_t dst_n, const int64_t *dst_stride, const T *delta,
                                   const T *partial, const int64_t *dims, int64_t n, int64_t R) {
    if (n * R <= 0) return cudaSuccess;
    GenericIdx g;
    for (int d = 0; d < kMaxDims; ++d) { g.dims[d] = dims[d]; g.dst_stride[d] = dst_stride[d]; }
    k_mul_acc_generic<T><<<grid_for(n * R, 2), kThreads, 0, s>>>(dst, dst_n, g, delta, partial, n, R);
    return cudaGetLastError();
}

// dst[i + r*n] += delta[clamp(i) + r*delta_n]   (reduction_increment_deriv!, propagation.jl:217-223: the adjoint of sum over dims)
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_bcast_acc(T *__restrict__ dst, GenericIdx g, const T *__restrict__ delta, int64_t n, int64_t R, int64_t delta_n, int overwrite) {
    int64_t tot = n * R;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < tot; i += st) {
        int64_t r = i / n, e = i - r * n, rem = e, off = 0;
#pragma unroll
        for (int d = 0; d < kMaxDims; ++d) {
            off += (rem % g.dims[d]) * g.dst_stride[d];
            rem /= g.dims[d];
        }
        dst[i] = overwrite ? delta[off + r * delta_n] : dst[i] + delta[off + r * delta_n];
    }
}
template <typename T>
cudaError_t launch_bcast_acc(cudaStream_t s, T *dst, const int64_t *dims, const int64_t *delta_stride, const T *delta, int64_t n, int64_t R, int64_t delta_n, bool overwrite) {
    if (n * R <= 0) return cudaSuccess;
    GenericIdx g;
    for (int d = 0; d < kMaxDims; ++d) { g.dims[d] = dims[d]; g.dst_stride[d] = delta_stride[d]; }
    k_bcast_acc<T><<<grid_for(n * R, 2), kThreads, 0, s>>>(dst, g, delta, n, R, delta_n, overwrite ? 1 : 0);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ gather / scatter
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_gather_cols(T *__restrict__ out, const T *__restrict__ src, int64_t src_n, const int64_t *__restrict__ map, int64_t n, int64_t K) {
    int64_t tot = n * K;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < tot; i += st) {
        int64_t c = i / n, k = i - c * n;
        out[i] = src[map[k] + c * src_n];
    }
}
template <typename T>
cudaError_t launch_gather_cols(cudaStream_t s, T *out, const T *src, int64_t src_n, const int64_t *map, int64_t n, int64_t K) {
    if (n * K <= 0) return cudaSuccess;
    k_gather_cols<T><<<grid_for(n * K, 2), kThreads, 0, s>>>(out, src, src_n, map, n, K);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_gather(cudaStream_t s, T *out, const T *src, const int64_t *map, int64_t n) {
    return launch_gather_cols<T>(s, out, src, 0, map, n, 1);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_scatter_add(T *dst, int64_t dst_n, const int64_t *__restrict__ map, const T *__restrict__ src, int64_t n, int64_t R) {
    int64_t tot = n * R;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < tot; i += st) {
        int64_t r = i / n, k = i - r * n;
        atomicAdd(dst + map[k] + r * dst_n, src[i]);
    }
}
template <typename T>
cudaError_t launch_scatter_add(cudaStream_t s, T *dst, int64_t dst_n, const int64_t *map, const T *src, int64_t n, int64_t R) {
    if (n * R <= 0) return cudaSuccess;
    k_scatter_add<T><<<grid_for(n * R, 2), kThreads, 0, s>>>(dst, dst_n, map, src, n, R);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ seeding / extraction
template <typename T>
__global__ void __launch_bounds__(kThreads) k_seed_onehot(T *der, int64_t m, int64_t R, int64_t i0) {
    int64_t tot = m * R;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < tot; i += st) {
        int64_t r = i / m, e = i - r * m;
        der[i] = (e == i0 + r) ? (T)1 : (T)0;
    }
}
template <typename T>
cudaError_t launch_seed_onehot(cudaStream_t s, T *der, int64_t m, int64_t R, int64_t i0) {
    if (m * R <= 0) return cudaSuccess;
    k_seed_onehot<T><<<grid_for(m * R, 4), kThreads, 0, s>>>(der, m, R, i0);
    return cudaGetLastError();
}
template <typename T>
__global__ void __launch_bounds__(kThreads) k_seed_gather(T *__restrict__ tv, const int64_t *__restrict__ map, int64_t n, int64_t K, int64_t c0) {
    int64_t tot = n * K;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < tot; i += st) {
        int64_t c = i / n, k = i - c * n;
        tv[i] = (map[k] == c0 + c) ? (T)1 : (T)0;
    }
}
template <typename T>
cudaError_t launch_seed_gather(cudaStream_t s, T *tv, const int64_t *map, int64_t n, int64_t K, int64_t c0) {
    if (n * K <= 0) return cudaSuccess;
    k_seed_gather<T><<<grid_for(n * K, 4), kThreads, 0, s>>>(tv, map, n, K, c0);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_seed_tangent(cudaStream_t s, T *tv, int64_t n, int64_t K, int64_t c0) {
    return launch_seed_onehot<T>(s, tv, n, K, c0);
}

// result[r + j*ld] = xder[j + r*n] : 32x32 shared-memory transpose, both sides coalesced
template <typename T>
__global__ void __launch_bounds__(256) k_rows_out(T *__restrict__ result, int64_t ld, const T *__restrict__ xder, int64_t n, int64_t R) {
    __shared__ T tile[32][33];
    int64_t j0 = blockIdx.x * 32LL, r0 = blockIdx.y * 32LL;
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        int64_t j = j0 + tx, r = r0 + ty + k;
        if (j < n && r < R) tile[ty + k][tx] = xder[j + r * n];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        int64_t r = r0 + tx, j = j0 + ty + k;
        if (j < n && r < R) result[r + j * ld] = tile[tx][ty + k];
    }
}
template <typename T>
cudaError_t launch_rows_out(cudaStream_t s, T *result, int64_t ld, const T *xder, int64_t n, int64_t R) {
    if (n * R <= 0) return cudaSuccess;
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((R + 31) / 32));
    k_rows_out<T><<<grid, 256, 0, s>>>(result, ld, xder, n, R);
    return cudaGetLastError();
}

// ======================================================================================================
// forward-over-reverse tangent sweeps (hessian!)
// ======================================================================================================
template <typename T, int N>
struct TanArgs {
    const T *partial[N];
    const T *tv[N];
    const T *second[N];
};
template <typename T, int N>
__global__ void __launch_bounds__(kThreads) k_tangent_forward(T *__restrict__ out, TanArgs<T, N> a, int64_t n, int64_t K) {
    int64_t tot = n * K;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < tot; i += st) {
        int64_t e = i % n;
        T acc = (T)0;
#pragma unroll
        for (int p = 0; p < N; ++p) if (a.tv[p]) acc += a.partial[p][e] * a.tv[p][i];     // an untracked argument has no tangent
        out[i] = acc;
    }
}
template <typename T>
cudaError_t launch_tangent_forward(cudaStream_t s, T *tv_out, int n_args, const T *const *partial, const T *const *tv_in, int64_t n, int64_t K) {
    if (n * K <= 0) return cudaSuccess;
    int grid = grid_for(n * K, 2);
#define RDB_TF(NN)                                                                     \
    case NN: {                                                                         \
        TanArgs<T, NN> a;                                                              \
        for (int p = 0; p < NN; ++p) { a.partial[p] = partial[p]; a.tv[p] = tv_in[p]; a.second[p] = nullptr; } \
        k_tangent_forward<T, NN><<<grid, kThreads, 0, s>>>(tv_out, a, n, K);           \
    } break;
    switch (n_args) { RDB_TF(1) RDB_TF(2) RDB_TF(3) RDB_TF(4) default: return cudaErrorInvalidValue; }
#undef RDB_TF
    return cudaGetLastError();
}

template <typename T, int N, bool OVERWRITE>
__global__ void __launch_bounds__(kThreads)
k_tangent_reverse(T *__restrict__ td, const T *__restrict__ dt, const T *__restrict__ delta, const T *__restrict__ partial_p,
                  TanArgs<T, N> a, int64_t n, int64_t K) {
    int64_t tot = n * K;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < tot; i += st) {
        int64_t e = i % n;
        T acc = (T)0;
#pragma unroll
        for (int q = 0; q < N; ++q) if (a.tv[q]) acc += a.second[q][e] * a.tv[q][i];      // an untracked argument has no tangent
        T v = delta[e] * acc;
        if (dt) v += dt[i] * partial_p[e];          // dt == nullptr: the output's adjoint tangent is known to be zero
        td[i] = OVERWRITE ? v : td[i] + v;
    }
}
template <typename T>
cudaError_t launch_tangent_reverse(cudaStream_t s, T *td_p, const T *dt, const T *delta, const T *partial_p, int n_args,
                                   const T *const *second_pq, const T *const *tv_in, int64_t n, int64_t K, bool overwrite) {
    if (n * K <= 0) return cudaSuccess;
    int grid = grid_for(n * K, 2);
#define RDB_TR(NN)                                                                                      \
    case NN: {                                                                                          \
        TanArgs<T, NN> a;                                                                               \
        for (int q = 0; q < NN; ++q) { a.partial[q] = nullptr; a.tv[q] = tv_in[q]; a.second[q] = second_pq[q]; } \
        if (overwrite) k_tangent_reverse<T, NN, true><<<grid, kThreads, 0, s>>>(td_p, dt, delta, partial_p, a, n, K); \
        else k_tangent_reverse<T, NN, false><<<grid, kThreads, 0, s>>>(td_p, dt, delta, partial_p, a, n, K);          \
    } break;
    switch (n_args) { RDB_TR(1) RDB_TR(2) RDB_TR(3) RDB_TR(4) default: return cudaErrorInvalidValue; }
#undef RDB_TR
    return cudaGetLastError();
}

// The same two sweeps for closures whose arguments BROADCAST against the output (propagation.jl:25-27 clamps: an argument
// dimension of length 1 keeps index 0): argument q's element for output element e is g_q(e) = sum_d idx_d(e) * stride[q][d]
// (stride 0 along clamped dimensions), its tangents live at g_q(e) + k * size[q].  Into a clamped argument many output elements
// add: atomics (only this generality path uses them; same-shape closures keep the kernels above).
template <typename T>
struct TanArgsG {
    const T *partial[kMaxArgs];
    const T *tv[kMaxArgs];
    const T *second[kMaxArgs];
    int64_t stride[kMaxArgs][kMaxDims];
    int64_t size[kMaxArgs];
    int64_t dims[kMaxDims];
    int n_args;
};
template <typename T>
__device__ __forceinline__ int64_t g_index(const TanArgsG<T> &a, int q, int64_t e) {
    int64_t off = 0, rem = e;
#pragma unroll
    for (int d = 0; d < kMaxDims; ++d) { off += (rem % a.dims[d]) * a.stride[q][d]; rem /= a.dims[d]; }
    return off;
}
template <typename T>
__global__ void __launch_bounds__(kThreads) k_tangent_forward_g(T *__restrict__ out, TanArgsG<T> a, int64_t n, int64_t K) {
    int64_t tot = n * K;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < tot; i += st) {
        const int64_t e = i % n, k = i / n;
        T acc = (T)0;
        for (int p = 0; p < a.n_args; ++p)
            if (a.tv[p]) acc += a.partial[p][e] * a.tv[p][g_index(a, p, e) + k * a.size[p]];
        out[i] = acc;
    }
}
template <typename T>
cudaError_t launch_tangent_forward_g(cudaStream_t s, T *tv_out, int n_args, const T *const *partial, const T *const *tv_in,
                                     const int64_t (*stride)[kMaxDims], const int64_t *sizes, const int64_t *dims, int64_t n, int64_t K) {
    if (n * K <= 0) return cudaSuccess;
    TanArgsG<T> a{};
    a.n_args = n_args;
    for (int d = 0; d < kMaxDims; ++d) a.dims[d] = dims[d];
    for (int p = 0; p < n_args; ++p) {
        a.partial[p] = partial[p]; a.tv[p] = tv_in[p]; a.second[p] = nullptr; a.size[p] = sizes[p];
        for (int d = 0; d < kMaxDims; ++d) a.stride[p][d] = stride[p][d];
    }
    k_tangent_forward_g<T><<<grid_for(n * K, 2), kThreads, 0, s>>>(tv_out, a, n, K);
    return cudaGetLastError();
}
// td_p[g_p(e), k] += delta[e] * sum_q second_pq[e] * tv_q[g_q(e), k] + dt[e, k] * partial_p[e]; `p_direct`: argument p has the
// output's shape (one writer per element, plain read-modify-write), otherwise atomicAdd.  The destination must hold zeros or the
// running sum already (the caller clears it when this is the first accumulation).
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_tangent_reverse_g(T *__restrict__ td, const T *__restrict__ dt, const T *__restrict__ delta, const T *__restrict__ partial_p,
                    TanArgsG<T> a, int p, int p_direct, int64_t n, int64_t K) {
    int64_t tot = n * K;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < tot; i += st) {
        const int64_t e = i % n, k = i / n;
        T acc = (T)0;
        for (int q = 0; q < a.n_args; ++q)
            if (a.tv[q]) acc += a.second[q][e] * a.tv[q][g_index(a, q, e) + k * a.size[q]];
        T v = delta[e] * acc;
        if (dt) v += dt[i] * partial_p[e];
        const int64_t dst = g_index(a, p, e) + k * a.size[p];
        if (p_direct) td[dst] += v;
        else atomicAdd(td + dst, v);
    }
}
template <typename T>
cudaError_t launch_tangent_reverse_g(cudaStream_t s, T *td_p, const T *dt, const T *delta, const T *partial_p, int n_args, int p,
                                     const T *const *second_pq, const T *const *tv_in, const int64_t (*stride)[kMaxDims],
                                     const int64_t *sizes, const int64_t *dims, int64_t n, int64_t K) {
    if (n * K <= 0) return cudaSuccess;
    TanArgsG<T> a{};
    a.n_args = n_args;
    for (int d = 0; d < kMaxDims; ++d) a.dims[d] = dims[d];
    for (int q = 0; q < n_args; ++q) {
        a.partial[q] = nullptr; a.tv[q] = tv_in[q]; a.second[q] = second_pq[q]; a.size[q] = sizes[q];
        for (int d = 0; d < kMaxDims; ++d) a.stride[q][d] = stride[q][d];
    }
    k_tangent_reverse_g<T><<<grid_for(n * K, 2), kThreads, 0, s>>>(td_p, dt, delta, partial_p, a, p, sizes[p] == n ? 1 : 0, n, K);
    return cudaGetLastError();
}

// Sparse second-order term for an elementwise instruction whose arguments are all `getindex` views of the seeded input: the value
// tangent of argument q is the one-hot selector (map_q[i] == c0 + k), so
//   td_x[map_p[i] + k*n_x] += delta[i] * second_pq[i]   only at k = map_q[i] - c0.
// One thread per (i, p, q): O(n N^2) work instead of O(n K N) — the dense result block is just memset + these scattered adds.
template <typename T>
struct SparseArgs {
    const int64_t *map[kMaxArgs];
    const T *second[kMaxArgs][kMaxArgs];
};
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_tangent_reverse_sparse(T *td_x, int64_t n_x, int64_t K, int64_t c0, const T *__restrict__ delta, SparseArgs<T> a, int64_t n, int N) {
    int64_t tot = n * N * N;
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; t < tot; t += st) {
        int64_t i = t % n;
        int pq = (int)(t / n), p = pq / N, q = pq % N;
        int64_t k = a.map[q][i] - c0;
        if (k < 0 || k >= K) continue;
        atomicAdd(td_x + a.map[p][i] + k * n_x, delta[i] * a.second[p][q][i]);
    }
}
template <typename T>
cudaError_t launch_tangent_reverse_sparse(cudaStream_t s, T *td_x, int64_t n_x, int64_t K, int64_t c0, const T *delta, int n_args,
                                          const int64_t *const *maps, const T *const *second_pq, int64_t n) {
    if (n <= 0) return cudaSuccess;
    SparseArgs<T> a;
    for (int p = 0; p < n_args; ++p) {
        a.map[p] = maps[p];
        for (int q = 0; q < n_args; ++q) a.second[p][q] = second_pq[p * n_args + q];
    }
    k_tangent_reverse_sparse<T><<<grid_for(n * n_args * n_args, 1), kThreads, 0, s>>>(td_x, n_x, K, c0, delta, a, n, n_args);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------- hessian! of a separable objective, one kernel
// Two kernels chained by programmatic dependent launch.  The call's HBM traffic is a device fill of the block (7.5 TB/s for a plain
// fill kernel; a fused kernel whose CTAs also evaluated entries reached 5.3-5.7), so the fill stays a plain fill: many small CTAs
// in address order, nothing else in them.  It triggers the launch of its dependent at once; the patch kernel - one lane per column -
// evaluates the second partials of the elements that read x[column] while the fill is still running, waits for the fill to be
// complete and visible (griddepcontrol.wait), and adds its few values in tape order.  No atomics: a column belongs to one lane.
template <typename T>
__global__ void __launch_bounds__(kThreads) k_fill_block_pdl(T *__restrict__ H, int64_t n, int64_t K, int64_t ld) {
    // its own dependent (the patch kernel) may start at once: it only reads x until its griddepcontrol.wait.  This kernel is itself
    // launched as a programmatic dependent of whatever kernel precedes it on the stream (back-to-back hessian! calls: the previous
    // call's patch kernel, which triggers at its start): the CTAs become resident early and wait here, so no launch latency sits
    // between two calls, and nothing is written before the previous kernel is complete
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    using VT = typename Vec<T>::type;
    constexpr int V = Vec<T>::N;
    T zz[V] = {};
    const VT z = vec_pack<T>(zz);
    if (ld == n && ((((uintptr_t)H) & 15) == 0)) {                           // the block is one contiguous range
        VT *dst = reinterpret_cast<VT *>(H);
        const int64_t tot = K * n, nv = tot / V;
        const int64_t base = (int64_t)blockIdx.x * (4 * kThreads) + threadIdx.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) { const int64_t i = base + j * kThreads; if (i < nv) dst[i] = z; }
        if (blockIdx.x == 0) for (int64_t j = nv * V + threadIdx.x; j < tot; j += kThreads) H[j] = (T)0;
    } else {
        const int64_t per_col = (n + 4 * kThreads - 1) / (4 * kThreads);    // CTAs per column
        const int64_t k = blockIdx.x / per_col, r0 = (blockIdx.x % per_col) * (4 * kThreads);
        if (k < K) for (int j = 0; j < 4; ++j) { const int64_t i = r0 + j * kThreads + threadIdx.x; if (i < n) H[k * ld + i] = (T)0; }
    }
}
template <typename T, int N>
__global__ void __launch_bounds__(128) k_hessian_patch(const SepHessParams P) {
    asm volatile("griddepcontrol.launch_dependents;");                       // (the next call's fill kernel waits for this grid to complete)
    __shared__ int4 s_ops[kMaxExprRegs];
    for (int i = threadIdx.x; i < P.n_ops; i += blockDim.x) s_ops[i] = reinterpret_cast<const int4 *>(P.ops)[i];
    __syncthreads();
    T *H = reinterpret_cast<T *>(P.H);
    const T *x = reinterpret_cast<const T *>(P.x);
    constexpr int kMaxHeld = 4;                      // entries a lane keeps in registers; longer lists are evaluated after the wait
    int64_t h_row[kMaxHeld * N];
    T h_val[kMaxHeld * N];
    int held = 0;
    auto eval_entry = [&](int64_t e, int64_t *rows, T *vals) {
        // the element that reads x[c] as its argument q: row map[p][i] of this column gets delta * d2g / dx_p dx_q
        const int64_t iq = P.inv[e], i = iq / N;
        const int q = (int)(iq - i * N);
        T xs[N], val, dd[N], hh[Tri<N>::NH];
#pragma unroll
        for (int p = 0; p < N; ++p) { rows[p] = P.map[p] ? P.map[p][i] : i; xs[p] = x[rows[p]]; }
        eval_expr<T, N, 2>(s_ops, P.n_ops, P.fconst, xs, val, dd, hh);
#pragma unroll
        for (int p = 0; p < N; ++p) vals[p] = (T)P.delta * hh[p <= q ? tri(p, q, N) : tri(q, p, N)];
    };
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t e0 = 0, e1 = 0;
    if (k < P.K) {
        e0 = P.inv_off[P.c0 + k]; e1 = P.inv_off[P.c0 + k + 1];
        if (e1 - e0 <= kMaxHeld)
            for (int64_t e = e0; e < e1; ++e) { eval_entry(e, h_row + held * N, h_val + held * N); ++held; }
        else held = -1;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");                       // the fill is complete and its zeros are visible
    if (k >= P.K) return;
    T *col = H + k * P.ld;
    if (held >= 0) {
        // the column was zero-filled and belongs to this lane: values that share a row are summed in registers (tape order) and
        // every touched row is stored once - no read-modify-write round trip after the wait
#pragma unroll
        for (int e = 0; e < kMaxHeld * N; ++e) {
            if (e >= held * N) break;
            bool first = true;
#pragma unroll
            for (int f = 0; f < kMaxHeld * N; ++f) if (f < e && h_row[f] == h_row[e]) first = false;
            if (!first) continue;
            T acc = h_val[e];
#pragma unroll
            for (int f = 0; f < kMaxHeld * N; ++f) if (f > e && f < held * N && h_row[f] == h_row[e]) acc += h_val[f];
            col[h_row[e]] = acc;
        }
        return;
    }
    for (int64_t e = e0; e < e1; ++e) {
        int64_t rows[N]; T vals[N];
        eval_entry(e, rows, vals);
#pragma unroll
        for (int p = 0; p < N; ++p) col[rows[p]] += vals[p];
    }
}
template <typename T, int N>
static cudaError_t launch_patch_n(cudaStream_t s, const SepHessParams &P) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((P.K + 127) / 128)); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_hessian_patch<T, N>, P);
}
template <typename T>
cudaError_t launch_hessian_separable(cudaStream_t s, const SepHessParams &P) {
    if (P.K <= 0) return cudaSuccess;
    T *H = reinterpret_cast<T *>(P.H);
    const bool contiguous = P.ld == P.n_x && ((((uintptr_t)H) & 15) == 0);
    const int64_t per_cta = 4 * kThreads * (contiguous ? Vec<T>::N : 1);
    const int64_t ctas = contiguous ? (P.K * P.n_x + per_cta - 1) / per_cta : P.K * ((P.n_x + per_cta - 1) / per_cta);
    if (ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)ctas); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_fill_block_pdl<T>, H, P.n_x, P.K, P.ld);
        if (e != cudaSuccess) return e;
    }
    switch (P.n_args) {
        case 1: return launch_patch_n<T, 1>(s, P);
        case 2: return launch_patch_n<T, 2>(s, P);
        case 3: return launch_patch_n<T, 3>(s, P);
        case 4: return launch_patch_n<T, 4>(s, P);
    }
    return cudaErrorInvalidValue;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) k_colsum(T *__restrict__ out, const T *__restrict__ x, int64_t n, int64_t K, double scale) {
    int64_t k = blockIdx.x;
    if (k >= K) return;
    double local = 0.0;
    const T *col = x + k * n;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) local += (double)col[i];
    double s = block_sum(local);
    if (threadIdx.x == 0) out[k] = (T)(scale * s);
}
template <typename T>
cudaError_t launch_colsum(cudaStream_t s, T *out, const T *x, int64_t n, int64_t K, T scale) {
    if (K <= 0) return cudaSuccess;
    k_colsum<T><<<(unsigned)K, kThreads, 0, s>>>(out, x, n, K, (double)scale);
    return cudaGetLastError();
}

template <typename T>
__global__ void __launch_bounds__(kThreads) k_copy2d(T *__restrict__ dst, int64_t ldd, const T *__restrict__ src, int64_t lds, int64_t n, int64_t K) {
    int64_t tot = n * K;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, st = (int64_t)gridDim.x * blockDim.x;
    for (; i < tot; i += st) {
        int64_t k = i / n, e = i - k * n;
        dst[e + k * ldd] = src[e + k * lds];
    }
}
template <typename T>
cudaError_t launch_copy2d(cudaStream_t s, T *dst, int64_t ldd, const T *src, int64_t lds, int64_t n, int64_t K) {
    if (n * K <= 0) return cudaSuccess;
    k_copy2d<T><<<grid_for(n * K, 4), kThreads, 0, s>>>(dst, ldd, src, lds, n, K);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ explicit instantiation
#define RDB_INST(T)                                                                                                       \
    template cudaError_t launch_fill<T>(cudaStream_t, T *, int64_t, T);                                                   \
    template cudaError_t launch_add<T>(cudaStream_t, T *, const T *, const T *, T, int64_t, T *);                         \
    template cudaError_t launch_block_sums<T>(cudaStream_t, const T *, int64_t, T *);                                     \
    template cudaError_t launch_finish_sum<T>(cudaStream_t, const T *, T *, T, bool);                                     \
    template cudaError_t launch_dot_block_sums<T>(cudaStream_t, const T *, const T *, int64_t, T *);                      \
    template cudaError_t launch_scal_acc<T>(cudaStream_t, T *, const T *, const T *, int64_t, int64_t, bool);             \
    template cudaError_t launch_bcast_acc<T>(cudaStream_t, T *, const int64_t *, const int64_t *, const T *, int64_t, int64_t, int64_t, bool); \
    template cudaError_t launch_axpy<T>(cudaStream_t, T *, const T *, T, int64_t, bool);                                  \
    template cudaError_t launch_acc_scalar<T>(cudaStream_t, T *, int64_t, int64_t, const T *, T, bool);                   \
    template cudaError_t launch_expr_forward<T>(cudaStream_t, const ExprParams &);                                        \
    template cudaError_t launch_mul_acc<T>(cudaStream_t, T *, const T *, const T *, int64_t, int64_t, bool);              \
    template cudaError_t launch_mul_acc_colvec<T>(cudaStream_t, T *, const T *, const T *, int64_t, int64_t, int64_t, T *, int64_t, bool); \
    template cudaError_t launch_mul_acc_colvec_fused<T>(cudaStream_t, T *, const T *, const T *, int64_t, int64_t, int64_t, T *, int64_t, bool, T *, const T *, bool, bool, T); \
    template cudaError_t launch_mul_acc_sdelta<T>(cudaStream_t, T *, const T *, T, const T *, int64_t, int64_t, bool); \
    template cudaError_t launch_mul_acc_generic<T>(cudaStream_t, T *, int64_t, const int64_t *, const T *, const T *, const int64_t *, int64_t, int64_t); \
    template cudaError_t launch_gather<T>(cudaStream_t, T *, const T *, const int64_t *, int64_t);                        \
    template cudaError_t launch_scatter_add<T>(cudaStream_t, T *, int64_t, const int64_t *, const T *, int64_t, int64_t); \
    template cudaError_t launch_seed_onehot<T>(cudaStream_t, T *, int64_t, int64_t, int64_t);                             \
    template cudaError_t launch_rows_out<T>(cudaStream_t, T *, int64_t, const T *, int64_t, int64_t);                     \
    template cudaError_t launch_tangent_forward<T>(cudaStream_t, T *, int, const T *const *, const T *const *, int64_t, int64_t); \
    template cudaError_t launch_tangent_reverse<T>(cudaStream_t, T *, const T *, const T *, const T *, int, const T *const *, const T *const *, int64_t, int64_t, bool); \
    template cudaError_t launch_tangent_forward_g<T>(cudaStream_t, T *, int, const T *const *, const T *const *, const int64_t (*)[kMaxDims], const int64_t *, const int64_t *, int64_t, int64_t); \
    template cudaError_t launch_tangent_reverse_g<T>(cudaStream_t, T *, const T *, const T *, const T *, int, int, const T *const *, const T *const *, const int64_t (*)[kMaxDims], const int64_t *, const int64_t *, int64_t, int64_t); \
    template cudaError_t launch_gather_cols<T>(cudaStream_t, T *, const T *, int64_t, const int64_t *, int64_t, int64_t); \
    template cudaError_t launch_tangent_reverse_sparse<T>(cudaStream_t, T *, int64_t, int64_t, int64_t, const T *, int, const int64_t *const *, const T *const *, int64_t); \
    template cudaError_t launch_hessian_separable<T>(cudaStream_t, const SepHessParams &);                               \
    template cudaError_t launch_seed_tangent<T>(cudaStream_t, T *, int64_t, int64_t, int64_t);                            \
    template cudaError_t launch_seed_gather<T>(cudaStream_t, T *, const int64_t *, int64_t, int64_t, int64_t);            \
    template cudaError_t launch_colsum<T>(cudaStream_t, T *, const T *, int64_t, int64_t, T);                             \
    template cudaError_t launch_copy2d<T>(cudaStream_t, T *, int64_t, const T *, int64_t, int64_t, int64_t);
RDB_INST(double)
RDB_INST(float)

}  // namespace rdb
