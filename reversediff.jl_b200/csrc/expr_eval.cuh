// Dual-number evaluation of the scalar-expression bytecode (program.h ExprOp): one ForwardDiff Dual{N} evaluation per call
// (value + N partials; ORDER == 2 also the packed second partials).  Shared by the elementwise sweeps (kernels.cu) and the
// scalar-instruction blocks (scalar.cu).  Compile the including file with -fmad=false (see kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "program.h"

namespace rdb {

template <typename T> struct M;
template <> struct M<double> {
    static __device__ __forceinline__ double tanh_(double x) { return tanh(x); }
    static __device__ __forceinline__ double exp_(double x) { return exp(x); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double sin_(double x) { return sin(x); }
    static __device__ __forceinline__ double cos_(double x) { return cos(x); }
    static __device__ __forceinline__ double pow_(double x, double y) { return pow(x, y); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
    static __device__ __forceinline__ double log1p_(double x) { return log1p(x); }
    static __device__ __forceinline__ double expm1_(double x) { return expm1(x); }
    static __device__ __forceinline__ double asin_(double x) { return asin(x); }
    static __device__ __forceinline__ double acos_(double x) { return acos(x); }
    static __device__ __forceinline__ double atan_(double x) { return atan(x); }
    static __device__ __forceinline__ double sinh_(double x) { return sinh(x); }
    static __device__ __forceinline__ double cosh_(double x) { return cosh(x); }
    static __device__ __forceinline__ double tan_(double x) { return tan(x); }
    static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
    static __device__ __forceinline__ double hypot_(double x, double y) { return hypot(x, y); }
    static __device__ __forceinline__ double exp2_(double x) { return exp2(x); }
    static __device__ __forceinline__ double exp10_(double x) { return exp10(x); }
    static __device__ __forceinline__ double log2_(double x) { return log2(x); }
    static __device__ __forceinline__ double log10_(double x) { return log10(x); }
    static __device__ __forceinline__ double cbrt_(double x) { return cbrt(x); }
    static __device__ __forceinline__ double asinh_(double x) { return asinh(x); }
    static __device__ __forceinline__ double acosh_(double x) { return acosh(x); }
    static __device__ __forceinline__ double atanh_(double x) { return atanh(x); }
};
template <> struct M<float> {
    static __device__ __forceinline__ float tanh_(float x) { return tanhf(x); }
    static __device__ __forceinline__ float exp_(float x) { return expf(x); }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float sin_(float x) { return sinf(x); }
    static __device__ __forceinline__ float cos_(float x) { return cosf(x); }
    static __device__ __forceinline__ float pow_(float x, float y) { return powf(x, y); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
    static __device__ __forceinline__ float log1p_(float x) { return log1pf(x); }
    static __device__ __forceinline__ float expm1_(float x) { return expm1f(x); }
    static __device__ __forceinline__ float asin_(float x) { return asinf(x); }
    static __device__ __forceinline__ float acos_(float x) { return acosf(x); }
    static __device__ __forceinline__ float atan_(float x) { return atanf(x); }
    static __device__ __forceinline__ float sinh_(float x) { return sinhf(x); }
    static __device__ __forceinline__ float cosh_(float x) { return coshf(x); }
    static __device__ __forceinline__ float tan_(float x) { return tanf(x); }
    static __device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
    static __device__ __forceinline__ float hypot_(float x, float y) { return hypotf(x, y); }
    static __device__ __forceinline__ float exp2_(float x) { return exp2f(x); }
    static __device__ __forceinline__ float exp10_(float x) { return exp10f(x); }
    static __device__ __forceinline__ float log2_(float x) { return log2f(x); }
    static __device__ __forceinline__ float log10_(float x) { return log10f(x); }
    static __device__ __forceinline__ float cbrt_(float x) { return cbrtf(x); }
    static __device__ __forceinline__ float asinh_(float x) { return asinhf(x); }
    static __device__ __forceinline__ float acosh_(float x) { return acoshf(x); }
    static __device__ __forceinline__ float atanh_(float x) { return atanhf(x); }
};

template <typename T>
__device__ __forceinline__ T powi(T x, int n) {   // Base.literal_pow: repeated product
    if (n == 0) return (T)1;
    bool neg = n < 0;
    if (neg) n = -n;
    T r = x;
    for (int i = 1; i < n; ++i) r = r * x;
    return neg ? (T)1 / r : r;
}

template <int N> struct Tri { static constexpr int NH = N * (N + 1) / 2; };
__device__ __forceinline__ int tri(int p, int q, int N) {   // p <= q packed index
    return p * N - p * (p - 1) / 2 + (q - p);
}

// `^` with Dual operands.  imm (expr.py pow_imm): bits 0-1 flavor, bit 2 base is a plain Real, bit 3 exponent is a plain Real.
// Flavors 0/1 are ForwardDiff's three methods (dual.jl, "exponentiation"):
//   Dual^Real  : (y == 0 || iszero(partials(x))) ? 0 : partials(x) * y * x^(y-1)          -- log(x) is never formed
//   Real^Dual  : ((x == 0 && y > 0) ? 0 : x^y * log(x)) * partials(y)
//   Dual^Dual  : partials(x) * (y x^(y-1)) + partials(y) * logval,  logval = isconstant(y) ? 1 : ((x == 0 && y > 0) ? 0 : x^y log(x))
// flavor 0 tests isconstant / iszero on the whole partials vector (one Dual{N} evaluation: ∇broadcast, @forward closures, scalar
// instructions), flavor 1 per component (tracker_∇broadcast evaluates one Dual{1} per argument, broadcast.jl:221-224); flavor 2
// is (broadcast,^): base partial e*b^(e-1) and exponent partial log(b)*b^e, cached separately (elementwise.jl:633-643).
template <typename T, int N>
__device__ __forceinline__ void pow_dual(T va, T vb, int imm, const T *Da, const T *Db, T &y, T *Dr) {
    const int flavor = imm & 3;
    const bool a_plain = (imm & 4) != 0, b_plain = (imm & 8) != 0;
    y = M<T>::pow_(va, vb);
    if (a_plain && b_plain) {
#pragma unroll
        for (int p = 0; p < N; ++p) Dr[p] = (T)0;
        return;
    }
    const T pw1 = M<T>::pow_(va, vb - (T)1);
    if (flavor == 2) {
        const T fa = vb * pw1, fb = M<T>::log_(va) * y;
#pragma unroll
        for (int p = 0; p < N; ++p) Dr[p] = (Da[p] != (T)0 ? fa * Da[p] : (T)0) + (Db[p] != (T)0 ? fb * Db[p] : (T)0);
        return;
    }
    if (b_plain) {
        bool all0 = true;
#pragma unroll
        for (int p = 0; p < N; ++p) all0 = all0 && Da[p] == (T)0;
#pragma unroll
        for (int p = 0; p < N; ++p) {
            const bool z = vb == (T)0 || (flavor == 1 ? Da[p] == (T)0 : all0);
            Dr[p] = z ? (T)0 : (Da[p] * vb) * pw1;
        }
        return;
    }
    const T lg = (va == (T)0 && vb > (T)0) ? (T)0 : y * M<T>::log_(va);
    if (a_plain) {
#pragma unroll
        for (int p = 0; p < N; ++p) Dr[p] = lg * Db[p];
        return;
    }
    const T powval = vb * pw1;
    bool cst = true;
#pragma unroll
    for (int p = 0; p < N; ++p) cst = cst && Db[p] == (T)0;
#pragma unroll
    for (int p = 0; p < N; ++p) {
        const T logval = (flavor == 1 ? Db[p] == (T)0 : cst) ? (T)1 : lg;
        Dr[p] = Da[p] * powval + Db[p] * logval;
    }
}

template <typename T, int N, int ORDER>
__device__ __forceinline__ void eval_expr(const int4 *__restrict__ ops, int n_ops, const double *__restrict__ fc,
                                          const T *x, T &val, T *dout, T *hout, int r0 = N) {
    constexpr int NH = Tri<N>::NH;
    T V[kMaxExprRegs];
    T D[kMaxExprRegs][N];
    T H[ORDER == 2 ? kMaxExprRegs : 1][NH];
#pragma unroll
    for (int p = 0; p < N; ++p) {
        V[p] = x[p];
#pragma unroll
        for (int q = 0; q < N; ++q) D[p][q] = (p == q) ? (T)1 : (T)0;
        if (ORDER == 2) {
#pragma unroll
            for (int h = 0; h < NH; ++h) H[p][h] = (T)0;
        }
    }
    int r = r0;     // op j writes register r0 + j: r0 = the expression's own argument count (<= N)
    for (int o = 0; o < n_ops; ++o, ++r) {
        const int4 op = ops[o];
        const int code = op.x, a = op.y, b = op.z, imm = op.w;
        T f0, f1 = (T)0, f2 = (T)0;
        bool unary = true;
        switch (code) {
            case E_CONST: {
                V[r] = (T)fc[imm];
#pragma unroll
                for (int p = 0; p < N; ++p) D[r][p] = (T)0;
                if (ORDER == 2) {
#pragma unroll
                    for (int h = 0; h < NH; ++h) H[r][h] = (T)0;
                }
                unary = false;
            } break;
            case E_ARG: {
                V[r] = V[imm];
#pragma unroll
                for (int p = 0; p < N; ++p) D[r][p] = D[imm][p];
                if (ORDER == 2) {
#pragma unroll
                    for (int h = 0; h < NH; ++h) H[r][h] = H[imm][h];
                }
                unary = false;
            } break;
            case E_ADD:
            case E_SUB: {
                const T s = code == E_ADD ? (T)1 : (T)-1;
                V[r] = code == E_ADD ? V[a] + V[b] : V[a] - V[b];
#pragma unroll
                for (int p = 0; p < N; ++p) D[r][p] = D[a][p] + s * D[b][p];
                if (ORDER == 2) {
#pragma unroll
                    for (int h = 0; h < NH; ++h) H[r][h] = H[a][h] + s * H[b][h];
                }
                unary = false;
            } break;
            case E_MUL: {
                const T va = V[a], vb = V[b];
                V[r] = va * vb;
                if (ORDER == 2) {
#pragma unroll
                    for (int p = 0; p < N; ++p)
#pragma unroll
                        for (int q = p; q < N; ++q)
                            H[r][tri(p, q, N)] = H[a][tri(p, q, N)] * vb + D[a][p] * D[b][q] + D[a][q] * D[b][p] +
                                                 va * H[b][tri(p, q, N)];
                }
#pragma unroll
                for (int p = 0; p < N; ++p) D[r][p] = D[a][p] * vb + va * D[b][p];
                unary = false;
            } break;
            case E_DIV: {
                const T q0 = V[a] / V[b], ib = (T)1 / V[b];
                T da[N];
#pragma unroll
                for (int p = 0; p < N; ++p) da[p] = (D[a][p] - q0 * D[b][p]) * ib;
                if (ORDER == 2) {
#pragma unroll
                    for (int p = 0; p < N; ++p)
#pragma unroll
                        for (int q = p; q < N; ++q)
                            H[r][tri(p, q, N)] = (H[a][tri(p, q, N)] - da[p] * D[b][q] - da[q] * D[b][p] -
                                                  q0 * H[b][tri(p, q, N)]) * ib;
                }
                V[r] = q0;
#pragma unroll
                for (int p = 0; p < N; ++p) D[r][p] = da[p];
                unary = false;
            } break;
            case E_POW: {
                const T va = V[a], vb = V[b];
                T y;
                pow_dual<T, N>(va, vb, imm, D[a], D[b], y, D[r]);
                V[r] = y;
                if (ORDER == 2) {
                    // y_pq = fa a_pq + fb b_pq + faa a_p a_q + fab (a_p b_q + a_q b_p) + fbb b_p b_q; a plain Real operand has no
                    // tangents, and its coefficients (log of a non-positive base...) must not touch the result
                    const bool b_off = (imm & 8) != 0;
                    const bool a_off = (imm & 4) != 0 || (b_off && vb == (T)0);          // Dual^Real with y == 0: constant
                    const T pw1 = M<T>::pow_(va, vb - (T)1), la = M<T>::log_(va);
                    const T fa = vb * pw1, fb = (va == (T)0 && vb > (T)0) ? (T)0 : y * la;
                    const T faa = vb * (vb - (T)1) * M<T>::pow_(va, vb - (T)2), fab = pw1 * ((T)1 + vb * la), fbb = fb * la;
#pragma unroll
                    for (int p = 0; p < N; ++p)
#pragma unroll
                        for (int q = p; q < N; ++q) {
                            T h = (T)0;
                            if (!a_off) h = fa * H[a][tri(p, q, N)] + faa * D[a][p] * D[a][q];
                            if (!b_off) h = h + fb * H[b][tri(p, q, N)] + fbb * D[b][p] * D[b][q];
                            if (!a_off && !b_off) h = h + fab * (D[a][p] * D[b][q] + D[a][q] * D[b][p]);
                            H[r][tri(p, q, N)] = h;
                        }
                }
                unary = false;
            } break;
            case E_MAX:
            case E_MIN: {
                const bool pick_a = code == E_MAX ? (V[a] > V[b]) : (V[a] < V[b]);
                const int s = pick_a ? a : b;
                V[r] = V[s];
#pragma unroll
                for (int p = 0; p < N; ++p) D[r][p] = D[s][p];
                if (ORDER == 2) {
#pragma unroll
                    for (int h = 0; h < NH; ++h) H[r][h] = H[s][h];
                }
                unary = false;
            } break;
            case E_NEG: f0 = -V[a]; f1 = (T)-1; f2 = (T)0; break;
            case E_POWI: {
                const T xx = V[a];
                f0 = powi(xx, imm);
                f1 = imm != 0 ? (T)imm * powi(xx, imm - 1) : (T)0;
                if (ORDER == 2) f2 = (imm != 0 && imm != 1) ? (T)(imm * (imm - 1)) * powi(xx, imm - 2) : (T)0;
            } break;
            case E_TANH: { const T t = M<T>::tanh_(V[a]); f0 = t; f1 = (T)1 - t * t; f2 = (T)-2 * t * f1; } break;
            case E_EXP: { const T e = M<T>::exp_(V[a]); f0 = e; f1 = e; f2 = e; } break;
            case E_LOG: { f0 = M<T>::log_(V[a]); f1 = (T)1 / V[a]; f2 = (T)-1 / (V[a] * V[a]); } break;
            case E_SQRT: { const T s = M<T>::sqrt_(V[a]); f0 = s; f1 = (T)1 / ((T)2 * s); f2 = (T)-1 / ((T)4 * s * V[a]); } break;
            case E_SIN: { f0 = M<T>::sin_(V[a]); f1 = M<T>::cos_(V[a]); f2 = -f0; } break;
            case E_COS: { f0 = M<T>::cos_(V[a]); f1 = -M<T>::sin_(V[a]); f2 = -f0; } break;
            case E_ABS2: { f0 = V[a] * V[a]; f1 = (T)2 * V[a]; f2 = (T)2; } break;
            case E_ABS: { f0 = M<T>::abs_(V[a]); f1 = signbit(V[a]) ? (T)-1 : (T)1; f2 = (T)0; } break;   // abs(d::Dual) = signbit(d) ? -d : d
            case E_INV: { const T rr = (T)1 / V[a]; f0 = rr; f1 = -rr * rr; f2 = (T)2 * rr * rr * rr; } break;
            case E_SIGMOID: { const T s = (T)1 / ((T)1 + M<T>::exp_(-V[a])); f0 = s; f1 = s * ((T)1 - s); f2 = f1 * ((T)1 - (T)2 * s); } break;
            // DiffRules: log1p' = inv(x + 1), expm1' = exp(x), asin' = inv(sqrt(1 - x^2)), acos' = -inv(sqrt(1 - x^2)),
            // atan' = inv(1 + x^2), sinh' = cosh(x), cosh' = sinh(x), tan' = 1 + tan(x)^2
            case E_LOG1P: { const T u = (T)1 / (V[a] + (T)1); f0 = M<T>::log1p_(V[a]); f1 = u; f2 = -u * u; } break;
            case E_EXPM1: { const T e = M<T>::exp_(V[a]); f0 = M<T>::expm1_(V[a]); f1 = e; f2 = e; } break;
            case E_ASIN: { const T w = (T)1 - V[a] * V[a], u = (T)1 / M<T>::sqrt_(w); f0 = M<T>::asin_(V[a]); f1 = u; f2 = V[a] * u / w; } break;
            case E_ACOS: { const T w = (T)1 - V[a] * V[a], u = (T)1 / M<T>::sqrt_(w); f0 = M<T>::acos_(V[a]); f1 = -u; f2 = -(V[a] * u / w); } break;
            case E_ATAN: { const T u = (T)1 / ((T)1 + V[a] * V[a]); f0 = M<T>::atan_(V[a]); f1 = u; f2 = (T)-2 * V[a] * u * u; } break;
            case E_SINH: { f0 = M<T>::sinh_(V[a]); f1 = M<T>::cosh_(V[a]); f2 = f0; } break;
            case E_COSH: { f0 = M<T>::cosh_(V[a]); f1 = M<T>::sinh_(V[a]); f2 = f0; } break;
            case E_TAN: { const T tt = M<T>::tan_(V[a]); f0 = tt; f1 = (T)1 + tt * tt; f2 = (T)2 * tt * f1; } break;
            // third tranche (DiffRules): exp2' = exp2(x) logtwo, exp10' = exp10(x) logten, log2' = inv(x) / logtwo, log10' = inv(x) / logten,
            // cbrt' = inv(3 cbrt(x)^2), asinh' = inv(sqrt(x^2 + 1)), acosh' = inv(sqrt(x^2 - 1)), atanh' = inv(1 - x^2),
            // sec' = sec(x) tan(x), csc' = -csc(x) cot(x), cot' = -(1 + cot(x)^2), sech' = -tanh(x) sech(x), csch' = -coth(x) csch(x),
            // coth' = -(csch(x)^2); the reciprocal functions are evaluated as Julia defines them (inv(cos(x)), ...)
            case E_EXP2: { const T l = (T)0.69314718055994530942; f0 = M<T>::exp2_(V[a]); f1 = f0 * l; f2 = f1 * l; } break;
            case E_EXP10: { const T l = (T)2.30258509299404568402; f0 = M<T>::exp10_(V[a]); f1 = f0 * l; f2 = f1 * l; } break;
            case E_LOG2: { const T l = (T)0.69314718055994530942, u = (T)1 / V[a]; f0 = M<T>::log2_(V[a]); f1 = u / l; f2 = -(f1 * u); } break;
            case E_LOG10: { const T l = (T)2.30258509299404568402, u = (T)1 / V[a]; f0 = M<T>::log10_(V[a]); f1 = u / l; f2 = -(f1 * u); } break;
            case E_CBRT: { const T c = M<T>::cbrt_(V[a]); f0 = c; f1 = (T)1 / ((T)3 * (c * c)); f2 = (T)-2 * f1 / ((T)3 * V[a]); } break;
            case E_ASINH: { const T u = (T)1 / M<T>::sqrt_(V[a] * V[a] + (T)1); f0 = M<T>::asinh_(V[a]); f1 = u; f2 = -(V[a] * u * u * u); } break;
            case E_ACOSH: { const T u = (T)1 / M<T>::sqrt_(V[a] * V[a] - (T)1); f0 = M<T>::acosh_(V[a]); f1 = u; f2 = -(V[a] * u * u * u); } break;
            case E_ATANH: { const T u = (T)1 / ((T)1 - V[a] * V[a]); f0 = M<T>::atanh_(V[a]); f1 = u; f2 = (T)2 * V[a] * u * u; } break;
            case E_SEC: { const T sc = (T)1 / M<T>::cos_(V[a]), tt = M<T>::tan_(V[a]); f0 = sc; f1 = sc * tt; f2 = sc * (tt * tt + sc * sc); } break;
            case E_CSC: { const T cs = (T)1 / M<T>::sin_(V[a]), ct = (T)1 / M<T>::tan_(V[a]); f0 = cs; f1 = -(cs * ct); f2 = cs * (ct * ct + cs * cs); } break;
            case E_COT: { const T ct = (T)1 / M<T>::tan_(V[a]); f0 = ct; f1 = -((T)1 + ct * ct); f2 = (T)-2 * ct * f1; } break;
            case E_SECH: { const T sh = (T)1 / M<T>::cosh_(V[a]), th = M<T>::tanh_(V[a]); f0 = sh; f1 = -(th * sh); f2 = sh * (th * th - sh * sh); } break;
            case E_CSCH: { const T ch = (T)1 / M<T>::sinh_(V[a]), ct = (T)1 / M<T>::tanh_(V[a]); f0 = ch; f1 = -(ct * ch); f2 = ch * (ct * ct + ch * ch); } break;
            case E_COTH: { const T ct = (T)1 / M<T>::tanh_(V[a]), ch = (T)1 / M<T>::sinh_(V[a]); f0 = ct; f1 = -(ch * ch); f2 = (T)-2 * ct * f1; } break;
            case E_ATAN2:
            case E_HYPOT: {
                // y = g(a, b) with partials (ga, gb) and second partials (gaa, gab, gbb):
                //   atan(y, x): (x / (y^2 + x^2), -y / (y^2 + x^2))   [DiffRules atan(x, y) with the roles of its names]
                //   hypot(x, y): (x / hypot, y / hypot)
                const T va = V[a], vb = V[b];
                T y, ga, gb, gaa = (T)0, gab = (T)0, gbb = (T)0;
                if (code == E_ATAN2) {
                    const T r2 = va * va + vb * vb;
                    y = M<T>::atan2_(va, vb); ga = vb / r2; gb = -va / r2;
                    if (ORDER == 2) { const T i4 = (T)1 / (r2 * r2); gaa = (T)-2 * va * vb * i4; gbb = (T)2 * va * vb * i4; gab = (va * va - vb * vb) * i4; }
                } else {
                    y = M<T>::hypot_(va, vb); ga = va / y; gb = vb / y;
                    if (ORDER == 2) { const T i3 = (T)1 / (y * y * y); gaa = vb * vb * i3; gbb = va * va * i3; gab = -(va * vb) * i3; }
                }
                if (ORDER == 2) {
#pragma unroll
                    for (int p = 0; p < N; ++p)
#pragma unroll
                        for (int q = p; q < N; ++q)
                            H[r][tri(p, q, N)] = ga * H[a][tri(p, q, N)] + gb * H[b][tri(p, q, N)] + gaa * D[a][p] * D[a][q] +
                                                 gab * (D[a][p] * D[b][q] + D[a][q] * D[b][p]) + gbb * D[b][p] * D[b][q];
                }
                T dn[N];
#pragma unroll
                for (int p = 0; p < N; ++p) dn[p] = ga * D[a][p] + gb * D[b][p];
                V[r] = y;
#pragma unroll
                for (int p = 0; p < N; ++p) D[r][p] = dn[p];
                unary = false;
            } break;
            default: f0 = (T)0; break;
        }
        if (unary) {
            if (ORDER == 2) {
#pragma unroll
                for (int p = 0; p < N; ++p)
#pragma unroll
                    for (int q = p; q < N; ++q)
                        H[r][tri(p, q, N)] = f2 * D[a][p] * D[a][q] + f1 * H[a][tri(p, q, N)];
            }
            V[r] = f0;
#pragma unroll
            for (int p = 0; p < N; ++p) D[r][p] = f1 * D[a][p];
        }
    }
    val = V[r - 1];
#pragma unroll
    for (int p = 0; p < N; ++p) dout[p] = D[r - 1][p];
    if (ORDER == 2) {
#pragma unroll
        for (int h = 0; h < NH; ++h) hout[h] = H[r - 1][h];
    }
}

}  // namespace rdb
