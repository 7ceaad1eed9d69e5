// NVRTC specialisation of elementwise instructions (kernel family a).
//
// The bytecode interpreter in kernels.cu keeps its SSA registers in local memory and dispatches every scalar op through a switch;
// with FP64 tanh on top it ran the MLP config's `∇broadcast` forward sweeps at ~1.4 TB/s.  At rdb_tape_load the same straight-line
// program is emitted as CUDA C — one Dual{N} evaluation per element in exactly the interpreter's arithmetic order, shapes and
// broadcast strides baked in as literals, partials of untracked arguments never computed — compiled for sm_100a with NVRTC
// (--fmad=false, like kernels.cu) and launched through the driver API.  No NVRTC at run time (dlopen fails) => the interpreter
// stays in use; both paths are covered by the same parity tests (RDB_JIT=0 forces the interpreter).
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "kernels.h"

namespace rdb {
namespace jit {

// ---- NVRTC through dlopen (no link-time dependency) ---------------------------------------------------------
typedef int nvrtcResult;
typedef struct _nvrtcProgram *nvrtcProgram;
struct Nvrtc {
    nvrtcResult (*createProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
    nvrtcResult (*compileProgram)(nvrtcProgram, int, const char *const *);
    nvrtcResult (*getCUBINSize)(nvrtcProgram, size_t *);
    nvrtcResult (*getCUBIN)(nvrtcProgram, char *);
    nvrtcResult (*getProgramLogSize)(nvrtcProgram, size_t *);
    nvrtcResult (*getProgramLog)(nvrtcProgram, char *);
    nvrtcResult (*destroyProgram)(nvrtcProgram *);
    bool ok = false;
};
static Nvrtc &nvrtc() {
    static Nvrtc n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    const char *e = getenv("RDB_JIT");
    if (e && e[0] == '0') return n;
    const char *names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    void *h = nullptr;
    for (const char *nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (h) break; }
    if (!h) return n;
#define RDB_SYM(field, sym) n.field = (decltype(n.field))dlsym(h, sym); if (!n.field) return n;
    RDB_SYM(createProgram, "nvrtcCreateProgram")
    RDB_SYM(compileProgram, "nvrtcCompileProgram")
    RDB_SYM(getCUBINSize, "nvrtcGetCUBINSize")
    RDB_SYM(getCUBIN, "nvrtcGetCUBIN")
    RDB_SYM(getProgramLogSize, "nvrtcGetProgramLogSize")
    RDB_SYM(getProgramLog, "nvrtcGetProgramLog")
    RDB_SYM(destroyProgram, "nvrtcDestroyProgram")
#undef RDB_SYM
    n.ok = true;
    return n;
}

// ---- driver API through cudaGetDriverEntryPoint ---------------------------------------------------------------
struct Drv {
    CUresult (*moduleLoadData)(CUmodule *, const void *);
    CUresult (*moduleGetFunction)(CUfunction *, CUmodule, const char *);
    CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **);
    bool ok = false;
};
static Drv &drv() {
    static Drv d;
    static bool tried = false;
    if (tried) return d;
    tried = true;
    cudaDriverEntryPointQueryResult q;
    void *p = nullptr;
#define RDB_DRV(field, name)                                                                                               \
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return d; \
    d.field = (decltype(d.field))p;
    RDB_DRV(moduleLoadData, "cuModuleLoadData")
    RDB_DRV(moduleGetFunction, "cuModuleGetFunction")
    RDB_DRV(launchKernel, "cuLaunchKernel")
#undef RDB_DRV
    d.ok = true;
    return d;
}

// ---- code generation ----------------------------------------------------------------------------------------------
static std::string lit(double c, bool f32) {
    char buf[64];
    snprintf(buf, sizeof(buf), f32 ? "%.9gf" : "%.17g", c);
    std::string s(buf);
    if (s.find_first_of(".en") == std::string::npos && !f32) s += ".0";          // integer-valued: keep it a double literal
    if (f32 && s.find_first_of(".e") == std::string::npos) s.insert(s.size() - 1, ".0");
    if (s.find("inf") != std::string::npos || s.find("nan") != std::string::npos) return "";   // not expressible: no JIT
    return "(" + s + ")";
}

struct Gen {
    std::ostringstream o;
    int N;
    bool f32;
    std::string V(int r) { return "v" + std::to_string(r); }
    std::string D(int r, int p) { return "d" + std::to_string(r) + "_" + std::to_string(p); }
    std::string one() { return f32 ? "1.0f" : "1.0"; }
    std::string zero() { return f32 ? "0.0f" : "0.0"; }
    std::string fn(const char *name) { return std::string(name) + (f32 ? "f" : ""); }
    void unary(int r, int a, const std::string &f0, const std::string &f1) {
        o << "    const T f0_" << r << " = " << f0 << ";\n    const T f1_" << r << " = " << f1 << ";\n";
        o << "    const T " << V(r) << " = f0_" << r << ";\n";
        for (int p = 0; p < N; ++p) o << "    const T " << D(r, p) << " = f1_" << r << " * " << D(a, p) << ";\n";
    }
};

// One elementwise instruction inside the element loop of a generated kernel.  `pfx` names its pointers (`<pfx>a0`, `<pfx>out`,
// `<pfx>p0`); src[a] non-empty: the argument is the value an earlier instruction of the same fused run computed for this very element
// (a register), not a load.  The instruction's value is left in `<pfx>y` (declared by the caller).  false: cannot be specialised.
static bool emit_instruction(Gen &g, const ExprParams &P, const int32_t *ops, const double *fconst, bool f32, const std::string &pfx,
                             const std::string *src) {
    const int N = P.n_args;
    g.N = N;
    auto &o = g.o;
    const bool small = P.n <= 0x7fffffffLL;
    o << "    {\n";
    for (int a = 0; a < N; ++a) {
        if (src && !src[a].empty()) { o << "    const T v" << a << " = " << src[a] << ";\n"; continue; }
        if (P.arg[a].same) { o << "    const T v" << a << " = " << pfx << "a" << a << "[i];\n"; continue; }
        // broadcast clamp (propagation.jl:25-27) as literal strides
        o << "    T v" << a << ";\n    {\n";
        if (small) o << "      unsigned rem = (unsigned)i; unsigned off = 0u;\n";
        else o << "      long long rem = i; long long off = 0;\n";
        for (int d = 0; d < kMaxDims; ++d) {
            const long long dim = P.dims[d], st = P.arg[a].stride[d];
            if (dim == 1) continue;
            const char *sfx = small ? "u" : "LL";
            if (st != 0) o << "      off += (rem % " << dim << sfx << ") * " << st << sfx << ";\n";
            o << "      rem /= " << dim << sfx << ";\n";
        }
        o << "      v" << a << " = " << pfx << "a" << a << "[off];\n    }\n";
    }
    for (int a = 0; a < N; ++a)
        for (int p = 0; p < N; ++p) o << "    const T " << g.D(a, p) << " = " << (a == p ? g.one() : g.zero()) << ";\n";
    int r = N;
    for (int k = 0; k < P.n_ops; ++k, ++r) {
        const int code = ops[4 * k], a = ops[4 * k + 1], b = ops[4 * k + 2], imm = ops[4 * k + 3];
        switch (code) {
            case E_CONST: {
                std::string c = lit(fconst[imm], f32);
                if (c.empty()) return false;
                o << "    const T " << g.V(r) << " = (T)" << c << ";\n";
                for (int p = 0; p < N; ++p) o << "    const T " << g.D(r, p) << " = " << g.zero() << ";\n";
            } break;
            case E_ARG:
                o << "    const T " << g.V(r) << " = " << g.V(imm) << ";\n";
                for (int p = 0; p < N; ++p) o << "    const T " << g.D(r, p) << " = " << g.D(imm, p) << ";\n";
                break;
            case E_ADD: case E_SUB: {
                const char *op = code == E_ADD ? " + " : " - ";
                o << "    const T " << g.V(r) << " = " << g.V(a) << op << g.V(b) << ";\n";
                for (int p = 0; p < N; ++p) o << "    const T " << g.D(r, p) << " = " << g.D(a, p) << op << g.D(b, p) << ";\n";
            } break;
            case E_MUL:
                o << "    const T " << g.V(r) << " = " << g.V(a) << " * " << g.V(b) << ";\n";
                for (int p = 0; p < N; ++p)
                    o << "    const T " << g.D(r, p) << " = " << g.D(a, p) << " * " << g.V(b) << " + " << g.V(a) << " * " << g.D(b, p) << ";\n";
                break;
            case E_DIV:
                o << "    const T " << g.V(r) << " = " << g.V(a) << " / " << g.V(b) << ";\n";
                o << "    const T ib_" << r << " = " << g.one() << " / " << g.V(b) << ";\n";
                for (int p = 0; p < N; ++p)
                    o << "    const T " << g.D(r, p) << " = (" << g.D(a, p) << " - " << g.V(r) << " * " << g.D(b, p) << ") * ib_" << r << ";\n";
                break;
            case E_POW: {
                // the three ForwardDiff `^` methods / the (broadcast,^) formulas, chosen statically from imm (expr_eval.cuh pow_dual)
                const int flavor = imm & 3;
                const bool a_plain = (imm & 4) != 0, b_plain = (imm & 8) != 0;
                const std::string R = std::to_string(r), z = g.zero();
                o << "    const T " << g.V(r) << " = " << g.fn("pow") << "(" << g.V(a) << ", " << g.V(b) << ");\n";
                if (a_plain && b_plain) {
                    for (int p = 0; p < N; ++p) o << "    const T " << g.D(r, p) << " = " << z << ";\n";
                    break;
                }
                o << "    const T pw_" << R << " = " << g.fn("pow") << "(" << g.V(a) << ", " << g.V(b) << " - " << g.one() << ");\n";
                if (flavor == 2) {
                    o << "    const T fa_" << R << " = " << g.V(b) << " * pw_" << R << ";\n";
                    o << "    const T fb_" << R << " = " << g.fn("log") << "(" << g.V(a) << ") * " << g.V(r) << ";\n";
                    for (int p = 0; p < N; ++p)
                        o << "    const T " << g.D(r, p) << " = (" << g.D(a, p) << " != " << z << " ? fa_" << R << " * " << g.D(a, p) << " : " << z
                          << ") + (" << g.D(b, p) << " != " << z << " ? fb_" << R << " * " << g.D(b, p) << " : " << z << ");\n";
                    break;
                }
                if (b_plain) {
                    o << "    const bool az_" << R << " = true";
                    for (int p = 0; p < N; ++p) o << " && " << g.D(a, p) << " == " << z;
                    o << ";\n";
                    for (int p = 0; p < N; ++p)
                        o << "    const T " << g.D(r, p) << " = (" << g.V(b) << " == " << z << " || "
                          << (flavor == 1 ? g.D(a, p) + " == " + z : "az_" + R) << ") ? " << z << " : (" << g.D(a, p) << " * " << g.V(b) << ") * pw_" << R << ";\n";
                    break;
                }
                o << "    const T lg_" << R << " = (" << g.V(a) << " == " << z << " && " << g.V(b) << " > " << z << ") ? " << z << " : "
                  << g.V(r) << " * " << g.fn("log") << "(" << g.V(a) << ");\n";
                if (a_plain) {
                    for (int p = 0; p < N; ++p) o << "    const T " << g.D(r, p) << " = lg_" << R << " * " << g.D(b, p) << ";\n";
                    break;
                }
                o << "    const T pv_" << R << " = " << g.V(b) << " * pw_" << R << ";\n";
                o << "    const bool cs_" << R << " = true";
                for (int p = 0; p < N; ++p) o << " && " << g.D(b, p) << " == " << z;
                o << ";\n";
                for (int p = 0; p < N; ++p)
                    o << "    const T " << g.D(r, p) << " = " << g.D(a, p) << " * pv_" << R << " + " << g.D(b, p) << " * (("
                      << (flavor == 1 ? g.D(b, p) + " == " + z : "cs_" + R) << ") ? " << g.one() << " : lg_" << R << ");\n";
            } break;
            case E_MAX: case E_MIN:
                o << "    const bool pk_" << r << " = " << g.V(a) << (code == E_MAX ? " > " : " < ") << g.V(b) << ";\n";
                o << "    const T " << g.V(r) << " = pk_" << r << " ? " << g.V(a) << " : " << g.V(b) << ";\n";
                for (int p = 0; p < N; ++p)
                    o << "    const T " << g.D(r, p) << " = pk_" << r << " ? " << g.D(a, p) << " : " << g.D(b, p) << ";\n";
                break;
            case E_NEG: g.unary(r, a, "-" + g.V(a), "-" + g.one()); break;
            case E_POWI:
                g.unary(r, a, "powi_(" + g.V(a) + ", " + std::to_string(imm) + ")",
                        imm != 0 ? "(T)" + std::to_string(imm) + " * powi_(" + g.V(a) + ", " + std::to_string(imm - 1) + ")" : g.zero());
                break;
            case E_TANH:
                o << "    const T t_" << r << " = " << g.fn("tanh") << "(" << g.V(a) << ");\n";
                g.unary(r, a, "t_" + std::to_string(r), g.one() + " - t_" + std::to_string(r) + " * t_" + std::to_string(r));
                break;
            case E_EXP:
                o << "    const T t_" << r << " = " << g.fn("exp") << "(" << g.V(a) << ");\n";
                g.unary(r, a, "t_" + std::to_string(r), "t_" + std::to_string(r));
                break;
            case E_LOG: g.unary(r, a, g.fn("log") + "(" + g.V(a) + ")", g.one() + " / " + g.V(a)); break;
            case E_SQRT:
                o << "    const T t_" << r << " = " << g.fn("sqrt") << "(" << g.V(a) << ");\n";
                g.unary(r, a, "t_" + std::to_string(r), g.one() + " / ((T)2 * t_" + std::to_string(r) + ")");
                break;
            case E_SIN: g.unary(r, a, g.fn("sin") + "(" + g.V(a) + ")", g.fn("cos") + "(" + g.V(a) + ")"); break;
            case E_COS: g.unary(r, a, g.fn("cos") + "(" + g.V(a) + ")", "-" + g.fn("sin") + "(" + g.V(a) + ")"); break;
            case E_ABS2: g.unary(r, a, g.V(a) + " * " + g.V(a), "(T)2 * " + g.V(a)); break;
            case E_ABS:
                g.unary(r, a, g.fn("fabs") + "(" + g.V(a) + ")", "signbit(" + g.V(a) + ") ? -" + g.one() + " : " + g.one());
                break;
            case E_INV:
                o << "    const T t_" << r << " = " << g.one() << " / " << g.V(a) << ";\n";
                g.unary(r, a, "t_" + std::to_string(r), "-t_" + std::to_string(r) + " * t_" + std::to_string(r));
                break;
            case E_SIGMOID:
                o << "    const T t_" << r << " = " << g.one() << " / (" << g.one() << " + " << g.fn("exp") << "(-" << g.V(a) << "));\n";
                g.unary(r, a, "t_" + std::to_string(r), "t_" + std::to_string(r) + " * (" + g.one() + " - t_" + std::to_string(r) + ")");
                break;
            case E_LOG1P: g.unary(r, a, g.fn("log1p") + "(" + g.V(a) + ")", g.one() + " / (" + g.V(a) + " + " + g.one() + ")"); break;
            case E_EXPM1: g.unary(r, a, g.fn("expm1") + "(" + g.V(a) + ")", g.fn("exp") + "(" + g.V(a) + ")"); break;
            case E_ASIN: g.unary(r, a, g.fn("asin") + "(" + g.V(a) + ")", g.one() + " / " + g.fn("sqrt") + "(" + g.one() + " - " + g.V(a) + " * " + g.V(a) + ")"); break;
            case E_ACOS: g.unary(r, a, g.fn("acos") + "(" + g.V(a) + ")", "-(" + g.one() + " / " + g.fn("sqrt") + "(" + g.one() + " - " + g.V(a) + " * " + g.V(a) + "))"); break;
            case E_ATAN: g.unary(r, a, g.fn("atan") + "(" + g.V(a) + ")", g.one() + " / (" + g.one() + " + " + g.V(a) + " * " + g.V(a) + ")"); break;
            case E_SINH: g.unary(r, a, g.fn("sinh") + "(" + g.V(a) + ")", g.fn("cosh") + "(" + g.V(a) + ")"); break;
            case E_COSH: g.unary(r, a, g.fn("cosh") + "(" + g.V(a) + ")", g.fn("sinh") + "(" + g.V(a) + ")"); break;
            case E_TAN:
                o << "    const T t_" << r << " = " << g.fn("tan") << "(" << g.V(a) << ");\n";
                g.unary(r, a, "t_" + std::to_string(r), g.one() + " + t_" + std::to_string(r) + " * t_" + std::to_string(r));
                break;
            case E_EXP2: case E_EXP10: {
                const std::string R = std::to_string(r), l = code == E_EXP2 ? "(T)0.69314718055994530942" : "(T)2.30258509299404568402";
                o << "    const T t_" << R << " = " << g.fn(code == E_EXP2 ? "exp2" : "exp10") << "(" << g.V(a) << ");\n";
                g.unary(r, a, "t_" + R, "t_" + R + " * " + l);
            } break;
            case E_LOG2: g.unary(r, a, g.fn("log2") + "(" + g.V(a) + ")", g.one() + " / " + g.V(a) + " / (T)0.69314718055994530942"); break;
            case E_LOG10: g.unary(r, a, g.fn("log10") + "(" + g.V(a) + ")", g.one() + " / " + g.V(a) + " / (T)2.30258509299404568402"); break;
            case E_CBRT: {
                const std::string R = std::to_string(r);
                o << "    const T t_" << R << " = " << g.fn("cbrt") << "(" << g.V(a) << ");\n";
                g.unary(r, a, "t_" + R, g.one() + " / ((T)3 * (t_" + R + " * t_" + R + "))");
            } break;
            case E_ASINH: g.unary(r, a, g.fn("asinh") + "(" + g.V(a) + ")", g.one() + " / " + g.fn("sqrt") + "(" + g.V(a) + " * " + g.V(a) + " + " + g.one() + ")"); break;
            case E_ACOSH: g.unary(r, a, g.fn("acosh") + "(" + g.V(a) + ")", g.one() + " / " + g.fn("sqrt") + "(" + g.V(a) + " * " + g.V(a) + " - " + g.one() + ")"); break;
            case E_ATANH: g.unary(r, a, g.fn("atanh") + "(" + g.V(a) + ")", g.one() + " / (" + g.one() + " - " + g.V(a) + " * " + g.V(a) + ")"); break;
            case E_SEC: case E_CSC: case E_COT: case E_SECH: case E_CSCH: case E_COTH: {
                // the reciprocal functions as Julia defines them: t = inv(cos | sin | tan | cosh | sinh | tanh)
                const std::string R = std::to_string(r), t = "t_" + R, u = "u_" + R;
                const char *inner = code == E_SEC ? "cos" : code == E_CSC ? "sin" : code == E_COT ? "tan" : code == E_SECH ? "cosh" : code == E_CSCH ? "sinh" : "tanh";
                o << "    const T " << t << " = " << g.one() << " / " << g.fn(inner) << "(" << g.V(a) << ");\n";
                if (code == E_SEC) { o << "    const T " << u << " = " << g.fn("tan") << "(" << g.V(a) << ");\n"; g.unary(r, a, t, t + " * " + u); }
                else if (code == E_CSC) { o << "    const T " << u << " = " << g.one() << " / " << g.fn("tan") << "(" << g.V(a) << ");\n"; g.unary(r, a, t, "-(" + t + " * " + u + ")"); }
                else if (code == E_COT) g.unary(r, a, t, "-(" + g.one() + " + " + t + " * " + t + ")");
                else if (code == E_SECH) { o << "    const T " << u << " = " << g.fn("tanh") << "(" << g.V(a) << ");\n"; g.unary(r, a, t, "-(" + u + " * " + t + ")"); }
                else if (code == E_CSCH) { o << "    const T " << u << " = " << g.one() << " / " << g.fn("tanh") << "(" << g.V(a) << ");\n"; g.unary(r, a, t, "-(" + u + " * " + t + ")"); }
                else { o << "    const T " << u << " = " << g.one() << " / " << g.fn("sinh") << "(" << g.V(a) << ");\n"; g.unary(r, a, t, "-(" + u + " * " + u + ")"); }
            } break;
            case E_ATAN2: case E_HYPOT: {
                const std::string R = std::to_string(r);
                if (code == E_ATAN2) {
                    o << "    const T r2_" << R << " = " << g.V(a) << " * " << g.V(a) << " + " << g.V(b) << " * " << g.V(b) << ";\n";
                    o << "    const T " << g.V(r) << " = " << g.fn("atan2") << "(" << g.V(a) << ", " << g.V(b) << ");\n";
                    o << "    const T ga_" << R << " = " << g.V(b) << " / r2_" << R << ";\n    const T gb_" << R << " = -" << g.V(a) << " / r2_" << R << ";\n";
                } else {
                    o << "    const T " << g.V(r) << " = " << g.fn("hypot") << "(" << g.V(a) << ", " << g.V(b) << ");\n";
                    o << "    const T ga_" << R << " = " << g.V(a) << " / " << g.V(r) << ";\n    const T gb_" << R << " = " << g.V(b) << " / " << g.V(r) << ";\n";
                }
                for (int p = 0; p < N; ++p)
                    o << "    const T " << g.D(r, p) << " = ga_" << R << " * " << g.D(a, p) << " + gb_" << R << " * " << g.D(b, p) << ";\n";
            } break;
            default: return false;
        }
    }
    o << "    " << pfx << "out[i] = " << g.V(r - 1) << ";\n";
    for (int a = 0; a < N; ++a) if (P.partial[a]) o << "    " << pfx << "p" << a << "[i] = " << g.D(r - 1, a) << ";\n";
    o << "    " << pfx << "y = " << g.V(r - 1) << ";\n";
    o << "    }\n";
    return true;
}

static void emit_prologue(Gen &g, bool f32) {
    g.o << "typedef " << (f32 ? "float" : "double") << " T;\n";
    g.o << "__device__ __forceinline__ T powi_(T x, int n) { if (n == 0) return (T)1; bool neg = n < 0; if (neg) n = -n; T r = x; "
           "for (int i = 1; i < n; ++i) r = r * x; return neg ? (T)1 / r : r; }\n";
    g.o << "extern \"C\" __global__ void __launch_bounds__(256) expr_k(";
}
static void emit_params(Gen &g, const ExprParams &P, const std::string &pfx, const std::string *src) {
    for (int a = 0; a < P.n_args; ++a) if (!src || src[a].empty()) g.o << "const T* __restrict__ " << pfx << "a" << a << ", ";
    g.o << "T* __restrict__ " << pfx << "out, ";
    for (int a = 0; a < P.n_args; ++a) if (P.partial[a]) g.o << "T* __restrict__ " << pfx << "p" << a << ", ";
}
static void emit_epilogue(Gen &g, bool reduce) {
    auto &o = g.o;
    o << "  }\n";
    if (reduce) {
        o << "  __shared__ double ws[8];\n"
             "  for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);\n"
             "  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = local;\n"
             "  __syncthreads();\n"
             "  if (threadIdx.x < 32) {\n"
             "    double v = threadIdx.x < 8 ? ws[threadIdx.x] : 0.0;\n"
             "    for (int s = 4; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);\n"
             "    if (threadIdx.x == 0) bsums[blockIdx.x] = v;\n"
             "  }\n";
    }
    o << "}\n";
}

// returns "" when the program cannot be specialised
static std::string generate(const ExprParams &P, const int32_t *ops, const double *fconst, bool f32, bool reduce) {
    Gen g;
    g.f32 = f32;
    emit_prologue(g, f32);
    emit_params(g, P, "", nullptr);
    g.o << "double* __restrict__ bsums) {\n  double local = 0.0;\n";
    g.o << "  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < " << P.n << "LL; i += gridDim.x * 256LL) {\n    T y;\n";
    if (!emit_instruction(g, P, ops, fconst, f32, "", nullptr)) return "";
    if (reduce) g.o << "    local += (double)y;\n";
    emit_epilogue(g, reduce);
    return g.o.str();
}

// A run of adjacent same-shape elementwise instructions as ONE kernel (north star (a)): every instruction still stores its value and
// its partials - the reverse sweep and `value(x)` see exactly what the separate kernels would have left - but a value consumed by a
// later instruction of the run travels in a register instead of through HBM, and the run is one launch.
static std::string generate_run(const JitRunItem *items, int n_items, const double *fconst, bool f32, bool reduce) {
    Gen g;
    g.f32 = f32;
    emit_prologue(g, f32);
    std::vector<std::vector<std::string>> src(n_items, std::vector<std::string>(kMaxArgs));
    for (int k = 0; k < n_items; ++k) {
        for (int a = 0; a < items[k].p.n_args; ++a)
            if (items[k].from_item[a] >= 0) src[k][a] = "s" + std::to_string(items[k].from_item[a]) + "_y";
        emit_params(g, items[k].p, "s" + std::to_string(k) + "_", src[k].data());
    }
    g.o << "double* __restrict__ bsums) {\n  double local = 0.0;\n";
    g.o << "  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < " << items[0].p.n << "LL; i += gridDim.x * 256LL) {\n";
    for (int k = 0; k < n_items; ++k) {
        g.o << "    T s" << k << "_y;\n";
        if (!emit_instruction(g, items[k].p, items[k].ops, fconst, f32, "s" + std::to_string(k) + "_", src[k].data())) return "";
    }
    if (reduce) g.o << "    local += (double)s" << (n_items - 1) << "_y;\n";
    emit_epilogue(g, reduce);
    return g.o.str();
}

static std::map<std::string, CUfunction> g_cache;

static CUfunction compile(const std::string &src) {
    auto it = g_cache.find(src);
    if (it != g_cache.end()) return it->second;
    g_cache[src] = nullptr;
    Nvrtc &n = nvrtc();
    Drv &d = drv();
    if (!n.ok || (!d.ok && !getenv("RDB_JIT_DUMP"))) return nullptr;
    nvrtcProgram prog = nullptr;
    if (n.createProgram(&prog, src.c_str(), "rdb_expr.cu", 0, nullptr, nullptr) != 0) return nullptr;
    const char *opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "-default-device", "--std=c++17"};
    nvrtcResult rc = n.compileProgram(prog, 4, opts);
    if (getenv("RDB_JIT_DUMP")) fprintf(stderr, "[rdb jit] NVRTC rc=%d for\n%s\n", rc, src.c_str());
    if (rc != 0) {
        if (getenv("RDB_JIT_VERBOSE")) {
            size_t ls = 0;
            n.getProgramLogSize(prog, &ls);
            std::vector<char> log(ls + 1, 0);
            n.getProgramLog(prog, log.data());
            fprintf(stderr, "[rdb jit] NVRTC failed (%d):\n%s\n---- source ----\n%s\n", rc, log.data(), src.c_str());
        }
        n.destroyProgram(&prog);
        return nullptr;
    }
    size_t sz = 0;
    if (n.getCUBINSize(prog, &sz) != 0 || sz == 0) { n.destroyProgram(&prog); return nullptr; }
    std::vector<char> cubin(sz);
    n.getCUBIN(prog, cubin.data());
    n.destroyProgram(&prog);
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    if (!d.ok) return nullptr;
    if (d.moduleLoadData(&mod, cubin.data()) != CUDA_SUCCESS) return nullptr;
    if (d.moduleGetFunction(&fn, mod, "expr_k") != CUDA_SUCCESS) return nullptr;
    g_cache[src] = fn;
    return fn;
}

}  // namespace jit

// Build (or fetch) the specialised kernel for one elementwise instruction; nullptr => use the interpreter.
void *jit_expr_kernel(const ExprParams &p, const int32_t *host_ops, const double *host_fconst, bool f32, bool reduce) {
    if (p.order != 1) return nullptr;
    std::string src = jit::generate(p, host_ops, host_fconst, f32, reduce);
    if (src.empty()) return nullptr;
    return (void *)jit::compile(src);
}

void *jit_run_kernel(const JitRunItem *items, int n_items, const double *host_fconst, bool f32, bool reduce) {
    for (int k = 0; k < n_items; ++k) if (items[k].p.order != 1 || items[k].p.n != items[0].p.n) return nullptr;
    std::string src = jit::generate_run(items, n_items, host_fconst, f32, reduce);
    if (src.empty()) return nullptr;
    return (void *)jit::compile(src);
}

cudaError_t jit_run_launch(cudaStream_t st, void *fn, const JitRunItem *items, int n_items, void *block_sums, bool reduce) {
    std::vector<void *> args;
    std::vector<const void *> ptrs;
    ptrs.reserve((size_t)n_items * (2 * kMaxArgs + 1) + 1);
    for (int k = 0; k < n_items; ++k) {
        const ExprParams &p = items[k].p;
        for (int i = 0; i < p.n_args; ++i) if (items[k].from_item[i] < 0) ptrs.push_back(p.arg[i].ptr);
        ptrs.push_back(p.out);
        for (int i = 0; i < p.n_args; ++i) if (p.partial[i]) ptrs.push_back(p.partial[i]);
    }
    ptrs.push_back(block_sums);
    for (auto &q : ptrs) args.push_back((void *)&q);
    const int64_t n = items[0].p.n;
    int64_t blocks = reduce ? kReduceBlocks : (n + 511) / 512;
    if (!reduce && blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    CUresult r = jit::drv().launchKernel((CUfunction)fn, (unsigned)blocks, 1, 1, 256, 1, 1, 0, (CUstream)st, args.data(), nullptr);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorLaunchFailure;
}

cudaError_t jit_expr_launch(cudaStream_t st, void *fn, const ExprParams &p, bool reduce) {
    void *args[2 * kMaxArgs + 3];
    const void *a[kMaxArgs];
    void *pp[kMaxArgs];
    void *out = p.out, *bs = p.block_sums;
    int n = 0;
    for (int i = 0; i < p.n_args; ++i) { a[i] = p.arg[i].ptr; args[n++] = &a[i]; }
    args[n++] = &out;
    for (int i = 0; i < p.n_args; ++i) if (p.partial[i]) { pp[i] = p.partial[i]; args[n++] = &pp[i]; }
    args[n++] = &bs;
    int64_t blocks = reduce ? kReduceBlocks : (p.n + 511) / 512;
    if (!reduce && blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    CUresult r = jit::drv().launchKernel((CUfunction)fn, (unsigned)blocks, 1, 1, 256, 1, 1, 0, (CUstream)st, args, nullptr);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorLaunchFailure;
}

}  // namespace rdb
