/* ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/replay.py for the rules).
 *
 * C restatement of the two loops the reference runs over a tape of ScalarInstructions, so that the CPU baseline of the
 * scalar-tape legs is compiled code (like the reference's Julia) and not a Python-interpreter benchmark:
 *
 *   forward   scalar_forward_exec!   src/derivatives/scalars.jl:80-123   (pull_value!, one ForwardDiff.derivative!/gradient!
 *                                    evaluation, value!(output), cache[] = partials), in tape order  src/tape.jl:75-80
 *   reverse   scalar_reverse_exec!   src/derivatives/scalars.jl:52-74    (increment_deriv!(input, deriv(output) * partial);
 *                                    unseed!(output)), in reverse tape order                        src/tape.jl:85-90
 *
 * ForwardDiff (Project.toml:25, "0.10, 1") / DiffRules ("1.4") are not under /root/reference; the Dual arithmetic restated
 * here is the same rule table as oracle/replay.py dual_eval, which the tests cross-check against these loops.
 *
 * The instruction table is the int64 body of the wire format (reversediff.jl_b200/program.py ScalarBlock):
 *   (expr_id, a_space, a_idx, b_space, b_idx); space -1 = slot of this run, -2 = literal, -3 = absent, >= 0 = ref buffer.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o libscalar_block.so scalar_block.c -lm   (no FMA contraction: Julia does
 * not contract a*b+c either).
 */
#include <math.h>
#include <stdint.h>

enum { E_CONST = 1, E_ADD, E_SUB, E_MUL, E_DIV, E_NEG, E_POWI, E_POW, E_TANH, E_EXP, E_LOG, E_SQRT, E_SIN, E_COS,
       E_ABS2, E_ABS, E_MAX, E_MIN, E_INV, E_SIGMOID, E_ARG,
       E_LOG1P, E_EXPM1, E_ASIN, E_ACOS, E_ATAN, E_SINH, E_COSH, E_TAN, E_ATAN2, E_HYPOT,
       E_EXP2, E_EXP10, E_LOG2, E_LOG10, E_CBRT, E_ASINH, E_ACOSH, E_ATANH, E_SEC, E_CSC, E_COT, E_SECH, E_CSCH, E_COTH };
enum { SP_POOL = -1, SP_LITERAL = -2, SP_NONE = -3 };
#define MAX_REGS 34

static double powi(double x, int n) {   /* Base.literal_pow spelled as the repeated product the device uses */
    if (n == 0) return 1.0;
    int neg = n < 0;
    if (neg) n = -n;
    double r = x;
    for (int i = 1; i < n; ++i) r = r * x;
    return neg ? 1.0 / r : r;
}

/* one Dual{2} evaluation: value and the partials w.r.t. argument 0 and 1 */
static int dual2(const int32_t *ops, int64_t n_ops, const double *fc, const double *x, int n_args, double *val, double *d) {
    double V[MAX_REGS], D0[MAX_REGS], D1[MAX_REGS];
    for (int p = 0; p < 2; ++p) { V[p] = p < n_args ? x[p] : 0.0; D0[p] = p == 0; D1[p] = p == 1; }
    int r = n_args;
    for (int64_t o = 0; o < n_ops; ++o, ++r) {
        const int code = ops[4 * o], a = ops[4 * o + 1], b = ops[4 * o + 2], imm = ops[4 * o + 3];
        double f0 = 0, f1 = 0;
        int unary = 1;
        switch (code) {
            case E_CONST: V[r] = fc[imm]; D0[r] = D1[r] = 0; unary = 0; break;
            case E_ARG: V[r] = V[imm]; D0[r] = D0[imm]; D1[r] = D1[imm]; unary = 0; break;
            case E_ADD: V[r] = V[a] + V[b]; D0[r] = D0[a] + D0[b]; D1[r] = D1[a] + D1[b]; unary = 0; break;
            case E_SUB: V[r] = V[a] - V[b]; D0[r] = D0[a] - D0[b]; D1[r] = D1[a] - D1[b]; unary = 0; break;
            case E_MUL: { double va = V[a], vb = V[b]; V[r] = va * vb; D0[r] = D0[a] * vb + va * D0[b]; D1[r] = D1[a] * vb + va * D1[b]; unary = 0; } break;
            case E_DIV: { double q = V[a] / V[b], ib = 1.0 / V[b]; double t0 = (D0[a] - q * D0[b]) * ib, t1 = (D1[a] - q * D1[b]) * ib;
                          V[r] = q; D0[r] = t0; D1[r] = t1; unary = 0; } break;
            case E_POW: {
                /* ForwardDiff dual.jl "exponentiation": scalar instructions dualise TrackedReal arguments only (macros.jl:101-124);
                   imm bit 2 / bit 3 = base / exponent is a plain Real (flavor bits 0-1 are always 0 on a scalar tape) */
                double vx = V[a], vy = V[b], expv = pow(vx, vy), t0, t1;
                int x_real = (imm & 4) != 0, y_real = (imm & 8) != 0;
                if (x_real && y_real) { t0 = t1 = 0.0; }
                else if (y_real) {                       /* Dual^Real */
                    if (vy == 0.0 || (D0[a] == 0.0 && D1[a] == 0.0)) t0 = t1 = 0.0;
                    else { double pw = pow(vx, vy - 1.0); t0 = D0[a] * vy * pw; t1 = D1[a] * vy * pw; }
                } else if (x_real) {                     /* Real^Dual */
                    double deriv = (vx == 0.0 && vy > 0.0) ? 0.0 : expv * log(vx);
                    t0 = deriv * D0[b]; t1 = deriv * D1[b];
                } else {                                 /* Dual^Dual */
                    double powval = vy * pow(vx, vy - 1.0), logval;
                    if (D0[b] == 0.0 && D1[b] == 0.0) logval = 1.0;
                    else if (vx == 0.0 && vy > 0.0) logval = 0.0;
                    else logval = expv * log(vx);
                    t0 = D0[a] * powval + D0[b] * logval; t1 = D1[a] * powval + D1[b] * logval;
                }
                V[r] = expv; D0[r] = t0; D1[r] = t1; unary = 0; } break;
            case E_MAX: case E_MIN: { int pick_a = code == E_MAX ? (V[a] > V[b]) : (V[a] < V[b]); int s = pick_a ? a : b;
                          V[r] = V[s]; D0[r] = D0[s]; D1[r] = D1[s]; unary = 0; } break;
            case E_NEG: f0 = -V[a]; f1 = -1.0; break;
            case E_POWI: f0 = powi(V[a], imm); f1 = imm != 0 ? (double)imm * powi(V[a], imm - 1) : 0.0; break;
            case E_TANH: { double t = tanh(V[a]); f0 = t; f1 = 1.0 - t * t; } break;
            case E_EXP: { double e = exp(V[a]); f0 = e; f1 = e; } break;
            case E_LOG: f0 = log(V[a]); f1 = 1.0 / V[a]; break;
            case E_SQRT: { double s = sqrt(V[a]); f0 = s; f1 = 1.0 / (2.0 * s); } break;
            case E_SIN: f0 = sin(V[a]); f1 = cos(V[a]); break;
            case E_COS: f0 = cos(V[a]); f1 = -sin(V[a]); break;
            case E_ABS2: f0 = V[a] * V[a]; f1 = 2.0 * V[a]; break;
            case E_ABS: f0 = fabs(V[a]); f1 = signbit(V[a]) ? -1.0 : 1.0; break;   /* abs(d::Dual) = signbit(d) ? -d : d */
            case E_INV: { double rr = 1.0 / V[a]; f0 = rr; f1 = -rr * rr; } break;
            case E_SIGMOID: { double s = 1.0 / (1.0 + exp(-V[a])); f0 = s; f1 = s * (1.0 - s); } break;
            case E_LOG1P: f0 = log1p(V[a]); f1 = 1.0 / (V[a] + 1.0); break;
            case E_EXPM1: f0 = expm1(V[a]); f1 = exp(V[a]); break;
            case E_ASIN: f0 = asin(V[a]); f1 = 1.0 / sqrt(1.0 - V[a] * V[a]); break;
            case E_ACOS: f0 = acos(V[a]); f1 = -(1.0 / sqrt(1.0 - V[a] * V[a])); break;
            case E_ATAN: f0 = atan(V[a]); f1 = 1.0 / (1.0 + V[a] * V[a]); break;
            case E_SINH: f0 = sinh(V[a]); f1 = cosh(V[a]); break;
            case E_COSH: f0 = cosh(V[a]); f1 = sinh(V[a]); break;
            case E_TAN: { double tt = tan(V[a]); f0 = tt; f1 = 1.0 + tt * tt; } break;
            case E_EXP2: f0 = exp2(V[a]); f1 = f0 * 0.69314718055994530942; break;
            case E_EXP10: f0 = pow(10.0, V[a]); f1 = f0 * 2.30258509299404568402; break;
            case E_LOG2: f0 = log2(V[a]); f1 = 1.0 / V[a] / 0.69314718055994530942; break;
            case E_LOG10: f0 = log10(V[a]); f1 = 1.0 / V[a] / 2.30258509299404568402; break;
            case E_CBRT: { double c = cbrt(V[a]); f0 = c; f1 = 1.0 / (3.0 * (c * c)); } break;
            case E_ASINH: f0 = asinh(V[a]); f1 = 1.0 / sqrt(V[a] * V[a] + 1.0); break;
            case E_ACOSH: f0 = acosh(V[a]); f1 = 1.0 / sqrt(V[a] * V[a] - 1.0); break;
            case E_ATANH: f0 = atanh(V[a]); f1 = 1.0 / (1.0 - V[a] * V[a]); break;
            case E_SEC: { double sc = 1.0 / cos(V[a]); f0 = sc; f1 = sc * tan(V[a]); } break;
            case E_CSC: { double cs = 1.0 / sin(V[a]); f0 = cs; f1 = -(cs * (1.0 / tan(V[a]))); } break;
            case E_COT: { double ct = 1.0 / tan(V[a]); f0 = ct; f1 = -(1.0 + ct * ct); } break;
            case E_SECH: { double sh = 1.0 / cosh(V[a]); f0 = sh; f1 = -(tanh(V[a]) * sh); } break;
            case E_CSCH: { double ch = 1.0 / sinh(V[a]); f0 = ch; f1 = -((1.0 / tanh(V[a])) * ch); } break;
            case E_COTH: { double ch = 1.0 / sinh(V[a]); f0 = 1.0 / tanh(V[a]); f1 = -(ch * ch); } break;
            case E_ATAN2: { double va = V[a], vb = V[b], r2 = va * va + vb * vb, ga = vb / r2, gb = -va / r2;
                            double t0 = ga * D0[a] + gb * D0[b], t1 = ga * D1[a] + gb * D1[b];
                            V[r] = atan2(va, vb); D0[r] = t0; D1[r] = t1; unary = 0; } break;
            case E_HYPOT: { double va = V[a], vb = V[b], y = hypot(va, vb), ga = va / y, gb = vb / y;
                            double t0 = ga * D0[a] + gb * D0[b], t1 = ga * D1[a] + gb * D1[b];
                            V[r] = y; D0[r] = t0; D1[r] = t1; unary = 0; } break;
            default: return code ? code : -1;
        }
        if (unary) { V[r] = f0; D0[r] = f1 * D0[a]; D1[r] = f1 * D1[a]; }
    }
    *val = V[r - 1]; d[0] = D0[r - 1]; d[1] = D1[r - 1];
    return 0;
}

int scalar_block_forward(int64_t n, const int64_t *body, const int64_t *expr_table, const int32_t *expr_words, const double *fconst,
                         double *const *ref_val, double *val, double *pa, double *pb) {
    for (int64_t i = 0; i < n; ++i) {                         /* forward_pass!: tape order */
        const int64_t *I = body + 5 * i;
        const int64_t *E = expr_table + 3 * I[0];
        double x[2] = {0, 0};
        int na = 0;
        for (int k = 0; k < 2; ++k) {                         /* pull_value!(a); pull_value!(b) */
            const int64_t sp = I[1 + 2 * k], ix = I[2 + 2 * k];
            if (sp == SP_NONE) continue;
            x[na++] = sp == SP_POOL ? val[ix] : (sp == SP_LITERAL ? fconst[ix] : ref_val[sp][ix]);
        }
        double v, d[2];
        int rc = dual2(expr_words + E[0], E[1], fconst, x, (int)E[2], &v, d);
        if (rc) return rc;
        val[i] = v;                                           /* value!(output, ...) */
        pa[i] = d[0];                                         /* cache[] = partials */
        pb[i] = d[1];
    }
    return 0;
}

int scalar_block_reverse(int64_t n, const int64_t *body, double *const *ref_der, double *der, const double *pa, const double *pb) {
    for (int64_t i = n - 1; i >= 0; --i) {                    /* reverse_pass!: reverse tape order */
        const int64_t *I = body + 5 * i;
        const double d = der[i];
        for (int k = 0; k < 2; ++k) {
            const int64_t sp = I[1 + 2 * k], ix = I[2 + 2 * k];
            const double p = k == 0 ? pa[i] : pb[i];
            if (sp == SP_POOL) der[ix] = der[ix] + d * p;                   /* increment_deriv!(input, deriv(output) * partial) */
            else if (sp >= 0) ref_der[sp][ix] = ref_der[sp][ix] + d * p;    /* pull_deriv!, add, push_deriv! on origin.deriv[index] */
        }
        der[i] = 0.0;                                         /* unseed!(output) */
    }
    return 0;
}
