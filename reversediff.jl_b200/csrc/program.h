// Wire format of the flat tape program (mirror of reversediff.jl_b200/program.py; see that file for
// the reference citations of every field).  Plain packed PODs, little-endian, 8-byte aligned sections.
#pragma once
#include <cstdint>

namespace rdb {

constexpr uint64_t kMagic = 0x3030324244525052ull;
constexpr uint32_t kVersion = 2;
constexpr int kWireArgs = 4;   // operand slots of one instruction record on the wire
constexpr int kMaxArgs = 8;    // operands of one instruction (records with more than kWireArgs continue in OP_MORE_OPERANDS records)
constexpr int kMaxDims = 4;
constexpr int kMaxExprRegs = 32;

enum BufKind : int32_t { BUF_TA = 0, BUF_TR = 1, BUF_CONST = 2 };
enum OperandClass : int32_t { CLS_TA = 0, CLS_IDX = 1, CLS_CONST = 2, CLS_TR = 3 };
enum IdxKind : int32_t { IDX_NONE = 0, IDX_EXPLICIT = 1, IDX_TRANSPOSE = 2 };

enum Opcode : int32_t {
    OP_MUL = 1, OP_ADD = 2, OP_SUB = 3, OP_NEG = 4, OP_SUM = 5, OP_MEAN = 6, OP_NBCAST = 7, OP_MAP = 8,
    OP_BCAST = 9, OP_BCAST_ADD = 10, OP_BCAST_SUB = 11, OP_BCAST_MUL = 12, OP_BCAST_DIV = 13, OP_BCAST_LDIV = 14, OP_BCAST_POW = 15,
    OP_GETINDEX = 16, OP_DOT = 17, OP_SUMDIMS = 18, OP_TRACKER_BCAST = 19, OP_GETINDEX_GENERIC = 20,
    OP_SCALAR_BLOCK = 21,   // a run of ScalarInstructions; payload layout in program.py (ScalarBlock)
    OP_MORE_OPERANDS = 22,  // not an instruction: operands 5.. of the record before it (∇broadcast is N-ary, broadcast.jl:142-146)
};

enum ExprOp : int32_t {
    E_CONST = 1, E_ADD, E_SUB, E_MUL, E_DIV, E_NEG, E_POWI, E_POW, E_TANH, E_EXP, E_LOG, E_SQRT, E_SIN, E_COS,
    E_ABS2, E_ABS, E_MAX, E_MIN, E_INV, E_SIGMOID, E_ARG,
    E_LOG1P, E_EXPM1, E_ASIN, E_ACOS, E_ATAN, E_SINH, E_COSH, E_TAN, E_ATAN2, E_HYPOT,
    E_EXP2, E_EXP10, E_LOG2, E_LOG10, E_CBRT, E_ASINH, E_ACOSH, E_ATANH, E_SEC, E_CSC, E_COT, E_SECH, E_CSCH, E_COTH,
    E_LAST = E_COTH,
};

#pragma pack(push, 1)
struct Header {
    uint64_t magic;
    uint32_t version, dtype;
    uint32_t n_buffers, n_instr, n_inputs;
    int32_t output_buffer;
    uint64_t off_buffers, off_inputs, off_instr;
    uint64_t off_idx, n_idx;
    uint64_t off_expr, n_expr;
    uint64_t off_fconst, n_fconst;
    uint64_t off_const, n_const;
    uint64_t total_bytes;
};
static_assert(sizeof(Header) == 128, "Header layout");

struct BufferRec {
    int32_t kind, ndims;
    int64_t dims[kMaxDims];
    int32_t alias_of, pad;
    int64_t const_offset;
};
static_assert(sizeof(BufferRec) == 56, "BufferRec layout");

struct OperandRec {
    int32_t buffer, cls, tracked, idx_kind;
    int64_t idx_offset, idx_len;
    int32_t ndims, pad;
    int64_t dims[kMaxDims];
    int64_t bound[kMaxDims];
};
static_assert(sizeof(OperandRec) == 104, "OperandRec layout");

struct InstrRec {
    int32_t opcode, n_in, out, n_expr_args;
    int64_t expr_offset, expr_len;
    OperandRec in[kWireArgs];
};
static_assert(sizeof(InstrRec) == 32 + 4 * 104, "InstrRec layout");
#pragma pack(pop)

}  // namespace rdb
