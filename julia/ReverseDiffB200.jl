# ReverseDiffB200.jl — the reference-side binding a ReverseDiff maintainer would add so that
# `ReverseDiffB200.compile(tape)` replays GradientTape / JacobianTape / HessianTape on librdb200.so (B200, sm_100a).
#
# STATUS: complete but UNEXECUTED.  There is no `julia` binary in the build image nor on the GPU boxes, so this file has never
# run; the same lowering written in Python (reversediff.jl_b200/tracked.py + program.py) is what the test-suite exercises.  The
# contract between the two is byte-level: `lower(tape)` must emit exactly the bytes `TapeProgram.to_bytes()` emits for the same
# recorded function; tests/golden/*.program.sha256 hold the SHA-256 of the Python recorder's programs for the five BASELINE
# configs at small sizes with exactly-representable constants, and julia/golden_roundtrip.jl is the one-assert check to run the
# day a Julia runtime is at hand.
#
# What it does
#   1. lower(t)   walks `t.tape::Vector{AbstractInstruction}` (src/tape.jl:5-7,30-69): every TrackedArray gets a buffer id (keyed by
#                 the identity of its value array; `reshape` aliases share storage, src/tracked.jl:402-408), every index-linked
#                 Array{TrackedReal} operand (`capture`, src/tracked.jl:223-225) becomes an IDX operand with the explicit map
#                 `x[k].index - 1` (src/tracked.jl:47-57,348-351), plain arrays become CONST buffers, elementwise closures are run
#                 ONCE on symbolic reals (`Sym`) to obtain straight-line SSA bytecode (the closure `df` of the instruction cache,
#                 src/derivatives/broadcast.jl:147-153, cannot cross a C ABI), runs of ScalarInstructions become OP_SCALAR_BLOCKs;
#   2. compile(t) hands the bytes to rdb_tape_load (include/rdb200.h);
#   3. gradient!/jacobian!/hessian! methods for the compiled tape `ccall` the replay entry points.
# Unsupported instruction kinds (`@grad` / ChainRules closures, det, inv, cat, non-whitelisted scalar functions) make `lower` throw:
# the engine has no CPU fallback.
module ReverseDiffB200

using ReverseDiff
using ReverseDiff: AbstractTape, GradientTape, JacobianTape, HessianTape, AbstractInstruction, SpecialInstruction,
                   ScalarInstruction, InstructionTape, TrackedArray, TrackedReal, ForwardOptimize, value, deriv, istracked,
                   hasorigin, input_hook, output_hook, func_hook, NULL_INDEX
import ReverseDiff.DiffResults
using ReverseDiff.DiffResults: DiffResult
using ReverseDiff.StaticArrays: SVector
using LinearAlgebra
using Statistics

const LIB = get(ENV, "RDB200_LIB", "librdb200.so")

# ------------------------------------------------------------------------------------------------ wire constants (program.py)
const MAGIC = 0x3030324244525052
const VERSION = UInt32(2)
const F64, F32 = UInt32(0), UInt32(1)
const BUF_TA, BUF_TR, BUF_CONST = Int32(0), Int32(1), Int32(2)
const CLS_TA, CLS_IDX, CLS_CONST, CLS_TR = Int32(0), Int32(1), Int32(2), Int32(3)
const IDX_NONE, IDX_EXPLICIT, IDX_TRANSPOSE = Int32(0), Int32(1), Int32(2)
const OP_MUL, OP_ADD, OP_SUB, OP_NEG, OP_SUM, OP_MEAN, OP_NBCAST, OP_MAP, OP_BCAST = Int32.((1, 2, 3, 4, 5, 6, 7, 8, 9))
const OP_BCAST_ADD, OP_BCAST_SUB, OP_BCAST_MUL, OP_BCAST_DIV, OP_BCAST_LDIV, OP_BCAST_POW = Int32.((10, 11, 12, 13, 14, 15))
const OP_GETINDEX, OP_DOT, OP_SUMDIMS, OP_TRACKER_BCAST, OP_GETINDEX_GENERIC, OP_SCALAR_BLOCK = Int32.((16, 17, 18, 19, 20, 21))
const SP_POOL, SP_LITERAL, SP_NONE = Int64(-1), Int64(-2), Int64(-3)
const E_CONST, E_ADD, E_SUB, E_MUL, E_DIV, E_NEG, E_POWI, E_POW = Int32.((1, 2, 3, 4, 5, 6, 7, 8))
const E_TANH, E_EXP, E_LOG, E_SQRT, E_SIN, E_COS, E_ABS2, E_ABS = Int32.((9, 10, 11, 12, 13, 14, 15, 16))
const E_MAX, E_MIN, E_INV, E_SIGMOID, E_ARG = Int32.((17, 18, 19, 20, 21))
const E_LOG1P, E_EXPM1, E_ASIN, E_ACOS, E_ATAN, E_SINH, E_COSH, E_TAN, E_ATAN2, E_HYPOT = Int32.((22, 23, 24, 25, 26, 27, 28, 29, 30, 31))
const E_EXP2, E_EXP10, E_LOG2, E_LOG10, E_CBRT, E_ASINH, E_ACOSH, E_ATANH = Int32.((32, 33, 34, 35, 36, 37, 38, 39))
const E_SEC, E_CSC, E_COT, E_SECH, E_CSCH, E_COTH = Int32.((40, 41, 42, 43, 44, 45))
const WIRE_ARGS, MAX_ARGS, MAX_DIMS, MAX_EXPR_REGS = 4, 8, 4, 32     # 4 operand slots per record; operands 5.. in OP_MORE_OPERANDS records
const OP_MORE_OPERANDS = Int32(22)
const COMPACT_TRANSPOSE_THRESHOLD = 1 << 20

# ------------------------------------------------------------------------------------------------ symbolic reals (expr.py)
"One closure trace: SSA ops `(code, a, b, imm)`, registers 0..nargs-1 are the arguments (expr.py `Tracer`)."
mutable struct Tracer
    nargs::Int
    dual::Vector{Bool}          # argument reaches the closure as a ForwardDiff Dual (only `^` cares, expr.py pow_imm)
    pow_flavor::Int
    ops::Vector{NTuple{4,Int32}}
    fconsts::Vector{Float64}
end
Tracer(nargs::Int, dual::Vector{Bool}, flavor::Int = 0) = Tracer(nargs, dual, flavor, NTuple{4,Int32}[], Float64[])

"Symbolic `Real` used to run an elementwise / scalar closure once at compile time (expr.py `Sym`)."
struct Sym <: Real
    tr::Tracer
    reg::Int32
    dual::Bool
end

function emit(tr::Tracer, code::Int32, a::Integer = 0, b::Integer = 0, imm::Integer = 0; dual::Bool = true)
    push!(tr.ops, (code, Int32(a), Int32(b), Int32(imm)))
    reg = tr.nargs + length(tr.ops) - 1
    reg >= MAX_EXPR_REGS && error("scalar closure too large (> $MAX_EXPR_REGS SSA values)")
    return Sym(tr, Int32(reg), dual)
end
function constant(tr::Tracer, c::Real)
    push!(tr.fconsts, Float64(c))
    return emit(tr, E_CONST, 0, 0, length(tr.fconsts) - 1; dual = false)
end
lift(tr::Tracer, x::Sym) = (x.tr === tr || error("symbolic value from a different closure trace"); x)
lift(tr::Tracer, x::Real) = constant(tr, x)

pow_imm(flavor::Int, a_dual::Bool, b_dual::Bool) = (flavor & 3) | (a_dual ? 0 : 4) | (b_dual ? 0 : 8)

function binop(code::Int32, a, b)
    tr = a isa Sym ? a.tr : b.tr
    sa, sb = lift(tr, a), lift(tr, b)              # a literal is materialised when the operation runs, left operand first
    imm = code == E_POW ? pow_imm(tr.pow_flavor, sa.dual, sb.dual) : 0
    return emit(tr, code, sa.reg, sb.reg, imm; dual = sa.dual || sb.dual)
end
unop(code::Int32, a::Sym) = emit(a.tr, code, a.reg; dual = a.dual)

# the same argument-type list the reference uses to stay clear of method ambiguities (src/ReverseDiff.jl REAL_TYPES)
const REALS = (:Bool, :Integer, :Rational, :AbstractFloat, :AbstractIrrational, :Real)
for (f, code) in ((:+, :E_ADD), (:-, :E_SUB), (:*, :E_MUL), (:/, :E_DIV), (:max, :E_MAX), (:min, :E_MIN))
    @eval Base.$f(a::Sym, b::Sym) = binop($code, a, b)
    for R in REALS
        @eval Base.$f(a::Sym, b::$R) = binop($code, a, b)
        @eval Base.$f(a::$R, b::Sym) = binop($code, a, b)
    end
end
Base.:\(a::Sym, b::Sym) = binop(E_DIV, b, a)
Base.:^(a::Sym, b::Sym) = binop(E_POW, a, b)
Base.:^(a::Sym, p::Integer) = emit(a.tr, E_POWI, a.reg, 0, p; dual = a.dual)      # Base.literal_pow: repeated product
Base.literal_pow(::typeof(^), a::Sym, ::Val{p}) where {p} = emit(a.tr, E_POWI, a.reg, 0, p; dual = a.dual)
for R in (:Rational, :AbstractFloat, :AbstractIrrational, :Real)
    @eval Base.:^(a::Sym, b::$R) = binop(E_POW, a, b)
end
for R in REALS
    @eval Base.:^(a::$R, b::Sym) = binop(E_POW, a, b)
end
Base.:-(a::Sym) = unop(E_NEG, a)
Base.:+(a::Sym) = a
for (f, code) in ((:tanh, :E_TANH), (:exp, :E_EXP), (:log, :E_LOG), (:sqrt, :E_SQRT), (:sin, :E_SIN), (:cos, :E_COS),
                  (:abs2, :E_ABS2), (:abs, :E_ABS), (:inv, :E_INV),
                  (:log1p, :E_LOG1P), (:expm1, :E_EXPM1), (:asin, :E_ASIN), (:acos, :E_ACOS), (:atan, :E_ATAN), (:sinh, :E_SINH),
                  (:cosh, :E_COSH), (:tan, :E_TAN), (:exp2, :E_EXP2), (:exp10, :E_EXP10), (:log2, :E_LOG2), (:log10, :E_LOG10),
                  (:cbrt, :E_CBRT), (:asinh, :E_ASINH), (:acosh, :E_ACOSH), (:atanh, :E_ATANH), (:sec, :E_SEC), (:csc, :E_CSC),
                  (:cot, :E_COT), (:sech, :E_SECH), (:csch, :E_CSCH), (:coth, :E_COTH))
    @eval Base.$f(a::Sym) = unop($code, a)
end
for (f, code) in ((:atan, :E_ATAN2), (:hypot, :E_HYPOT))            # two-argument atan(y, x) and hypot
    @eval Base.$f(a::Sym, b::Sym) = binop($code, a, b)
    for R in REALS
        @eval Base.$f(a::Sym, b::$R) = binop($code, a, b)
        @eval Base.$f(a::$R, b::Sym) = binop($code, a, b)
    end
end
if isdefined(ReverseDiff, :LogExpFunctions)
    @eval ReverseDiff.LogExpFunctions.logistic(a::Sym) = unop(E_SIGMOID, a)
end
Base.zero(a::Sym) = constant(a.tr, 0.0)
Base.one(a::Sym) = constant(a.tr, 1.0)
Base.float(a::Sym) = a
Base.convert(::Type{Sym}, a::Sym) = a
nobranch(args...) = error("data-dependent control flow cannot be recorded on a compiled tape (reference docs/src/api.md:39-44)")
for f in (:<, :<=, :(==), :isless, :isequal)
    @eval Base.$f(a::Sym, b::Sym) = nobranch()
    @eval Base.$f(a::Sym, b::Real) = nobranch()
    @eval Base.$f(a::Real, b::Sym) = nobranch()
end
for f in (:iszero, :isone, :isnan, :isinf, :isfinite, :signbit)
    @eval Base.$f(a::Sym) = nobranch()
end

"""
    trace(f, nargs; dual, pow_flavor, literal_args) -> (ops, fconsts)

Run `f` on `nargs` symbolic reals (expr.py `trace`).  The result is the last SSA value.
"""
function trace(call, nargs::Int; dual::Vector{Bool} = fill(true, nargs), pow_flavor::Int = 0)
    nargs > MAX_ARGS && error("elementwise closures with more than $MAX_ARGS array arguments are not supported")
    tr = Tracer(nargs, dual, pow_flavor)
    out = call(ntuple(i -> Sym(tr, Int32(i - 1), dual[i]), nargs)...)
    out = lift(tr, out)
    if out.reg < nargs                                  # the closure returned one of its arguments
        out = emit(tr, E_ARG, 0, 0, out.reg; dual = out.dual)
    elseif out.reg != nargs + length(tr.ops) - 1        # result must be the last SSA value
        z = constant(tr, 0.0)
        out = emit(tr, E_ADD, out.reg, z.reg; dual = out.dual)
    end
    return (copy(tr.ops), copy(tr.fconsts))
end

# ------------------------------------------------------------------------------------------------ tape program (program.py)
struct Operand
    buffer::Int32
    cls::Int32
    tracked::Bool
    dims::Vector{Int64}
    idx_kind::Int32
    idx_map::Union{Nothing,Vector{Int64}}
    bound::Union{Nothing,Vector{Int64}}
end
Operand(buffer, cls, tracked, dims; idx_kind = IDX_NONE, idx_map = nothing, bound = nothing) =
    Operand(Int32(buffer), cls, tracked, collect(Int64, dims), idx_kind, idx_map, bound)

const Expr4 = Tuple{Vector{NTuple{4,Int32}},Vector{Float64}}

"A maximal run of consecutive ScalarInstructions, shipped as ONE tape entry (program.py `ScalarBlock`)."
mutable struct ScalarBlock
    token_buffer::Int32
    spill_buffer::Int32
    n_spill::Int
    expr_id::Vector{Int64}
    a_space::Vector{Int64}; a_idx::Vector{Int64}
    b_space::Vector{Int64}; b_idx::Vector{Int64}
    exprs::Vector{Tuple{Vector{NTuple{4,Int32}},Vector{Float64},Int}}
    expr_key::Dict{Any,Int}
    refbufs::Vector{Int64}
    ref_ord::Dict{Int,Int}
    literals::Vector{Float64}
    lit_ord::Dict{UInt64,Int}
    exports::Vector{NTuple{3,Int64}}            # (slot, ref ordinal, element)
    export_of::Dict{Int,Tuple{Int,Int}}         # slot -> (buffer, element) of its first export
end
ScalarBlock(token) = ScalarBlock(Int32(token), Int32(-1), 0, Int64[], Int64[], Int64[], Int64[], Int64[],
                                 Tuple{Vector{NTuple{4,Int32}},Vector{Float64},Int}[], Dict{Any,Int}(), Int64[], Dict{Int,Int}(),
                                 Float64[], Dict{UInt64,Int}(), NTuple{3,Int64}[], Dict{Int,Tuple{Int,Int}}())
Base.length(b::ScalarBlock) = length(b.expr_id)
function expr!(b::ScalarBlock, ops, fconsts, nargs::Int)
    key = (copy(ops), reinterpret(UInt64, fconsts) |> collect, nargs)
    return get!(b.expr_key, key) do
        push!(b.exprs, (ops, fconsts, nargs)); length(b.exprs) - 1
    end
end
ref!(b::ScalarBlock, buffer::Integer) = get!(b.ref_ord, Int(buffer)) do
    push!(b.refbufs, buffer); length(b.refbufs) - 1
end
literal!(b::ScalarBlock, v::Real) = get!(b.lit_ord, reinterpret(UInt64, Float64(v))) do      # bit pattern keeps -0.0 / NaN payloads apart
    push!(b.literals, Float64(v)); length(b.literals) - 1
end
function export!(b::ScalarBlock, slot::Integer, buffer::Integer, elem::Integer)
    push!(b.exports, (Int64(slot), Int64(ref!(b, buffer)), Int64(elem)))
    get!(b.export_of, Int(slot), (Int(buffer), Int(elem)))
    return nothing
end
function push_scalar!(b::ScalarBlock, eid::Int, a::Tuple{Int64,Int64}, bb::Tuple{Int64,Int64})
    push!(b.expr_id, eid); push!(b.a_space, a[1]); push!(b.a_idx, a[2]); push!(b.b_space, bb[1]); push!(b.b_idx, bb[2])
    return length(b.expr_id) - 1
end

mutable struct Instr
    opcode::Int32
    inputs::Vector{Operand}
    out::Int32
    expr::Union{Nothing,Expr4}
    n_expr_args::Int32
    scalar::Union{Nothing,ScalarBlock}
end
mutable struct Buffer
    kind::Int32
    dims::Vector{Int64}
    alias_of::Int32
    const_data::Union{Nothing,Array}
end
mutable struct TapeProgram
    dtype::UInt32
    buffers::Vector{Buffer}
    inputs::Vector{Int32}
    instrs::Vector{Instr}
    output_buffer::Int32
end
TapeProgram(dtype) = TapeProgram(dtype, Buffer[], Int32[], Instr[], Int32(-1))
function add_buffer!(p::TapeProgram, kind::Int32, dims; alias_of = -1, const_data = nothing)
    length(dims) > MAX_DIMS && error("arrays of more than $MAX_DIMS dimensions are not supported")
    push!(p.buffers, Buffer(kind, collect(Int64, dims), Int32(alias_of), const_data))
    return Int32(length(p.buffers) - 1)
end

align8(n::Integer) = (n + 7) & ~7
padded(dims::Vector{Int64}, fill::Int64) = Int64[i <= length(dims) ? dims[i] : fill for i in 1:MAX_DIMS]

"Serialise exactly as `TapeProgram.to_bytes` does (program.py:313-441): same section order, same padding, same chunk order."
function to_bytes(p::TapeProgram)
    T = p.dtype == F64 ? Float64 : Float32
    for it in p.instrs                                           # finalize(): a block's spill buffer is sized when recording is over
        if it.scalar !== nothing && it.scalar.spill_buffer >= 0
            p.buffers[it.scalar.spill_buffer + 1].dims = Int64[max(1, it.scalar.n_spill)]
        end
    end
    bufs = IOBuffer(); consts = IOBuffer()
    for b in p.buffers
        write(bufs, b.kind, Int32(length(b.dims)))
        foreach(d -> write(bufs, d), padded(b.dims, Int64(1)))
        off = Int64(-1)
        if b.kind == BUF_CONST
            off = Int64(position(consts))
            write(consts, vec(convert(Array{T}, b.const_data)))          # column-major
            write(consts, zeros(UInt8, align8(position(consts)) - position(consts)))
        end
        write(bufs, b.alias_of, Int32(0), off)
    end
    ins = IOBuffer(); idx = IOBuffer(); ex = IOBuffer(); fc = IOBuffer()
    n_records = 0
    n_idx() = position(idx) ÷ 8; n_ex() = position(ex) ÷ 4; n_fc() = position(fc) ÷ 8
    function put_expr(ops, fconsts)                                      # literal indices become absolute in the program-wide pool
        base = n_fc(); start = n_ex()
        for (code, a, b, imm) in ops
            write(ex, code, a, b, code == E_CONST ? Int32(imm + base) : imm)
        end
        foreach(c -> write(fc, c), fconsts)
        return start, length(ops)
    end
    for it in p.instrs
        inputs = it.inputs
        if it.scalar !== nothing
            blk = it.scalar
            table = Int64[]
            for (ops, fconsts, na) in blk.exprs
                start, n = put_expr(ops, fconsts)
                append!(table, (start, n, na))
            end
            lit_base = n_fc()
            foreach(c -> write(fc, c), blk.literals)
            n = length(blk)
            payload = Int64[n, length(blk.refbufs), length(blk.exports), length(blk.exprs)]
            append!(payload, blk.refbufs); append!(payload, table)
            for e in blk.exports; append!(payload, e); end
            for i in 1:n
                ai = blk.a_space[i] == SP_LITERAL ? blk.a_idx[i] + lit_base : blk.a_idx[i]
                bi = blk.b_space[i] == SP_LITERAL ? blk.b_idx[i] + lit_base : blk.b_idx[i]
                append!(payload, (blk.expr_id[i], blk.a_space[i], ai, blk.b_space[i], bi))
            end
            inputs = [Operand(it.out, CLS_TR, false, Int64[]; idx_kind = IDX_EXPLICIT, idx_map = payload)]
        end
        length(inputs) > MAX_ARGS && error("instructions with more than $MAX_ARGS inputs are not supported")
        eoff, elen = Int64(0), Int64(0)
        if it.expr !== nothing
            eoff, elen = put_expr(it.expr[1], it.expr[2])
        end
        write(ins, it.opcode, Int32(length(inputs)), it.out, it.n_expr_args, Int64(eoff), Int64(elen))
        nrec = 1 + (max(length(inputs), 1) - 1) ÷ WIRE_ARGS
        n_records += nrec
        for j in 1:(nrec * WIRE_ARGS)
            if j > WIRE_ARGS && (j - 1) % WIRE_ARGS == 0             # continuation record: operands j .. j+3 of this instruction
                write(ins, OP_MORE_OPERANDS, Int32(min(WIRE_ARGS, length(inputs) - (j - 1))), Int32(-1), Int32(0), Int64(0), Int64(0))
            end
            if j > length(inputs)
                write(ins, zeros(UInt8, 104)); continue
            end
            o = inputs[j]
            ioff, ilen = Int64(-1), Int64(0)
            if o.idx_kind == IDX_EXPLICIT
                ioff, ilen = Int64(n_idx()), Int64(length(o.idx_map))
                write(idx, o.idx_map)
            end
            write(ins, o.buffer, o.cls, Int32(o.tracked), o.idx_kind, ioff, ilen, Int32(length(o.dims)), Int32(0))
            foreach(d -> write(ins, d), padded(o.dims, Int64(1)))
            foreach(d -> write(ins, d), o.bound === nothing ? zeros(Int64, MAX_DIMS) : padded(o.bound, Int64(1)))
        end
    end
    sections = (take!(bufs), reinterpret(UInt8, p.inputs) |> collect, take!(ins), take!(idx), take!(ex), take!(fc), take!(consts))
    counts = (length(p.buffers), length(p.inputs), n_records)
    offs = Int[]; off = 128
    for s in sections
        push!(offs, off); off += align8(length(s))
    end
    out = IOBuffer()
    write(out, UInt64(MAGIC), VERSION, p.dtype, UInt32(counts[1]), UInt32(counts[3]), UInt32(counts[2]), p.output_buffer)
    write(out, UInt64(offs[1]), UInt64(offs[2]), UInt64(offs[3]))
    write(out, UInt64(offs[4]), UInt64(length(sections[4]) ÷ 8))
    write(out, UInt64(offs[5]), UInt64(length(sections[5]) ÷ 4))
    write(out, UInt64(offs[6]), UInt64(length(sections[6]) ÷ 8))
    write(out, UInt64(offs[7]), UInt64(length(sections[7])))
    write(out, UInt64(off))
    @assert position(out) == 128
    for s in sections
        write(out, s); write(out, zeros(UInt8, align8(length(s)) - length(s)))
    end
    return take!(out)
end

"`transpose_map` (program.py:505-515): index map of `adjoint(A)` for a rows x cols column-major parent.  Element (i, j) of the
cols x rows adjoint sits at k = i + j*cols and reads parent[j + i*rows] (all 0-based)."
function transpose_map(rows::Int, cols::Int)
    m = Vector{Int64}(undef, rows * cols)
    for j in 0:rows-1, i in 0:cols-1
        m[i + j * cols + 1] = j + i * rows
    end
    return m
end

# ------------------------------------------------------------------------------------------------ lowering
mutable struct Lowering
    prog::TapeProgram
    T::DataType
    arrays::IdDict{Any,Int32}                      # value array (identity) -> buffer id
    storage::Dict{Tuple{UInt,Int},Int32}           # (data pointer, length) -> root buffer: `reshape` aliases (src/tracked.jl:402-408)
    reals::IdDict{Any,Any}                         # TrackedReal -> Int32 buffer id | (ScalarBlock, slot)
    block::Union{Nothing,ScalarBlock}              # the block the next ScalarInstruction is appended to
end

root_buffer(p::TapeProgram, b::Integer) = (while p.buffers[b + 1].alias_of >= 0; b = p.buffers[b + 1].alias_of; end; Int32(b))

function register_array!(L::Lowering, t::TrackedArray)
    v = value(t)
    b = add_buffer!(L.prog, BUF_TA, size(v))
    L.arrays[v] = b
    v isa Array && (L.storage[(UInt(pointer(v)), length(v))] = b)
    return b
end
function buffer_of(L::Lowering, t::TrackedArray)
    v = value(t)
    haskey(L.arrays, v) && return L.arrays[v]
    base = v isa Base.ReshapedArray ? parent(v) : v
    key = base isa Array ? (UInt(pointer(base)), length(base)) : nothing
    if key !== nothing && haskey(L.storage, key)     # same storage, other shape: a reshape alias shares value AND deriv
        b = add_buffer!(L.prog, BUF_TA, size(v); alias_of = root_buffer(L.prog, L.storage[key]))
        L.arrays[v] = b
        return b
    end
    error("TrackedArray that is neither a tape input nor the output of a recorded instruction")
end

"The block consecutive ScalarInstructions share (tracked.py `InstructionTape.scalar_block`)."
function scalar_block!(L::Lowering)
    L.block !== nothing && return L.block
    token = add_buffer!(L.prog, BUF_TR, ())
    blk = ScalarBlock(token)
    push!(L.prog.instrs, Instr(OP_SCALAR_BLOCK, Operand[], token, nothing, Int32(0), blk))
    L.block = blk
    return blk
end

"`(buffer, elem)` through which readers outside the block see a slot; exports it on first use (tracked.py `_slot_buffer`)."
function slot_buffer!(L::Lowering, blk::ScalarBlock, slot::Int; want_whole::Bool = false)
    hit = get(blk.export_of, slot, nothing)
    if hit !== nothing && (!want_whole || L.prog.buffers[hit[1] + 1].kind == BUF_TR)
        return hit
    end
    if want_whole
        b = add_buffer!(L.prog, BUF_TR, ())
        export!(blk, slot, b, 0)
        hit !== nothing && (blk.export_of[slot] = hit)      # the first export stays the canonical one for scalar readers
        return (Int(b), 0)
    end
    if blk.spill_buffer < 0
        blk.spill_buffer = add_buffer!(L.prog, BUF_TA, (1,))
    end
    k = blk.n_spill
    blk.n_spill += 1
    export!(blk, slot, blk.spill_buffer, k)
    return (Int(blk.spill_buffer), k)
end

"Operand of a scalar instruction: (space, index) (tracked.py `_scalar_operand`)."
function scalar_operand!(L::Lowering, blk::ScalarBlock, x)
    if x isa TrackedReal
        where_ = get(L.reals, x, nothing)
        if where_ isa Tuple
            b, slot = where_
            b === blk && return (SP_POOL, Int64(slot))
            buf, k = slot_buffer!(L, b, slot)
            return (Int64(ref!(blk, buf)), Int64(k))
        elseif where_ isa Int32                            # output of an array instruction (sum, dot, ...): size-1 buffer
            return (Int64(ref!(blk, where_)), Int64(0))
        elseif hasorigin(x)                                # scalar getindex records nothing (src/tracked.jl:348-351)
            return (Int64(ref!(blk, root_buffer(L.prog, buffer_of(L, x.origin)))), Int64(x.index - 1))
        end
        error("TrackedReal without a producer on this tape")
    end
    x isa Real || error("scalar instructions take TrackedReal / Real arguments")
    return (SP_LITERAL, Int64(literal!(blk, x)))           # capture(::Real) (src/tape.jl:24): frozen by value
end

function record_scalar!(L::Lowering, ex::Expr4, out::TrackedReal, a, b = nothing)
    blk = scalar_block!(L)
    eid = expr!(blk, ex[1], ex[2], b === nothing ? 1 : 2)
    oa = scalar_operand!(L, blk, a)
    ob = b === nothing ? (SP_NONE, Int64(0)) : scalar_operand!(L, blk, b)
    slot = push_scalar!(blk, eid, oa, ob)
    L.reals[out] = (blk, slot)
    return slot
end
identity_expr() = (NTuple{4,Int32}[(E_ARG, Int32(0), Int32(0), Int32(0))], Float64[])

"Size-1 buffer that carries a TrackedReal (operand class TR of array instructions; output of a GradientTape)."
function scalar_buffer!(L::Lowering, t::TrackedReal)
    where_ = get(L.reals, t, nothing)
    where_ isa Int32 && return where_
    if where_ === nothing                                  # element of an array: route it through a slot of its own
        hasorigin(t) || error("TrackedReal without a producer on this tape")
        ghost = TrackedReal(value(t), zero(value(t)))      # stands for the identity instruction's output in the slot table
        record_scalar!(L, identity_expr(), ghost, t)
        where_ = L.reals[ghost]
    end
    blk, slot = where_
    return Int32(slot_buffer!(L, blk, slot; want_whole = true)[1])
end

"An Array{TrackedReal} used as tape OUTPUT: element k of a fresh buffer mirrors xs[k] (tracked.py `collect`)."
function collect_output!(L::Lowering, xs::AbstractArray)
    buf = add_buffer!(L.prog, BUF_TA, (length(xs),))
    for (k, x) in enumerate(xs)
        blk, slot = if x isa TrackedReal && get(L.reals, x, nothing) isa Tuple
            L.reals[x]
        elseif x isa TrackedReal
            ghost = TrackedReal(value(x), zero(value(x)))
            record_scalar!(L, identity_expr(), ghost, x)
            L.reals[ghost]
        else                                               # a plain Real in the output: a literal slot
            b = scalar_block!(L)
            eid = expr!(b, NTuple{4,Int32}[(E_CONST, Int32(0), Int32(0), Int32(0))], Float64[Float64(x)], 0)
            (b, push_scalar!(b, eid, (SP_NONE, Int64(0)), (SP_NONE, Int64(0))))
        end
        export!(blk, slot, buf, k - 1)
    end
    return buf
end

bound_of(dims, out_dims) = Int64[i <= length(dims) ? dims[i] : 1 for i in 1:length(out_dims)]   # index_bound, propagation.jl:25-27

"`capture`d instruction input -> operand (tracked.py `_capture`; src/tracked.jl:223-225)."
function operand!(L::Lowering, x; bound_to = nothing)
    bnd(dims) = bound_to === nothing ? nothing : bound_of(collect(Int64, dims), bound_to)
    if x isa TrackedArray
        return Operand(buffer_of(L, x), CLS_TA, true, size(x); bound = bnd(size(x)))
    elseif x isa TrackedReal
        return Operand(scalar_buffer!(L, x), CLS_TR, true, (); bound = bnd(()))
    elseif x isa AbstractArray && eltype(x) <: TrackedReal
        # index-linked dense copy of e.g. adjoint(a): every element must be a scalar getindex of ONE TrackedArray
        isempty(x) && error("empty tracked operand")
        origin = first(x).origin
        all(e -> hasorigin(e) && e.origin === origin, x) || error("Array{TrackedReal} operand whose elements do not share one origin array")
        b = buffer_of(L, origin)
        m = Int64[e.index - 1 for e in vec(x)]
        if ndims(origin) == 2 && length(m) > COMPACT_TRANSPOSE_THRESHOLD && size(x) == reverse(size(origin)) &&
           m == transpose_map(size(origin, 1), size(origin, 2))
            return Operand(b, CLS_IDX, true, size(x); idx_kind = IDX_TRANSPOSE, bound = bnd(size(x)))
        end
        return Operand(b, CLS_IDX, true, size(x); idx_kind = IDX_EXPLICIT, idx_map = m, bound = bnd(size(x)))
    elseif x isa AbstractArray{<:Real}
        b = add_buffer!(L.prog, BUF_CONST, size(x); const_data = convert(Array{L.T}, x))
        return Operand(b, CLS_CONST, false, size(x); bound = bnd(size(x)))
    end
    error("operand of type $(typeof(x)) cannot be lowered")
end

function new_output!(L::Lowering, out)
    if out isa TrackedArray
        return register_array!(L, out)
    elseif out isa TrackedReal
        b = add_buffer!(L.prog, BUF_TR, ())
        L.reals[out] = b
        return b
    end
    error("instruction output of type $(typeof(out)) cannot be lowered")
end

function push_instr!(L::Lowering, opcode::Int32, inputs::Vector{Operand}, out::Int32; expr = nothing, nargs = 0)
    L.block = nothing                                      # an array instruction ends the current scalar block (tape order is kept)
    push!(L.prog.instrs, Instr(opcode, inputs, out, expr, Int32(nargs), nothing))
end

closure_field(df, name::Symbol) = getfield(df, name)
"The user closure behind the `df` of a map/broadcast cache (src/derivatives/elementwise.jl:230-310): `df` captured `f::ForwardOptimize`."
function forward_closure(df)
    for name in fieldnames(typeof(df))
        v = getfield(df, name)
        v isa Core.Box && (v = v.contents)
        v isa ForwardOptimize && return v.f
    end
    for name in fieldnames(typeof(df))                      # nested `let` closures (the TrackedReal x Array methods)
        v = getfield(df, name)
        v isa Function && return forward_closure(v)
    end
    error("cannot find the @forward closure inside the instruction cache")
end

const INFIX = Dict{Any,Tuple{Int32,Function}}(
    (+) => (OP_BCAST_ADD, (a, b) -> a + b), (-) => (OP_BCAST_SUB, (a, b) -> a - b), (*) => (OP_BCAST_MUL, (a, b) -> a * b),
    (/) => (OP_BCAST_DIV, (a, b) -> a / b), (\) => (OP_BCAST_LDIV, (a, b) -> b / a), (^) => (OP_BCAST_POW, (a, b) -> a^b))

function lower_elementwise!(L::Lowering, opcode::Int32, call, args::Tuple, out; dual, flavor = 0)
    ex = trace(call, length(args); dual = collect(Bool, dual), pow_flavor = flavor)
    ob = new_output!(L, out)
    odims = collect(Int64, size(out))
    ops = Operand[operand!(L, a; bound_to = odims) for a in args]
    push_instr!(L, opcode, ops, ob; expr = ex, nargs = length(args))
end

"0-based linear index map of a `getindex` iterable (src/tracked.jl:291-316,342-362)."
linear_map(t::TrackedArray, idx) = Int64[LinearIndices(size(t))[CartesianIndex(i)] - 1 for i in idx][:]

function lower_special!(L::Lowering, ins::SpecialInstruction)
    f, input, out, cache = ins.func, ins.input, ins.output, ins.cache
    if f === (*)
        ob = new_output!(L, out)
        push_instr!(L, OP_MUL, Operand[operand!(L, input[1]), operand!(L, input[2])], ob)
    elseif f === (+) || (f === (-) && input isa Tuple)
        ob = new_output!(L, out)
        push_instr!(L, f === (+) ? OP_ADD : OP_SUB, Operand[operand!(L, input[1]), operand!(L, input[2])], ob)
    elseif f === (-)
        ob = new_output!(L, out)
        push_instr!(L, OP_NEG, Operand[operand!(L, input)], ob)
    elseif f === sum || f === Statistics.mean
        ob = new_output!(L, out)
        push_instr!(L, f === sum ? OP_SUM : OP_MEAN, Operand[operand!(L, input)], ob)
    elseif f === LinearAlgebra.dot
        ob = new_output!(L, out)
        push_instr!(L, OP_DOT, Operand[operand!(L, input[1]), operand!(L, input[2])], ob)
    elseif f === sum!
        ob = haskey(L.arrays, value(out)) ? buffer_of(L, out) : new_output!(L, out)
        x = input::TrackedArray
        push_instr!(L, OP_SUMDIMS, Operand[Operand(buffer_of(L, x), CLS_TA, true, size(x); bound = collect(Int64, size(out)))], ob)
    elseif f === ReverseDiff.∇broadcast
        # cache = (results, df, bounds); df captured (result, f, untracked, inds)  (src/derivatives/broadcast.jl:142-160)
        df = cache[2]
        g, untracked, inds = closure_field(df, :f), closure_field(df, :untracked), closure_field(df, :inds)
        N = length(input)
        call = (s...) -> ReverseDiff.splatcall(g, SVector{N,Sym}(s...), untracked, inds)
        lower_elementwise!(L, OP_NBCAST, call, Tuple(input), out; dual = fill(true, N))
    elseif f === ReverseDiff.tracker_∇broadcast
        # plain Reals are closed over as literals, TrackedReals stay operands of class TR (tracked.py `_record_tracker_bcast`)
        g = cache[1]
        live = Any[]; slots = Any[]
        for a in input
            if a isa TrackedArray || a isa TrackedReal || a isa AbstractArray
                push!(live, a); push!(slots, length(live))
            else
                push!(slots, Float64(a))
            end
        end
        call = (s...) -> g((sl isa Int ? s[sl] : sl for sl in slots)...)
        lower_elementwise!(L, OP_TRACKER_BCAST, call, Tuple(live), out; dual = fill(true, length(live)), flavor = 1)
    elseif f === map || f === broadcast
        args = input isa Tuple ? input : (input,)
        g = forward_closure(cache[2])
        lower_elementwise!(L, f === map ? OP_MAP : OP_BCAST, g, args, out; dual = [istracked(a) for a in args])
    elseif f isa Tuple && length(f) == 2 && f[1] === broadcast && haskey(INFIX, f[2])
        opcode, g = INFIX[f[2]]
        lower_elementwise!(L, opcode, g, Tuple(input), out; dual = [true, true], flavor = f[2] === (^) ? 2 : 0)
    elseif f === getindex
        t, idx = input
        m = idx isa AbstractArray{<:CartesianIndex} ? Int64[LinearIndices(size(t))[i] - 1 for i in vec(idx)] : linear_map(t, idx)
        ob = new_output!(L, out)
        push_instr!(L, OP_GETINDEX, Operand[Operand(buffer_of(L, t), CLS_TA, true, size(t); idx_kind = IDX_EXPLICIT, idx_map = m)], ob)
    elseif f isa Tuple && length(f) == 2 && f[1] === getindex
        t, inds = input
        cinds = CartesianIndices(map(i -> 1:length(i), inds))
        m = Int64[LinearIndices(size(t))[CartesianIndex(map(getindex, inds, Tuple(c)))] - 1 for c in cinds][:]
        ob = new_output!(L, out)
        push_instr!(L, OP_GETINDEX_GENERIC, Operand[Operand(buffer_of(L, t), CLS_TA, true, size(t); idx_kind = IDX_EXPLICIT, idx_map = m)], ob)
    elseif f === convert
        # convert(::Type{<:TrackedReal}, t::TrackedReal) (src/tracked.jl:237-257): value copy forward, adjoint pass-through in reverse
        # = the identity scalar instruction
        record_scalar!(L, identity_expr(), out, input)
    else
        error("instruction `$(f)` cannot be replayed on the device (no CPU fallback): @grad / ChainRules closures, det, inv, cat " *
              "and non-whitelisted functions are outside the tape-replay path")
    end
    return nothing
end

function lower_scalar!(L::Lowering, ins::ScalarInstruction)
    f, input, out = ins.func, ins.input, ins.output
    if input isa Tuple
        a, b = input
        if f === (^) && b isa Integer && a isa TrackedReal      # x^2 -> literal_pow: repeated product (tracked.py TrackedReal.__pow__)
            return record_scalar!(L, (NTuple{4,Int32}[(E_POWI, Int32(0), Int32(0), Int32(b))], Float64[]), out, a)
        end
        ex = trace(f, 2; dual = Bool[a isa TrackedReal, b isa TrackedReal])
        return record_scalar!(L, ex, out, a, b)
    end
    return record_scalar!(L, trace(f, 1; dual = Bool[true]), out, input)
end

"""
    lower(t::AbstractTape) -> Vector{UInt8}

The flat tape program of `t` (layout: reversediff.jl_b200/program.py; parsed by csrc/program.h and oracle/replay.py).
"""
function lower(t::AbstractTape)
    xs = input_hook(t) isa Tuple ? input_hook(t) : (input_hook(t),)
    T = eltype(value(first(xs)))
    T === Float64 || T === Float32 || error("only Float64 and Float32 tapes are supported")
    L = Lowering(TapeProgram(T === Float64 ? F64 : F32), T, IdDict{Any,Int32}(), Dict{Tuple{UInt,Int},Int32}(), IdDict{Any,Any}(), nothing)
    for x in xs
        x isa TrackedArray || error("tape inputs must be arrays")
        push!(L.prog.inputs, register_array!(L, x))
    end
    for ins in t.tape
        ins isa ScalarInstruction ? lower_scalar!(L, ins) : lower_special!(L, ins::SpecialInstruction)
    end
    out = output_hook(t)
    L.prog.output_buffer = out isa TrackedReal ? scalar_buffer!(L, out) :
                           out isa TrackedArray ? buffer_of(L, out) :
                           out isa AbstractArray ? collect_output!(L, out) :
                           error("tape output of type $(typeof(out)) cannot be lowered")
    return to_bytes(L.prog)
end

# ------------------------------------------------------------------------------------------------ C ABI (include/rdb200.h)
struct RdbOptions
    fuse_elementwise::Int32
    seed_block::Int32
    use_graph::Int32
    profile::Int32
    gemm_f64_mode::Int32
    ozaki_slices::Int32
    reserved::NTuple{2,Int32}
end
RdbOptions(; fuse = 1, seed_block = 0, graph = 1, profile = 0, gemm_f64_mode = 0, ozaki_slices = 0) =
    RdbOptions(fuse, seed_block, graph, profile, gemm_f64_mode, ozaki_slices, (0, 0))

last_error() = unsafe_string(ccall((:rdb_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("librdb200: ", last_error())

"`CompiledTape` (src/api/tape.jl:55-59) whose executors live on the device."
mutable struct B200CompiledTape{T<:AbstractTape} <: AbstractTape
    tape::T
    ctx::Ptr{Cvoid}
    handle::Ptr{Cvoid}
    function B200CompiledTape(t::T, program::Vector{UInt8}; device = 0, opts = RdbOptions()) where {T<:AbstractTape}
        ctx = ccall((:rdb_ctx_create, LIB), Ptr{Cvoid}, (Cint,), device)
        ctx == C_NULL && error("librdb200: ", last_error())          # no GPU => error, never a CPU fallback
        h = GC.@preserve program ccall((:rdb_tape_load, LIB), Ptr{Cvoid}, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t, Ref{RdbOptions}),
                                       ctx, program, length(program), opts)
        if h == C_NULL
            msg = last_error()
            ccall((:rdb_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx)
            error("librdb200: ", msg)
        end
        ct = new{T}(t, ctx, h)
        finalizer(ct) do c
            ccall((:rdb_tape_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.handle)
            ccall((:rdb_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.ctx)
        end
        return ct
    end
end

ReverseDiff.func_hook(ct::B200CompiledTape) = func_hook(ct.tape)
ReverseDiff.input_hook(ct::B200CompiledTape) = input_hook(ct.tape)
ReverseDiff.output_hook(ct::B200CompiledTape) = output_hook(ct.tape)
Base.length(ct::B200CompiledTape) = length(ct.tape)

"`ReverseDiff.compile(t)` (src/api/tape.jl:146-149) for the B200 engine."
compile(t::AbstractTape; kw...) = B200CompiledTape(t, lower(t); kw...)

inputs_tuple(input) = input isa Tuple ? input : (input,)

# value!(input_hook(t), input) (src/api/tape.jl:40-44).  rdb_tape_set_input only ENQUEUES the copy (include/rdb200.h): the caller
# keeps `xs` alive (GC.@preserve) until the replay call that follows has returned - see the API methods below.
function set_inputs!(ct::B200CompiledTape, xs::Tuple)
    length(xs) == length(inputs_tuple(input_hook(ct))) || throw(ArgumentError("wrong number of inputs"))
    for (i, x) in enumerate(xs)
        x isa Array || throw(ArgumentError("inputs must be dense column-major Arrays (IndexLinear, src/tracked.jl:77)"))
        size(x) == size(inputs_tuple(input_hook(ct))[i]) || throw(DimensionMismatch("input $i has a different shape than the recorded one"))
        check(ccall((:rdb_tape_set_input, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint), ct.handle, i - 1, x, 0))
    end
end

gradient_array(r::DiffResult) = DiffResults.gradient(r)
gradient_array(r::AbstractArray) = r

# gradient!(result, tape::CompiledGradient, input)            src/api/gradients.jl:78-82
function ReverseDiff.gradient!(result, ct::B200CompiledTape{<:GradientTape}, input)
    xs = inputs_tuple(input)
    rs = result isa Tuple ? result : (result,)
    arrs = map(gradient_array, rs)
    ptrs = Ptr{Cvoid}[pointer(a) for a in arrs]
    val = Ref{Float64}(0.0)                                   # (an F32 tape writes 4 bytes; read back through the tape's eltype)
    GC.@preserve xs arrs ptrs begin                           # inputs AND results stay rooted across set_input -> rdb_gradient
        set_inputs!(ct, xs)
        check(ccall((:rdb_gradient, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cvoid}), ct.handle, ptrs, 0, val))
    end
    T = eltype(first(arrs))
    v = T === Float32 ? reinterpret(Float32, [val[]])[1] : val[]
    out = map(r -> r isa DiffResult ? DiffResults.value!(r, v) : r, rs)      # extract_result!, src/api/utils.jl:96-100
    return result isa Tuple ? out : out[1]
end

# jacobian!(result, tape::CompiledJacobian, input)            src/api/jacobians.jl:120-124
# rows = (begin, end) 0-based half-open shard of output-seed rows (multi-GPU, SURVEY.md section 8e); the result then has
# size(result, 1) = end - begin rows.
function ReverseDiff.jacobian!(result, ct::B200CompiledTape{<:JacobianTape}, input; rows = nothing, slot = nothing)
    xs = inputs_tuple(input)
    rs = result isa Tuple ? result : (result,)
    m = length(output_hook(ct))
    rb, re = rows === nothing ? (0, m) : rows
    GC.@preserve xs rs begin
        set_inputs!(ct, xs)
        for (i, r) in enumerate(rs)                           # tuple inputs: one result per input (src/api/utils.jl:66-71)
            slot !== nothing && slot != i - 1 && continue
            J = r isa DiffResult ? DiffResults.jacobian(r) : r
            yv = r isa DiffResult ? DiffResults.value(r) : nothing
            check(ccall((:rdb_jacobian, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}),
                        ct.handle, i - 1, rb, re, J, re - rb, 0, yv === nothing ? C_NULL : pointer(yv)))
        end
    end
    return result
end

# hessian!(result, tape::CompiledHessian, input)             src/api/hessians.jl:86-97
# The reference's HessianTape IS the scalar tape of x -> gradient(f, x) (src/api/tape.jl:285-293): hessian! = its Jacobian, row i =
# one one-hot seeded reverse sweep, all seeds batched on the device.
function ReverseDiff.hessian!(result, ct::B200CompiledTape{<:HessianTape}, input::AbstractArray; rows = nothing)
    n = length(input)
    rb, re = rows === nothing ? (0, n) : rows
    H = result isa DiffResult ? DiffResults.hessian(result) : result
    g = result isa DiffResult ? DiffResults.gradient(result) : nothing
    GC.@preserve input H g begin
        set_inputs!(ct, (input,))
        check(ccall((:rdb_jacobian, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}),
                    ct.handle, 0, rb, re, H, re - rb, 0, g === nothing ? C_NULL : pointer(g)))   # value(output) of this tape = the gradient
    end
    if result isa DiffResult
        result = DiffResults.value!(result, func_hook(ct)(input))     # src/api/hessians.jl:96: the reference re-evaluates f too
    end
    return result
end

# The north star's mechanism: forward-over-reverse on the ARRAY tape of f (a compiled GradientTape), K tangent columns per sweep.
# cols = (begin, end) 0-based half-open column shard (multi-GPU).
function hessian_forward_over_reverse!(result, ct::B200CompiledTape{<:GradientTape}, input::AbstractArray; cols = nothing)
    n = length(input)
    cb, ce = cols === nothing ? (0, n) : cols
    H = result isa DiffResult ? DiffResults.hessian(result) : result
    g = result isa DiffResult ? DiffResults.gradient(result) : nothing
    val = Ref{Float64}(0.0)
    GC.@preserve input H g begin
        set_inputs!(ct, (input,))
        check(ccall((:rdb_hessian, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                    ct.handle, cb, ce, H, n, 0, g === nothing ? C_NULL : pointer(g), result isa DiffResult ? val : C_NULL))
    end
    result isa DiffResult && (result = DiffResults.value!(result, val[]))
    return result
end

# ---- multi-GPU result assembly: one process per GPU (MPI.jl or Distributed), NCCL inside librdb200.so ---------------------------
mutable struct Comm
    ptr::Ptr{Cvoid}
    world::Int
    rank::Int
end
"Rank 0 calls `unique_id()` and ships the 128 bytes to the other ranks with its own transport; then every rank calls `Comm(ct, ...)`."
function unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:rdb_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    return id
end
function Comm(ct::B200CompiledTape, world::Int, rank::Int, id::Vector{UInt8})
    p = GC.@preserve id ccall((:rdb_comm_create, LIB), Ptr{Cvoid}, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ct.ctx, world, rank, id)
    p == C_NULL && error("librdb200: ", last_error())
    c = Comm(p, world, rank)
    finalizer(x -> ccall((:rdb_comm_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ptr), c)
    return c
end
"Contiguous shard of `n` rows / columns owned by `rank` (api.py `shard_range`)."
shard_range(n::Int, world::Int, rank::Int) = (rank * (n ÷ world), rank < world - 1 ? (rank + 1) * (n ÷ world) : n)
shard_bounds(n::Int, world::Int) = Int64[[shard_range(n, world, r)[1] for r in 0:world-1]; n]
"""
    gather!(result_dev, block_dev, comm, kind, n_total, other_dim; T = Float64, root = -1)

`rdb_gather_result`: kind 0 = row blocks of jacobian!, kind 1 = column blocks of hessian_forward_over_reverse!; device pointers
(e.g. `CuPtr`s converted with `reinterpret(Ptr{Cvoid}, ...)`).  The assembled matrix is the reference's column-major result
(src/api/utils.jl:44-58).
"""
function gather!(result_dev::Ptr{Cvoid}, block_dev::Ptr{Cvoid}, comm::Comm, kind::Int, n_total::Int, other_dim::Int; T = Float64, root = -1)
    bounds = shard_bounds(n_total, comm.world)
    GC.@preserve bounds check(ccall((:rdb_gather_result, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Cvoid}, Cint),
        comm.ptr, T === Float64 ? 0 : 1, kind, block_dev, bounds, other_dim, result_dev, root))
    return result_dev
end

end # module
