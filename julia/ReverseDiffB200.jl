# ReverseDiffB200.jl — the reference-side binding a ReverseDiff maintainer would add to make
# `compile(tape; backend=:b200)` replay on librdb200.so.
#
# STATUS: UNVERIFIED.  There is no `julia` binary in the build image or on the GPU boxes, so this file has never
# been executed; the same lowering, written in Python (reversediff.jl_b200/tracked.py + program.py), is what the
# tests exercise.  It is kept here because INTEGRATION.md must show the binding for the reference's own language.
#
# What it does: walk `t.tape::Vector{AbstractInstruction}` (src/tape.jl:5-7,30-69), assign a buffer id to every
# TrackedArray/TrackedReal (objectid of its value array; reshape aliases share storage, src/tracked.jl:402-408),
# emit the flat tape program (layout: reversediff.jl_b200/program.py), and `ccall` the C ABI of include/rdb200.h.
module ReverseDiffB200

using ReverseDiff
using ReverseDiff: AbstractTape, GradientTape, JacobianTape, HessianTape, SpecialInstruction, ScalarInstruction,
                   TrackedArray, TrackedReal, value, deriv, istracked, input_hook, output_hook

const LIB = get(ENV, "RDB200_LIB", "librdb200.so")

struct RdbOptions
    fuse_elementwise::Int32
    seed_block::Int32
    use_graph::Int32
    profile::Int32
    gemm_f64_mode::Int32
    ozaki_slices::Int32
    reserved::NTuple{2,Int32}
end
RdbOptions(; fuse=1, seed_block=0, graph=1, profile=0, gemm_f64_mode=0, ozaki_slices=0) =
    RdbOptions(fuse, seed_block, graph, profile, gemm_f64_mode, ozaki_slices, (0, 0))

last_error() = unsafe_string(ccall((:rdb_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("librdb200: ", last_error())

mutable struct B200CompiledTape{T<:AbstractTape} <: AbstractTape
    tape::T
    ctx::Ptr{Cvoid}
    handle::Ptr{Cvoid}
    function B200CompiledTape(t::T, program::Vector{UInt8}; device=0, opts=RdbOptions()) where {T}
        ctx = ccall((:rdb_ctx_create, LIB), Ptr{Cvoid}, (Cint,), device)
        ctx == C_NULL && error("librdb200: ", last_error())          # no GPU => error, never a CPU fallback
        h = GC.@preserve program ccall((:rdb_tape_load, LIB), Ptr{Cvoid}, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t, Ref{RdbOptions}),
                                       ctx, program, length(program), opts)
        h == C_NULL && (ccall((:rdb_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx); error("librdb200: ", last_error()))
        ct = new{T}(t, ctx, h)
        finalizer(ct) do c
            ccall((:rdb_tape_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.handle)
            ccall((:rdb_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.ctx)
        end
        return ct
    end
end

ReverseDiff.input_hook(ct::B200CompiledTape) = input_hook(ct.tape)
ReverseDiff.output_hook(ct::B200CompiledTape) = output_hook(ct.tape)

# ---------------------------------------------------------------------------------------------- lowering
# Opcode table (program.py).  Anything else — @grad / ChainRules closures (src/macros.jl:190-206,335-350), det, inv,
# cat, tracker_∇broadcast, ScalarInstruction — throws: the engine has no CPU fallback.
const OPCODES = IdDict{Any,Int32}(
    (*) => 1, (+) => 2, (-) => 3, sum => 5, ReverseDiff.mean => 6, ReverseDiff.∇broadcast => 7, map => 8, broadcast => 9,
    (broadcast, +) => 10, (broadcast, -) => 11, (broadcast, *) => 12, getindex => 16)

"Symbolic Real used to run an elementwise closure once and record straight-line SSA bytecode (expr.py)."
struct Sym <: Real
    reg::Int32
    ops::Vector{NTuple{4,Int32}}
    fconsts::Vector{Float64}
end
# (methods for + - * / ^ tanh exp log sqrt sin cos abs2 abs max min inv on Sym emit one op each, exactly as
#  reversediff.jl_b200/expr.py does; comparisons on Sym throw "data-dependent control flow")

"""
    lower(t::AbstractTape) -> Vector{UInt8}

Emit the flat tape program.  Operand classes follow `capture` (src/tracked.jl:223-225):
TrackedArray -> TA; Array{<:TrackedReal} with origin links -> IDX, map[k] = x[k].index - 1 (src/tracked.jl:47-57,348-351);
plain arrays -> CONST (copied into the program).  `index_bound` caches (src/derivatives/propagation.jl:25-27) are shipped
as operand dims; the engine re-derives the clamp.
"""
function lower(t::AbstractTape)
    error("ReverseDiffB200.lower: write-only sketch — see reversediff.jl_b200/tracked.py for the executable lowering")
end

ReverseDiff.compile(t::AbstractTape, ::Val{:b200}; kw...) = B200CompiledTape(t, lower(t); kw...)

# ---------------------------------------------------------------------------------------------- replay API
set_input!(ct, slot, x::Array) =
    GC.@preserve x check(ccall((:rdb_tape_set_input, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint), ct.handle, slot, x, 0))

function seeded_forward!(ct::B200CompiledTape, input)
    xs = input isa Tuple ? input : (input,)
    for (i, x) in enumerate(xs)
        set_input!(ct, i - 1, x)            # value!(input_hook(t)[i], x)   src/api/tape.jl:40-44
    end
end

# gradient!(result, tape, input)            src/api/gradients.jl:78-82
function ReverseDiff.gradient!(result, ct::B200CompiledTape{<:GradientTape}, input)
    seeded_forward!(ct, input)
    rs = result isa Tuple ? result : (result,)
    arrs = map(r -> r isa ReverseDiff.DiffResult ? ReverseDiff.DiffResults.gradient(r) : r, rs)
    ptrs = Ptr{Cvoid}[pointer(a) for a in arrs]
    val = Ref{Float64}(0.0)
    GC.@preserve arrs check(ccall((:rdb_gradient, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cvoid}), ct.handle, ptrs, 0, val))
    return result                            # DiffResult values: DiffResults.value!(r, val[]) as src/api/utils.jl:96-100
end

# jacobian!(result, tape, input)            src/api/jacobians.jl:120-124 ; rows = seed-row shard for multi-GPU
function ReverseDiff.jacobian!(result::AbstractMatrix, ct::B200CompiledTape{<:JacobianTape}, input; rows=nothing, slot=0)
    seeded_forward!(ct, input)
    m = length(output_hook(ct))
    rb, re = rows === nothing ? (0, m) : rows
    GC.@preserve result check(ccall((:rdb_jacobian, LIB), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}), ct.handle, slot, rb, re, result, size(result, 1), 0, C_NULL))
    return result
end

# hessian!(result, tape, input)             src/api/hessians.jl:86-90 ; cols = column shard for multi-GPU
function ReverseDiff.hessian!(result::AbstractMatrix, ct::B200CompiledTape{<:HessianTape}, input::AbstractArray; cols=nothing)
    seeded_forward!(ct, input)
    n = length(input)
    cb, ce = cols === nothing ? (0, n) : cols
    GC.@preserve result check(ccall((:rdb_hessian, LIB), Cint,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}, Ptr{Cvoid}), ct.handle, cb, ce, result, n, 0, C_NULL, C_NULL))
    return result
end

end # module
