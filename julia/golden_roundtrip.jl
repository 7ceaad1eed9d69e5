# golden_roundtrip.jl — the one-assert check of the Julia lowering against the Python recorder (tests/golden/make_program_hashes.py).
#
#     julia --project=<env with ReverseDiff> julia/golden_roundtrip.jl
#
# Records the five BASELINE configs with ReverseDiff.jl itself (same functions, same shapes, the same exactly-representable constants),
# lowers each tape with ReverseDiffB200.lower and compares the SHA-256 of the bytes with tests/golden/<cfg>.program.sha256.  Needs no
# GPU and no librdb200.so.  UNEXECUTED so far: no Julia runtime exists in the build image.
include(joinpath(@__DIR__, "ReverseDiffB200.jl"))
using .ReverseDiffB200
using ReverseDiff
using ReverseDiff: GradientTape, JacobianTape, @forward
using SHA

pattern(rows, cols, a = 7, b = 13) = [((a * i + b * j) % 17 - 8) / 16 for i in 0:rows-1, j in 0:cols-1]

f_gradient_example(a, b) = sum(a' * b + a * b')                       # examples/gradient.jl:8
make_f_mlp(X) = (W1, b1, W2, b2) -> sum(tanh.(W2 * tanh.(W1 * X .+ b1) .+ b2))
make_f_jacobian(A, B) = x -> (A * x + x) .* tanh.(B * x)
rosen = @forward((o, e) -> 100.0 * (e - o^2)^2 + (1.0 - o)^2)
f_rosenbrock(x) = sum(map(rosen, x[1:2:end], x[2:2:end]))

function programs()
    out = Dict{String,Vector{UInt8}}()
    out["cfg1"] = ReverseDiffB200.lower(GradientTape(f_gradient_example, (ones(100, 100), ones(100, 100))))
    out["cfg2_f64"] = ReverseDiffB200.lower(GradientTape(f_gradient_example, (ones(4096, 4096), ones(4096, 4096))))
    out["cfg2_f32"] = ReverseDiffB200.lower(GradientTape(f_gradient_example, (ones(Float32, 4096, 4096), ones(Float32, 4096, 4096))))
    d = h = o = 8; N = 32
    out["cfg3"] = ReverseDiffB200.lower(GradientTape(make_f_mlp(pattern(d, N)), (ones(h, d), ones(h), ones(o, h), ones(o))))
    n = 16
    out["cfg4"] = ReverseDiffB200.lower(JacobianTape(make_f_jacobian(pattern(n, n), pattern(n, n, 5, 3)), ones(n)))
    # cfg 5: the ARRAY tape of f (what hessian_forward_over_reverse! replays) is the GradientTape of f
    out["cfg5"] = ReverseDiffB200.lower(GradientTape(f_rosenbrock, ones(16384)))
    return out
end

function main()
    golden = joinpath(@__DIR__, "..", "tests", "golden")
    ok = true
    for (name, raw) in sort(collect(programs()); by = first)
        want, len = split(strip(read(joinpath(golden, "$name.program.sha256"), String)))
        got = bytes2hex(sha256(raw))
        same = got == want && length(raw) == parse(Int, len)
        println(rpad(name, 10), same ? "identical to the Python recorder's program" : "DIFFERS: $(length(raw)) bytes, sha256 $got (want $len bytes, $want)")
        ok &= same
    end
    ok || error("the Julia lowering and the Python recorder disagree")
end

main()
