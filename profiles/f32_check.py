import sys, numpy as np, time
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
rd = g.load_package(); orc = g.load_oracle()
n = 4096
for seed in (2, 7):
    rng = np.random.default_rng(1)
    a0 = np.asfortranarray(rng.random((n, n)).astype(np.float32)); b0 = np.asfortranarray(rng.random((n, n)).astype(np.float32))
    tape = rd.GradientTape(rd.workloads.f_gradient_example, (a0, b0))
    ct = rd.compile(tape)
    a, b = rd.workloads.inputs_gradient_example(n, np.float32, seed=seed)
    ga, gb = rd.gradient(ct, (a, b))
    val32, (oa, ob) = orc.OracleTape(tape.program_bytes()).gradient([a, b])
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    ta = b64.sum(1)[:, None] + b64.sum(0)[None, :]
    rel = lambda x, y: float(np.max(np.abs(x.astype(np.float64) - y.astype(np.float64))) / np.max(np.abs(y)))
    print("seed", seed, "gpu vs f32 oracle", rel(ga, oa), rel(gb, ob), "| gpu vs f64 truth", rel(ga, ta), "| f32 oracle vs truth", rel(oa, ta))
# signed data
rng = np.random.default_rng(3)
a = np.asfortranarray(rng.standard_normal((n, n)).astype(np.float32)); b = np.asfortranarray(rng.standard_normal((n, n)).astype(np.float32))
tape = rd.GradientTape(rd.workloads.f_gradient_example, (a, b)); ct = rd.compile(tape)
ga, gb = rd.gradient(ct, (a, b))
_, (oa, ob) = orc.OracleTape(tape.program_bytes()).gradient([a, b])
print("randn: gpu vs f32 oracle", rel(ga, oa), rel(gb, ob))
