"""CPU checks of the C-ABI boundary: the library loads and exports every symbol include/rdb200.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest


def _declared_symbols():
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "rdb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rdb_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_exported(rd):
    lib = ctypes.CDLL(rd.engine.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"librdb200.so does not export {n}"
    assert sorted(rd.engine.EXPORTED_SYMBOLS) == names      # the ctypes binding covers the whole header


def test_no_cpu_fallback(rd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rd.engine.EngineError, match="no CPU fallback"):
        rd.engine.Context(0)


def test_oracle_is_not_imported_by_the_product(rd):
    """The oracle is test infrastructure: no product source may import it, dlopen it, spawn it or even name a path under oracle/
    in code (prose in comments / docstrings is fine), and importing the product must leave no oracle module or oracle-built library
    in the process."""
    import io, sys, tokenize
    root = os.path.join(os.path.dirname(__file__), "..", "reversediff.jl_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                src = open(path).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                # every token that is code or a string literal other than a docstring-position string
                toks = list(tokenize.generate_tokens(io.StringIO(src).readline))
                for i, t in enumerate(toks):
                    if t.type == tokenize.NAME and t.string == "oracle":
                        raise AssertionError(f"{f}:{t.start[0]} names `oracle` in code")
                    if t.type == tokenize.STRING and re.search(r"oracle[/\\.]|libscalar_block|scalar_block\.c", t.string):
                        prev = next((p_ for p_ in reversed(toks[:i]) if p_.type not in (tokenize.NL, tokenize.NEWLINE, tokenize.INDENT,
                                                                                         tokenize.DEDENT, tokenize.COMMENT)), None)
                        is_doc = prev is None or prev.string in (":",) or prev.type == tokenize.ENCODING
                        assert is_doc, f"{f}:{t.start[0]} has a string literal naming the oracle outside a docstring"
            elif f.endswith((".cu", ".h", ".cuh", ".cpp")):
                code = re.sub(r"//[^\n]*|/\*.*?\*/", "", open(path).read(), flags=re.S)
                assert "oracle" not in code and "scalar_block.so" not in code, f"{f} names the oracle in code"
    assert not [m for m in sys.modules if m == "oracle" or m.startswith("oracle.")] or "oracle.replay" in sys.modules   # (the test session itself may have it)
    import subprocess
    probe = ("import sys, os; sys.path.insert(0, %r); import __graft_entry__ as g; g.load_package(); "
             "bad = [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; "
             "maps = open('/proc/self/maps').read(); "
             "assert not bad and 'libscalar_block' not in maps and '/oracle/' not in maps, (bad,); print('clean')"
             % os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
    r = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "clean" in r.stdout, r.stdout + r.stderr


def test_col_panel_planner(rd):
    """Host-only planner behind the panelled uploads / panelled last accumulations (engine.cu plan_col_panels): the panels tile
    the columns exactly, every panel but the last is a whole number of 64-column tiles, and the 128x64-tile waves they cost on
    148 SMs stay within one wave of the unsplit GEMM (four equal panels of a 4096^2 result would cost 16 against 13.8)."""
    import ctypes as C
    lib = rd.engine.lib()
    for rows, cols in [(4096, 4096), (2048, 2048), (8192, 8192), (4096, 2048), (1024, 65536), (3000, 5000), (2048, 16384)]:        # sizes that reach the panel routes (>= 32 MB, >= 2048 columns)
        w = (C.c_int64 * 4)()
        n = lib.rdb_plan_col_panels(rows, cols, w)
        widths = [int(x) for x in w if x > 0]
        assert n == len(widths) and 1 <= n <= 4
        assert sum(widths) == cols
        assert all(x % 64 == 0 for x in widths[:-1])
        rt = -(-rows // 128)
        waves = sum(-(-(rt * -(-x // 64)) // 148) for x in widths)
        ideal = rt * -(-cols // 64) / 148.0
        assert waves <= -(-ideal // 1) + 1 or waves <= ideal * 1.12, (rows, cols, widths, waves, ideal)      # within one wave of the unsplit GEMM
    w = (C.c_int64 * 4)()
    assert lib.rdb_plan_col_panels(4096, 4096, w) == 3 and list(w)[:3] == [1472, 1472, 1152]


def test_bundle_scheduler(rd):
    """Host-only list scheduler of the scalar-block bundle kernels: every node lands in exactly one (bundle, lane) slot, at most 32
    per bundle, strictly after all its predecessors; a chain with independent side work needs exactly as many bundles as the chain
    is long (critical path first), a wide layer of independent nodes ceil(n/32); cycles are refused."""
    import ctypes as C
    import numpy as np
    lib = rd.engine.lib()

    def run(n, edges):
        off = [0] * (n + 1)
        for u, _ in edges:
            off[u + 1] += 1
        for i in range(n):
            off[i + 1] += off[i]
        fill = off[:-1].copy(); succ = [0] * len(edges)
        for u, v in edges:
            succ[fill[u]] = v; fill[u] += 1
        o = (C.c_int64 * (n + 1))(*off); s_ = (C.c_int32 * max(len(edges), 1))(*succ)
        b = (C.c_int32 * max(n, 1))(); l = (C.c_int32 * max(n, 1))()
        nb = lib.rdb_schedule_bundles(n, o, s_, b, l)
        return nb, list(b)[:n], list(l)[:n]

    def check(n, edges, nb, b, l):
        assert nb >= 0 and all(0 <= x < nb for x in b) and all(0 <= x < 32 for x in l)
        assert len({(x, y) for x, y in zip(b, l)}) == n                       # one slot per node
        for u, v in edges:
            assert b[u] < b[v]                                                # strictly after every predecessor

    # rosenbrock-like: a chain acc_k -> acc_{k+1} of length m, each link also fed by an independent 5-node term
    m = 300
    edges, n = [], 0
    chain = []
    for k in range(m):
        term = list(range(n, n + 5)); n += 5
        edges += [(term[i], term[i + 1]) for i in range(4)]
        acc = n; n += 1
        edges.append((term[-1], acc))
        if chain:
            edges.append((chain[-1], acc))
        chain.append(acc)
    nb, b, l = run(n, edges)
    check(n, edges, nb, b, l)
    assert nb == m + 5                                                        # the critical path: first term (5) + m chain links
    # a wide independent layer
    nb, b, l = run(1000, [])
    check(1000, [], nb, b, l)
    assert nb == 32 and max(b.count(x) for x in set(b)) == 32                # ceil(1000 / 32) bundles, full ones first
    # random DAG
    rng = np.random.default_rng(0)
    n = 2000
    edges = sorted({(int(u), int(v)) for u, v in zip(rng.integers(0, n, 6000), rng.integers(0, n, 6000)) if u < v})
    nb, b, l = run(n, edges)
    check(n, edges, nb, b, l)
    # a cycle
    assert run(3, [(0, 1), (1, 2), (2, 0)])[0] == -1
