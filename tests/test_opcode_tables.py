"""The scalar-function opcodes (`ExprOp`) live in five places - the recorder (`program.py`), the engine (`csrc/program.h`), the two
oracles (`oracle/replay.py`, `oracle/scalar_block.c`) and the Julia shim (`julia/ReverseDiffB200.jl`).  They are separate
spellings on purpose (engine and oracle parse the wire format independently); this test keeps their numbering identical."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_enum(path, first):
    """names of a C enum that starts with `first = 1`, in order -> {name: value}"""
    src = open(path).read()
    m = re.search(r"enum[^{]*\{\s*(" + first + r"\s*=\s*1.*?)\}", src, re.S)
    assert m, path
    out, val = {}, 0
    for tok in m.group(1).replace("\n", " ").split(","):
        tok = tok.strip()
        if not tok:
            continue
        if "=" in tok:
            name, rhs = [p.strip() for p in tok.split("=")]
            val = out[rhs] if rhs in out else int(rhs)
        else:
            name, val = tok, val + 1
        out[name] = val
    return out


def test_scalar_opcodes_agree_everywhere(rd, orc):
    P = rd.program
    py = {n: getattr(P, n) for n in dir(P) if re.fullmatch(r"E_[A-Z0-9]+", n) and isinstance(getattr(P, n), int)}
    assert py["E_LAST"] == max(v for n, v in py.items() if n != "E_LAST")
    assert sorted(v for n, v in py.items() if n != "E_LAST") == list(range(1, py["E_LAST"] + 1))      # dense numbering, no gaps
    assert set(P.E_NAMES) == {v for n, v in py.items() if n != "E_LAST"}

    eng = _c_enum(os.path.join(ROOT, "reversediff.jl_b200", "csrc", "program.h"), "E_CONST")
    assert eng == py

    corc = _c_enum(os.path.join(ROOT, "oracle", "scalar_block.c"), "E_CONST")
    assert corc == {n: v for n, v in py.items() if n != "E_LAST"}

    porc = {n: getattr(orc, n) for n in dir(orc) if re.fullmatch(r"E_[A-Z0-9]+", n) and isinstance(getattr(orc, n), int)}
    assert porc == {n: v for n, v in py.items() if n != "E_LAST"}

    jl = open(os.path.join(ROOT, "julia", "ReverseDiffB200.jl")).read()
    jvals = {}
    for m in re.finditer(r"const\s+((?:E_[A-Z0-9]+\s*,\s*)*E_[A-Z0-9]+)\s*=\s*Int32\.\(\(([^)]*)\)\)", jl):
        names = [n.strip() for n in m.group(1).split(",")]
        vals = [int(v) for v in m.group(2).split(",") if v.strip()]
        assert len(names) == len(vals)
        jvals.update(zip(names, vals))
    assert jvals == {n: v for n, v in py.items() if n != "E_LAST"}


def test_instruction_opcodes_agree_everywhere(rd, orc):
    P = rd.program
    py = {n: getattr(P, n) for n in dir(P) if re.fullmatch(r"OP_[A-Z0-9_]+", n) and isinstance(getattr(P, n), int)}
    assert sorted(py.values()) == list(range(1, max(py.values()) + 1))
    src = open(os.path.join(ROOT, "reversediff.jl_b200", "csrc", "program.h")).read()
    m = re.search(r"enum Opcode[^{]*\{(.*?)\};", src, re.S)
    body = re.sub(r"//[^\n]*", "", m.group(1))
    eng = {n: int(v) for n, v in re.findall(r"(OP_[A-Z0-9_]+)\s*=\s*(\d+)", body)}
    assert eng == py
    porc = {n: getattr(orc, n) for n in dir(orc) if re.fullmatch(r"OP_[A-Z0-9_]+", n) and isinstance(getattr(orc, n), int)}
    assert porc == py
    jl = open(os.path.join(ROOT, "julia", "ReverseDiffB200.jl")).read()
    jvals = {}
    for m in re.finditer(r"const\s+((?:OP_[A-Z0-9_]+\s*,\s*)*OP_[A-Z0-9_]+)\s*=\s*Int32\.?\(\(?([^)]*)\)", jl):
        names = [n.strip() for n in m.group(1).split(",")]
        vals = [int(v) for v in m.group(2).split(",") if v.strip()]
        assert len(names) == len(vals), m.group(0)
        jvals.update(zip(names, vals))
    assert jvals == py
