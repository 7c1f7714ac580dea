"""Static check of julia/OpenQuantumBaseB200.jl against include/oqb.h (no Julia in this image, so the glue cannot be executed):
every `ccall((:name, LIB), Cint, (argtypes…), …)` must name a symbol the library exports, pass as many arguments as the C
prototype declares, and use an argument type of the same class (pointer / 32-bit int / 64-bit int / size_t / double) in
every position — the kind of drift a never-run binding accumulates."""
import ctypes as C
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
JL = os.path.join(ROOT, "openquantumbase.jl_b200", "julia", "OpenQuantumBaseB200.jl")
HDR = os.path.join(ROOT, "include", "oqb.h")


def _balanced(text, start):
    """text[start] == '(' → index one past its matching ')'."""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == "(":
            depth += 1
        elif text[i] == ")":
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced parentheses")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls():
    src = open(JL, encoding="utf-8").read()
    src = re.sub(r"#=.*?=#", "", src, flags=re.S)
    src = "\n".join(l.split(" # ")[0] for l in src.split("\n"))
    calls = []
    for m in re.finditer(r"ccall\(", src):
        end = _balanced(src, m.end() - 1)
        args = _split_top(src[m.end():end - 1])
        sym = re.match(r"\(\s*:(\w+)\s*,\s*LIB\s*\)", args[0])
        assert sym, f"unrecognised ccall target: {args[0]!r}"
        assert args[2].startswith("(") and args[2].endswith(")"), args[2]
        types = _split_top(args[2][1:-1])
        calls.append((sym.group(1), args[1], types, len(args) - 3, src.count("\n", 0, m.start()) + 1))
    return calls


def c_prototypes():
    hdr = open(HDR, encoding="utf-8").read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", "", hdr)
    protos = {}
    for m in re.finditer(r"\b(int|const char\s*\*|void\s*\*|cudaStream_t|void)\s+(oqb_\w+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        params = [p.strip() for p in _split_top(m.group(3)) if p.strip() and p.strip() != "void"]
        protos[m.group(2)] = (m.group(1), params)
    return protos


def c_class(param):
    if "*" in param or "(" in param:
        return "ptr"
    t = param.replace("const", " ").split()
    t = " ".join(t[:-1]) if len(t) > 1 else t[0]  # drop the parameter name
    if t.endswith("_fn"):  # function-pointer typedefs (oqb_spectrum_fn, oqb_corr_fn): @cfunction pointers on the Julia side
        return "ptr"
    return {"int": "i32", "int32_t": "i32", "unsigned": "i32", "int64_t": "i64", "long long": "i64", "uint64_t": "i64", "size_t": "size",
            "double": "f64"}.get(t, t)


def jl_class(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t in ("Cstring", "Ptr", "Any"):
        return "ptr"
    return {"Cint": "i32", "Int32": "i32", "Cuint": "i32", "Int64": "i64", "UInt64": "i64", "Clonglong": "i64", "Csize_t": "size", "Float64": "f64",
            "Cdouble": "f64"}.get(t, t)


def test_the_glue_has_calls_and_the_header_parses():
    calls, protos = julia_ccalls(), c_prototypes()
    assert len(calls) >= 40 and len(protos) >= 80
    assert {"oqb_dense_rhs", "oqb_liouv_rhs", "oqb_traj_matvec", "oqb_comm_init_all"} <= {c[0] for c in calls}


def test_every_ccall_matches_its_c_prototype():
    protos = c_prototypes()
    import oqb_b200  # noqa: F401
    from oqb_b200 import _lib
    lib = _lib.load()
    problems = []
    for name, ret, types, nargs, line in julia_ccalls():
        where = f"OpenQuantumBaseB200.jl:{line} {name}"
        if not hasattr(lib, name):
            problems.append(f"{where}: not exported by liboqb_b200.so")
            continue
        if name not in protos:
            problems.append(f"{where}: not declared in include/oqb.h")
            continue
        cret, params = protos[name]
        if len(types) != len(params):
            problems.append(f"{where}: {len(types)} argument types, the prototype has {len(params)} parameters")
            continue
        if nargs != len(types):
            problems.append(f"{where}: {nargs} arguments passed for {len(types)} declared types")
        want_ret = {"int": "Cint", "const char*": "Cstring"}.get(re.sub(r"\s+", "", cret).replace("constchar*", "const char*"), None)
        if want_ret and ret.strip() != want_ret:
            problems.append(f"{where}: returns {ret}, the prototype returns {cret}")
        for k, (jt, cp) in enumerate(zip(types, params)):
            jc, cc = jl_class(jt), c_class(cp)
            if jc != cc and not (jc == "i64" and cc == "size") and not (jc == "size" and cc == "i64"):
                problems.append(f"{where}: argument {k + 1} is {jt} ({jc}), the prototype says `{cp}` ({cc})")
    assert not problems, "\n".join(problems)


def ctypes_class(t):
    if t is None:
        return "void"
    if t in (C.c_void_p, C.c_char_p) or hasattr(t, "_type_") and not isinstance(t._type_, str) or "CFunctionType" in str(getattr(t, "__mro__", "")):
        return "ptr"
    return {C.c_int: "i32", C.c_int32: "i32", C.c_uint: "i32", C.c_int64: "i64", C.c_uint64: "i64", C.c_longlong: "i64", C.c_size_t: "size",
            C.c_double: "f64"}.get(t, str(t))


def test_the_python_binding_matches_the_header_one_to_one():
    """openquantumbase.jl_b200/_lib.py SIGNATURES against include/oqb.h: same set of functions, same argument count, same
    argument class per position."""
    protos = c_prototypes()
    import oqb_b200  # noqa: F401
    from oqb_b200 import _lib
    problems = []
    for name in sorted(set(protos) - set(_lib.SIGNATURES)):
        problems.append(f"{name}: declared in include/oqb.h, no ctypes signature")
    for name, argtypes in _lib.SIGNATURES.items():
        if name not in protos:
            problems.append(f"{name}: ctypes signature without a declaration in include/oqb.h")
            continue
        params = protos[name][1]
        if len(params) != len(argtypes):
            problems.append(f"{name}: {len(argtypes)} ctypes arguments, the prototype has {len(params)}")
            continue
        for k, (at, cp) in enumerate(zip(argtypes, params)):
            ac, cc = ctypes_class(at), c_class(cp)
            if ac != cc and {ac, cc} != {"i64", "size"}:
                problems.append(f"{name}: argument {k + 1} is {at} ({ac}), the prototype says `{cp}` ({cc})")
    assert not problems, "\n".join(problems)
