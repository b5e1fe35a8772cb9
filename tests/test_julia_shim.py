"""CPU: the Julia shim (julia/ApproximateGPsB200Ext.jl) cannot be executed in this image (no julia), so every `ccall` and every
struct mirror in it is parsed and checked against include/agp.h: symbol exists, arity, and the C type class of every argument,
return value and struct field.  The ctypes binding (approximategps.jl_b200/_lib.py) is checked against the same parse."""
import ctypes as C
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HEADER = os.path.join(ROOT, "include", "agp.h")
SHIM = os.path.join(ROOT, "julia", "ApproximateGPsB200Ext.jl")

STRUCT_NAMES = {"AgpKernelComponent": "agp_kernel_component", "AgpKernel": "agp_kernel", "AgpLikelihood": "agp_likelihood", "AgpExpectation": "agp_expectation", "AgpSvgpParams": "agp_svgp_params",
                "AgpSvgpGrads": "agp_svgp_grads", "AgpLaplaceProblem": "agp_laplace_problem", "AgpLaplaceResult": "agp_laplace_result"}


def c_class(t: str) -> str:
    t = t.strip()
    if "*" in t or t.startswith("agp_newton_callback"):
        return "cstr" if re.fullmatch(r"const\s+char\s*\*", t) else "ptr"
    t = t.replace("const", "").strip()
    return {"int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "double": "f64", "void": "void"}.get(t, "struct:" + t)


def parse_header():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    structs = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.fullmatch(r"(.*?)(\w+)", decl, flags=re.S)
            fields.append((m.group(2), c_class(m.group(1))))
        structs[name] = fields
    src_nos = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", src, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"([\w\s\*]+?)\b(agp_\w+)\s*\(([^()]*)\)\s*;", src_nos):
        if "typedef" in ret:
            continue
        args = args.strip()
        alist = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        classes = []
        for a in alist:
            m = re.fullmatch(r"(.*?)(\w+)", a, flags=re.S)  # strip the parameter name
            classes.append(c_class(m.group(1) if ("*" in a or " " in a) else a))
        protos[name] = (c_class(ret), classes)
    return protos, structs


def jl_class(t: str) -> str:
    t = t.strip()
    if t.startswith("Ptr{") or t.startswith("Ref{"):
        return "ptr"
    if t in STRUCT_NAMES:
        return "struct:" + STRUCT_NAMES[t]
    return {"Int32": "i32", "Int64": "i64", "UInt64": "u64", "Float64": "f64", "Cstring": "cstr", "Cvoid": "void"}[t]


def split_top(s: str):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out if x.strip()]


def parse_shim():
    src = open(SHIM).read()
    src = re.sub(r"#[^\n]*", "", src)
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*lib\),\s*(\w+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        calls.append((m.group(1), jl_class(m.group(2)), [jl_class(t) for t in split_top(src[m.end():i - 1])]))
    structs = {}
    for name, body in re.findall(r"(?:mutable\s+)?struct\s+(Agp\w+)\n(.*?)\nend", src, flags=re.S):
        structs[name] = [(f, jl_class(t)) for f, t in re.findall(r"^\s*(\w+)::([\w\{\}]+)\s*$", body, flags=re.M)]
    return calls, structs


def compatible(c: str, j: str) -> bool:
    # a by-reference struct argument (const agp_kernel*) is a pointer on both sides; a C string return is Cstring
    return c == j


def test_every_ccall_matches_the_header():
    protos, _ = parse_header()
    calls, _ = parse_shim()
    assert len(calls) >= 25
    for sym, ret, args in calls:
        assert sym in protos, f"ccall of {sym}: not declared in include/agp.h"
        cret, cargs = protos[sym]
        assert compatible(cret, ret), (sym, "return", cret, ret)
        assert len(cargs) == len(args), (sym, "arity", cargs, args)
        for k, (c, j) in enumerate(zip(cargs, args)):
            assert compatible(c, j), (sym, f"argument {k + 1}", c, j)


def test_struct_mirrors_match_the_header():
    _, cstructs = parse_header()
    _, jstructs = parse_shim()
    assert set(jstructs) == set(STRUCT_NAMES)
    for jname, fields in jstructs.items():
        cfields = cstructs[STRUCT_NAMES[jname]]
        assert [n for n, _ in fields] == [n for n, _ in cfields], (jname, fields, cfields)
        for (n, j), (_, c) in zip(fields, cfields):
            assert compatible(c, j), (jname, n, c, j)


def test_the_shim_covers_the_reference_entry_points():
    """SURVEY.md section 8b "Julia methods to intercept": each has a method definition in the shim."""
    src = open(SHIM).read()
    for needle in ("function AbstractGPs.elbo(sva::SparseVariationalApproximation, lfx::LatentFiniteGP", "ChainRulesCore.rrule(::typeof(AbstractGPs.elbo)",
                   "_prior_kl(sva::SparseVariationalApproximation)", "function AbstractGPs.posterior(sva::SparseVariationalApproximation)",
                   "StatsBase.mean_and_var(f::SVAPosterior", "StatsBase.mean_and_cov(f::SVAPosterior", "Statistics.cov(f::SVAPosterior, x::AbstractVector, y::AbstractVector)",
                   "function AbstractGPs.posterior(la::LaplaceApproximation, lfx::LatentFiniteGP, ys)", "ApproximateGPs.approx_lml(la::LaplaceApproximation",
                   "function laplace_f_and_lml(lfx::LatentFiniteGP, ys; newton_kwargs...)", "ChainRulesCore.rrule(::typeof(newton_inner_loop)",
                   "ChainRulesCore.frule((_, _, _, ΔK), ::typeof(newton_inner_loop)", "function laplace_f_cov(cache::LazyLaplaceCache)",
                   "points(x::RowVecs)", "kernel_tangent(", "mean_tangent(", "input_tangent(", "lik_tangent(", "mutable struct LazyLaplaceCache", "laplace_call_K(",
                   "PosDefException(Int(ccall((:agp_last_error_info"):
        assert needle in src, needle


def test_ctypes_binding_matches_the_header():
    from agp_b200 import _lib as L

    protos, cstructs = parse_header()
    assert set(L.SYMBOLS) == set(protos), set(L.SYMBOLS) ^ set(protos)

    def cls(t):
        if t is None:
            return "void"
        if t is C.c_char_p:
            return "cstr"
        if t in (C.c_void_p,) or isinstance(t, type) and issubclass(t, (C._Pointer, C._CFuncPtr)):
            return "ptr"
        return {C.c_int32: "i32", C.c_int64: "i64", C.c_uint64: "u64", C.c_double: "f64"}[t]

    for name, (res, args) in L.SYMBOLS.items():
        cret, cargs = protos[name]
        assert cls(res) == cret, (name, res, cret)
        assert [cls(a) for a in args] == cargs, (name, [cls(a) for a in args], cargs)
    for pyname, cname in (("AgpKernelComponent", "agp_kernel_component"), ("AgpKernel", "agp_kernel"), ("AgpLikelihood", "agp_likelihood"), ("AgpExpectation", "agp_expectation"), ("AgpSvgpParams", "agp_svgp_params"),
                          ("AgpSvgpGrads", "agp_svgp_grads"), ("AgpLaplaceProblem", "agp_laplace_problem"), ("AgpLaplaceResult", "agp_laplace_result")):
        fields = getattr(L, pyname)._fields_
        assert [f[0] for f in fields] == [n for n, _ in cstructs[cname]], (pyname,)
