"""The Julia wrapper cannot be executed in this image (no Julia): what CAN be checked mechanically is checked here --
every `ccall` of julia/GPULoglike.jl names a function include/demcmc_b200.h declares, with the same number of
arguments and compatible argument types; the Julia mirrors of the C structs list the header's fields in the header's
order with compatible types; the ABI version matches; and the file does not redefine the package's own constructor
(`DEModel(args...; ...)` would replace src/structs.jl:176-189 for every CPU model)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "demcmc_b200.h")).read()
JL = open(os.path.join(ROOT, "julia", "GPULoglike.jl")).read()


def strip_comments(c):
    return re.sub(r"/\*.*?\*/", " ", c, flags=re.S)


def header_prototypes():
    protos = {}
    for m in re.finditer(r"\b(?:int|const char \*)\s*(demcmc_\w+)\s*\(([^;{]*?)\)\s*;", strip_comments(HDR)):
        args = m.group(2).strip()
        protos[m.group(1)] = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
    return protos


def header_struct(name):
    c = strip_comments(HDR)
    end = re.search(r"\}\s*" + name + r"\s*;", c).start()
    body = c[c.rindex("typedef struct", 0, end):end].split("{", 1)[1]
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        first, *rest = [v.strip() for v in decl.split(",")]
        m = re.match(r"(.*?)(\**)\s*(\w+)(\[\w*\])?$", first)
        ctype = m.group(1).replace("const", "").strip()
        fields.append((m.group(3), ctype, bool(m.group(2))))
        for nm in rest:
            fields.append((nm.lstrip("* "), ctype, nm.startswith("*")))
    return fields


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls():
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+), LIBDEMCMC\)", JL):
        i = m.end()
        depth, j = 1, i
        while depth:
            depth += {"(": 1, ")": -1}.get(JL[j], 0)
            j += 1
        parts = split_top(JL[i:j - 1].lstrip(", "))
        ret, argt, args = parts[0], parts[1], parts[2:]
        types = split_top(argt.strip()[1:-1].rstrip(","))
        calls.append((m.group(1), ret, [t for t in types if t], args))
    return calls


C2J = {"int32_t": {"Int32", "Cint"}, "int": {"Cint", "Int32"}, "int64_t": {"Int64"}, "uint64_t": {"UInt64"}, "double": {"Float64"},
       "uint8_t": {"UInt8"}}


def compatible(c_arg, jl_type):
    c = c_arg.replace("const", "").strip()
    if "*" in c or "[" in c:
        base = c.replace("*", " ").split()[0]
        if base.startswith("demcmc_handle"):
            return jl_type in ("Ptr{Cvoid}", "Ref{Ptr{Cvoid}}")
        if base.startswith("demcmc_"):                       # struct by pointer: Ref{CStruct}
            return jl_type.startswith("Ref{C") or jl_type.startswith("Ptr{C")
        inner = C2J.get(base, set())
        return any(jl_type in (f"Ptr{{{t}}}", f"Ref{{{t}}}") for t in inner)
    return jl_type in C2J.get(c.split()[0], set())


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 15
    for name, ret, types, args in calls:
        assert name in protos, f"ccall of {name}: not declared in include/demcmc_b200.h"
        assert len(types) == len(protos[name]), f"{name}: {len(types)} argument types in the ccall, {len(protos[name])} parameters in the header"
        assert len(args) == len(types), f"{name}: {len(args)} arguments passed for {len(types)} argument types"
        assert ret == ("Cstring" if name == "demcmc_last_error" else "Cint"), (name, ret)
        for c_arg, jt in zip(protos[name], types):
            assert compatible(c_arg, jt), f"{name}: header parameter `{c_arg}` bound as {jt}"


def julia_struct(name):
    body = re.search(r"struct " + name + r"\n(.*?)\nend", JL, flags=re.S).group(1)
    return [(m.group(1), m.group(2)) for m in re.finditer(r"^\s*(\w+)::([\w{}]+)", body, flags=re.M)]


def test_struct_mirrors_follow_the_header_field_by_field():
    for cname, jname in (("demcmc_prior", "CPrior"), ("demcmc_model", "CModel"), ("demcmc_config", "CConfig")):
        hf, jf = header_struct(cname), julia_struct(jname)
        assert [f[0] for f in hf] == [f[0] for f in jf], (cname, [f[0] for f in hf], [f[0] for f in jf])
        for (nm, ctype, ptr), (_, jt) in zip(hf, jf):
            if ptr:
                assert jt.startswith("Ptr{"), (cname, nm, jt)
            else:
                assert jt in C2J[ctype], (cname, nm, ctype, jt)
    # ... and the ctypes mirror the tests actually drive lists the same fields
    import demcmc_b200 as D
    for cname, cls in (("demcmc_prior", D._ffi.Prior), ("demcmc_model", D._ffi.Model), ("demcmc_config", D._ffi.Config), ("demcmc_counters", D._ffi.Counters)):
        assert [f[0] for f in header_struct(cname)] == [f[0] for f in cls._fields_], cname


def test_abi_version_and_symbols():
    ver = int(re.search(r"#define DEMCMC_ABI_VERSION (\d+)", HDR).group(1))
    assert int(re.search(r"const DEMCMC_ABI_VERSION = Int32\((\d+)\)", JL).group(1)) == ver
    import demcmc_b200 as D
    assert D._ffi.ABI_VERSION == ver
    assert sorted(header_prototypes()) == sorted(D._ffi.SYMBOLS)


def test_the_package_constructor_is_not_redefined():
    # a method DEModel(args...; ...) in the wrapper would REPLACE the package's keyword constructor (Julia does not
    # dispatch on keyword types); the GPU model must come in through its own entry points
    assert not re.search(r"function DEModel\(args\.\.\.", JL)
    assert "function GPUDEModel(" in JL and "function DEModel(loglike::GPULoglike;" in JL
    # ... and the initial particles must not go through evaluate_fitness! (it would call the plugin object on the host)
    body = JL[JL.index("function _sample_gpu("):]
    assert "gpu_sample_init(model, de, n_iter)" in body and not re.search(r"[^_]sample_init\(model, de, n_iter\)", body.replace("gpu_sample_init", "GSI"))
