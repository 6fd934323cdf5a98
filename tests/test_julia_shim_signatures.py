"""The Julia shim (gpr.jl_b200/julia/GPRB200.jl) cannot be executed in this image (no Julia toolchain), so its one
machine-checkable property is verified here: every ``ccall`` names a symbol that include/gprb200.h declares, with the
same number of arguments, C-compatible argument types in the same order, and the same return type; the two structs that
cross the boundary by reference (gprb_lbfgs_opts, gprb_opt_result) have the header's fields in the header's order.
A maintainer's drop-in also must not shadow GaussianProcesses.jl / Optim.jl exports."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gprb200.h")
SHIM = os.path.join(ROOT, "gpr.jl_b200", "julia", "GPRB200.jl")

# Julia ccall type -> set of acceptable C parameter types (normalised: no const, no parameter name, no spaces)
JL2C = {
    "Cint": {"int", "int32_t"}, "Int32": {"int32_t", "int"}, "Int64": {"int64_t"}, "Cstring": {"constchar*", "char*"},
    "Ptr{Cvoid}": {"gprb_ctx*", "gprb_batch*", "gprb_dataset*", "void*", "gprb_ctx*const*"},
    "Ref{Ptr{Cvoid}}": {"gprb_ctx**", "gprb_batch**", "gprb_dataset**"},
    "Ptr{Ptr{Cvoid}}": {"gprb_dataset*const*", "gprb_dataset**", "gprb_ctx**", "gprb_ctx*const*"},
    "Ptr{Float64}": {"double*"}, "Ptr{Ptr{Float64}}": {"double*const*"},
    "Ptr{UInt8}": {"uint8_t*"}, "Ptr{Int32}": {"int32_t*"},
    "Ref{LbfgsOpts}": {"gprb_lbfgs_opts*"}, "Ptr{OptResult}": {"gprb_opt_result*"},
}
JL_FIELD = {"Int32": "int32_t", "Float64": "double"}


def _strip_comments(src):
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def header_prototypes():
    src = _strip_comments(open(HEADER).read())
    protos = {}
    for m in re.finditer(r"\b([A-Za-z_0-9 \*]+?)\b\s*(gprb_[a-z_A-Z0-9]+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        ret = ret.replace("const ", "").strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = re.sub(r"\[\d*\]", "*", a.strip())            # int64_t out[4] -> pointer
                a = a.replace("const ", "").strip()
                mm = re.match(r"^(.*?[\*\s])\s*([A-Za-z_][A-Za-z_0-9]*)?$", a)
                typ = (mm.group(1) if mm else a).replace(" ", "")
                if a.endswith("*") or (mm and mm.group(2) is None):
                    typ = a.replace(" ", "")
                params.append(typ)
        protos[name] = (ret.replace(" ", ""), params)
    return protos


def header_struct(name):
    src = _strip_comments(open(HEADER).read())
    body = re.search(r"typedef struct " + name + r"\s*\{(.*?)\}\s*" + name + r"\s*;", src, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        typ, names = decl.split(None, 1)
        for nm in names.split(","):
            fields.append((typ, nm.strip()))
    return fields


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def shim_ccalls():
    src = open(SHIM).read()
    src = re.sub(r"#.*", "", src)
    calls = []
    for m in re.finditer(r"ccall\(\(:(gprb_[a-z_A-Z0-9]+),\s*LIB\),\s*([A-Za-z0-9_{}]+),\s*\(", src):
        name, ret = m.group(1), m.group(2)
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        types = _split_top(src[m.end():i - 1])
        # the values follow the type tuple: count them up to the ccall's closing parenthesis
        j, depth = i, 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        values = _split_top(src[i:j - 1].lstrip(", \n"))
        calls.append((name, ret, types, values))
    return calls


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    calls = shim_ccalls()
    assert len(calls) >= 14
    for name, ret, types, values in calls:
        assert name in protos, f"{name} is not declared in include/gprb200.h"
        cret, cparams = protos[name]
        assert cret in JL2C[ret], (name, ret, cret)
        assert len(types) == len(cparams), (name, types, cparams)
        assert len(values) == len(types), (name, "value count", values)
        for k, (jt, ct) in enumerate(zip(types, cparams)):
            assert jt in JL2C, (name, jt)
            assert ct in JL2C[jt], f"{name} arg {k}: Julia {jt} vs C {ct}"


def test_the_reference_facing_entry_points_are_bound():
    bound = {c[0] for c in shim_ccalls()}
    for sym in ["gprb_init", "gprb_datasets_create", "gprb_batch_create", "gprb_eval", "gprb_optimize", "gprb_predict",
                "gprb_predict_async", "gprb_predict_wait", "gprb_get_alpha", "gprb_gather", "gprb_comm_unique_id",
                "gprb_comm_init_rank", "gprb_batch_destroy", "gprb_dataset_destroy", "gprb_last_error"]:
        assert sym in bound, sym


def test_structs_match_field_for_field():
    src = open(SHIM).read()
    for jl, c in (("LbfgsOpts", "gprb_lbfgs_opts"), ("OptResult", "gprb_opt_result")):
        body = re.search(r"struct " + jl + r"\b(.*?)\nend", src, flags=re.S).group(1)
        body = re.sub(r"#.*", "", body)
        jfields = [(t.strip(), n.strip()) for n, t in re.findall(r"([A-Za-z_0-9]+)::([A-Za-z0-9]+)", body)]
        cfields = header_struct(c)
        assert [n for _, n in jfields] == [n for _, n in cfields], (jl, jfields, cfields)
        for (jt, n), (ct, _) in zip(jfields, cfields):
            assert JL_FIELD[jt] == ct, (jl, n, jt, ct)


def test_no_name_clashes_with_gaussianprocesses_or_optim():
    src = open(SHIM).read()
    exported = set(re.search(r"^export (.*)$", src, flags=re.M).group(1).replace(" ", "").split(","))
    clash = exported & {"GP", "GPE", "predict_y", "optimize!", "update_mll!", "update_mll_and_dmll!", "SEArd", "MeanZero", "LBFGS",
                        "optimize", "fit!", "Mean", "Kernel"}
    assert not clash, clash
    assert "B200Covariance<:CovarianceStrategy" in src.replace(" ", "")
    # the reference's typed call sites keep working: the objects stay GaussianProcesses.GPE
    assert re.search(r"GaussianProcesses\.optimize!\(gps::Vector\{<:GPE\}", src)
    assert re.search(r"GaussianProcesses\.predict_y\(gp::B200GPE", src)
