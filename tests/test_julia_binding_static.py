"""Static consistency of the Julia binding (rome.jl_b200/julia/RoMEB200.jl, and the snippet in INTEGRATION.md) with the C
header: `julia` is not in the image, so the shim cannot be executed; what can be checked is that every `ccall` names an
exported symbol and passes the argument list the header declares (arity and C type class of every argument and of the
return value), that `struct Buffers` mirrors `rome_b200_buffers` field for field, and that the enum / flag constants
carry the header's values."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "rome_b200.h")).read()
JL = open(os.path.join(ROOT, "rome.jl_b200", "julia", "RoMEB200.jl")).read()
MD = open(os.path.join(ROOT, "INTEGRATION.md")).read()


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


def _c_class(t):
    t = re.sub(r"\bconst\b", "", t).strip()
    if "*" in t or "[" in t:
        return "ptr"
    t = t.split()[0] if t.split() else t
    return {"int": "i32", "uint32_t": "u32", "uint64_t": "u64", "size_t": "usize", "void": "void"}[t]


def header_signatures():
    sigs = {}
    for m in re.finditer(r"ROME_B200_API\s+([^;]*?)\b(rome_b200_\w+)\s*\(([^;]*?)\)\s*;", HDR, re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        params = [] if args == "void" else _split_top(args)
        classes = []
        for p in params:
            ptr = "*" in p or "[" in p
            base = re.sub(r"\b\w+\s*(\[\d*\])?$", "", p).strip() if not ptr else p  # drop the parameter name
            classes.append("ptr" if ptr else _c_class(base))
        sigs[name] = ("ptr" if "*" in ret else _c_class(ret), classes)
    return sigs


def _jl_class(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return {"Cint": "i32", "Int32": "i32", "UInt32": "u32", "Cuint": "u32", "UInt64": "u64", "Csize_t": "usize",
            "Cvoid": "void"}[t]


def julia_ccalls(text):
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*LIB\),\s*([\w{}]+),\s*\(", text):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        types = [t for t in _split_top(text[m.end():i - 1]) if t]
        # the values passed after the type tuple, up to the ccall's closing parenthesis
        j, depth = i, 1
        while depth:
            depth += {"(": 1, ")": -1}.get(text[j], 0)
            j += 1
        nvals = len([v for v in _split_top(re.sub(r"#=.*?=#", "", text[i:j - 1]).lstrip(", \n")) if v])
        calls.append((m.group(1), m.group(2), types, nvals))
    return calls


def test_every_ccall_matches_the_header():
    sigs = header_signatures()
    assert len(sigs) >= 45 and "rome_b200_eval_host" in sigs
    seen = set()
    for src, text in (("RoMEB200.jl", JL), ("INTEGRATION.md", MD)):
        calls = julia_ccalls(text)
        assert calls, src
        for name, ret, types, nvals in calls:
            assert name in sigs, f"{src}: {name} is not declared in include/rome_b200.h"
            cret, cargs = sigs[name]
            assert _jl_class(ret) == cret, f"{src}: {name} returns {ret}, header says {cret}"
            assert [_jl_class(t) for t in types] == cargs, f"{src}: {name} passes {types}, header expects {cargs}"
            assert nvals == len(types), f"{src}: {name} passes {nvals} values for {len(types)} declared argument types"
            seen.add(name)
    for must in ("rome_b200_create", "rome_b200_destroy", "rome_b200_last_error", "rome_b200_set_particles", "rome_b200_get_particles",
                 "rome_b200_set_factors_pose2pose2", "rome_b200_set_factors_priorpose2", "rome_b200_set_factors_bearingrange",
                 "rome_b200_set_factors_pose3pose3", "rome_b200_set_factors_gaussian", "rome_b200_eval_host",
                 "rome_b200_set_product_plan", "rome_b200_product"):
        assert must in seen, must


def test_buffers_struct_and_constants_mirror_the_header():
    cfields = re.findall(r"(?:const\s+)?float\*\s*(\w+);", re.search(r"typedef struct rome_b200_buffers \{(.*?)\}", HDR, re.S).group(1))
    for text in (JL, MD):
        body = re.search(r"struct Buffers(.*?)\bend", text, re.S).group(1)
        assert re.findall(r"(\w+)::Ptr\{Cfloat\}", body) == cfields
    enums = {k: int(v) for k, v in re.findall(r"ROME_B200_(\w+)\s*=\s*(-?\d+)", HDR)}
    flags = {k: int(v) for k, v in re.findall(r"#define ROME_B200_(\w+)\s+(\d+)u", HDR)}
    for m in re.finditer(r"const ([\w,\s]+?)\s*=\s*\n?\s*((?:(?:Cint|UInt32)\(\d+\),?\s*)+)", JL):
        names = [n.strip() for n in m.group(1).split(",")]
        vals = [int(v) for v in re.findall(r"\((\d+)\)", m.group(2))]
        assert len(names) == len(vals)
        for n, v in zip(names, vals):
            assert {**enums, **flags}[n] == v, n
    checked = set(n.strip() for m in re.finditer(r"const ([\w,\s]+?)\s*=\s*\n?\s*(?:Cint|UInt32)\(", JL) for n in m.group(1).split(","))
    assert {"POSE2", "POSE3POSE3", "POSE3POSE3UNITTRANS", "RESIDUAL", "SAMPLE", "DECONV", "PRODUCT_REANCHOR"} <= checked
    # the flag word quoted in INTEGRATION.md
    assert "0x1b #=RESIDUAL|PROPOSAL_FWD|STATS|SAMPLE=#" in MD and (flags["RESIDUAL"] | flags["PROPOSAL_FWD"] | flags["STATS"] | flags["SAMPLE"]) == 0x1b


def test_julia_shim_covers_the_five_hot_families():
    """the batched drop-in (`DeviceGraph` + `approxConvBatch`) exists for all five hot families, and the per-family
    dimensions it hands to the library are the library's own (rome_b200_family_dims / engine.FAMILY)"""
    import rome_b200 as rb
    fam_of = re.search(r"const FAMILY_OF = Dict\{DataType,Cint\}\((.*?)\)\n", JL, re.S).group(1)
    pairs = dict(re.findall(r"(\w+) => (\w+)", fam_of))
    assert pairs == {"Pose2Pose2": "POSE2POSE2", "PriorPose2": "PRIORPOSE2", "Pose2Point2BearingRange": "BEARINGRANGE",
                     "Pose3Pose3": "POSE3POSE3", "PriorPose3": "PRIORPOSE3"}
    dims = re.search(r"const FAMILY_DIMS = Dict\{Cint,Tuple\}\((.*?)\)\nconst", JL, re.S).group(1)
    got = {m[0]: (m[1], m[2], int(m[3]), int(m[4]), int(m[5]), int(m[6]))
           for m in re.findall(r"(\w+) => \((\w+), (\w+), (\d+), (\d+), (\d+), (\d+)\)", dims)}
    vt = {rb.POSE2: "Pose2", rb.POINT2: "Point2", rb.POSE3: "Pose3", None: "nothing"}
    for name, fam in (("POSE2POSE2", rb.POSE2POSE2), ("PRIORPOSE2", rb.PRIORPOSE2), ("BEARINGRANGE", rb.BEARINGRANGE),
                      ("POSE3POSE3", rb.POSE3POSE3), ("PRIORPOSE3", rb.PRIORPOSE3)):
        v0, v1, dm, dr, ns, _, dfwd, _ = rb.FAMILY[fam]
        assert got[name] == (vt[v0], vt[v1], dm, dr, dfwd, ns), name
    assert "function approxConvBatch(dg::DeviceGraph, ::Type{F}; seed::UInt64=UInt64(0)) where {F}" in JL
    assert "function upload_factors!(dg::DeviceGraph)" in JL and "function upload_particles!(dg::DeviceGraph)" in JL
    assert "IncrementalInference.approxConvBelief(dfg::AbstractDFG, fc::DFGFactor{<:CommonConvWrapper{<:F}}" in JL
    # every flag constant of the header is mirrored (the regex of the test above walks them; the newest ones by name)
    for name in ("PRECISE", "ROUTED_ONLY", "BARRIER_WAIT", "BARRIER_SIGNAL"):
        assert re.search(rf"\b{name}\b", JL), name
