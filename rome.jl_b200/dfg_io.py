"""DFG file ingest / export for the hot path's graphs: `loadDFG` / `saveDFG`.

Reads the two on-disk layouts the reference ships, straight into the host graph whose arrays feed the device tables:

* DFG v0.25 (`test/testdata/g2otest.tar.gz`, loaded by `test/testG2oExportSE3.jl:21`): `dfg.json` (with `solverParams` and
  `addHistory`), `variables/<label>.json` (`variableType`, `solverData` = list of {`solveKey`, `vecval`, `dimval`,
  `initialized`}, `ppes`), `factors/<label>.json` (`fnctype`, `_variableOrderSymbols`, `data` = JSON string whose `fnc`
  holds the Packed* factor: `Z = {"_type": "IncrementalInference.PackedFullNormal", "mu", "cov"}`, or `bearstr` / `rangstr`).
* legacy 2020 layout (`examples/manhattan-batch-500-fg.tar.gz`): `<root>/{variables,factors}/<label>.json`, variables with
  `solverDataDict` / `ppeDict` JSON strings keyed by solve key (`softtype` names the variable type), factors whose `fnc`
  holds the belief as text: `FullNormal(\\ndim: 3\\nμ: [..]\\nΣ: [a b c; d e f; g h i]\\n)` under `datastr` / `str`.

`saveDFG` writes the v0.25 layout (tar.gz).  Only the factor families of this path are understood; anything else raises.
"""
from __future__ import annotations

import datetime
import io
import json
import os
import re
import tarfile

import numpy as np

from .factors import (MvNormal, Normal, Point2, Point3, Pose2, Pose2Point2BearingRange, Pose3, pack, unpack)
from .g2o import _natural_key
from .graph import FactorGraph, SolverParams, addFactor, addVariable, initfg

_VARTYPES = {"Pose2": Pose2, "Point2": Point2, "Pose3": Pose3, "Point3": Point3}
DFG_VERSION = "0.25.1"


# ---- beliefs as the legacy text forms -------------------------------------------------------------------------------
def _parse_fullnormal_text(txt: str) -> MvNormal:
    """`FullNormal(\\ndim: d\\nμ: [..]\\nΣ: [r1; r2; ..]\\n)` -- Distributions' `show` of an MvNormal (2020 files)"""
    m, s = re.search(r"μ: \[([^\]]*)\]", txt), re.search(r"Σ: \[([^\]]*)\]", txt)
    if not (m and s):
        raise ValueError(f"not a FullNormal text: {txt[:60]!r}")
    mu = np.array([float(v) for v in m.group(1).replace(",", " ").split()])
    Sigma = np.array([[float(v) for v in row.split()] for row in s.group(1).split(";")])
    return MvNormal(mu, Sigma)


def _parse_normal_text(txt: str) -> Normal:
    """`Normal{Float64}(μ=0.0, σ=0.03)` (also the positional `Normal(0.0, 0.03)`)"""
    nums = re.findall(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?", txt.split("(", 1)[1])
    return Normal(float(nums[0]), float(nums[1]))


def _belief(x):
    if isinstance(x, dict):
        if x["_type"].endswith("PackedNormal") or x["_type"].endswith("PackedFullNormal"):
            return x
        raise ValueError(f"unsupported belief type {x['_type']}")
    if isinstance(x, str):
        b = _parse_normal_text(x) if x.lstrip().startswith("Normal") else _parse_fullnormal_text(x)
        return ({"_type": "IncrementalInference.PackedNormal", "mu": b.mu, "sigma": b.sigma} if isinstance(b, Normal) else
                {"_type": "IncrementalInference.PackedFullNormal", "mu": b.mu.tolist(), "cov": b.Sigma.reshape(-1).tolist()})
    raise ValueError("belief must be a Packed* dictionary or its legacy text")


def _factor_from_json(d: dict):
    name = d["fnctype"].split(".")[-1]
    data = json.loads(d["data"]) if isinstance(d["data"], str) else d["data"]
    fnc = data["fnc"]
    if name == "Pose2Point2BearingRange":
        packed = {"_type": "Packed" + name, "bearstr": _belief(fnc["bearstr"]), "rangstr": _belief(fnc["rangstr"])}
    else:
        z = fnc["Z"] if "Z" in fnc else fnc.get("datastr", fnc.get("str"))
        if z is None:
            raise ValueError(f"factor {d['label']}: no belief field in {sorted(fnc)}")
        packed = {"_type": "Packed" + name, "Z": _belief(z)}
    try:
        return unpack(packed), data
    except KeyError as e:
        raise ValueError(f"factor {d['label']}: family {name} is outside this path") from e


def _jsonfield(v, default):
    if v is None:
        return default
    return json.loads(v) if isinstance(v, str) else v


# ---- reading -----------------------------------------------------------------------------------------------------
def _read_members(path: str) -> dict[str, dict]:
    """{relative path: parsed JSON} of a .tar.gz or an unpacked directory"""
    out = {}
    if os.path.isdir(path):
        for root, _, files in os.walk(path):
            for f in files:
                if f.endswith(".json"):
                    p = os.path.join(root, f)
                    out[os.path.relpath(p, path)] = json.load(open(p, encoding="utf-8"))
        return out
    if not os.path.exists(path) and os.path.exists(path + ".tar.gz"):
        path += ".tar.gz"  # loadDFG appends the extension itself
    with tarfile.open(path) as tf:
        for m in tf.getmembers():
            if m.isfile() and m.name.endswith(".json"):
                out[m.name] = json.load(io.TextIOWrapper(tf.extractfile(m), encoding="utf-8"))
    return out


def loadDFG(path: str, fg: FactorGraph | None = None, solveKey: str = "default") -> FactorGraph:
    """loadDFG(path) / loadDFG!(fg, path).  Variables enter in `addHistory` order when the file records it, else in
    natural label order; factors in timestamp order (their creation order), so table rows and `exportG2o` output follow
    the saved graph.  A variable's particles (`val`, [N][d] coordinates) come from the `solveKey` solver data when it
    is marked initialized; a "parametric" entry lands in `.parametric`; PPEs in `.ppes`."""
    members = _read_members(path)
    variables = {k: v for k, v in members.items() if os.path.basename(os.path.dirname(k)) == "variables"}
    factors = {k: v for k, v in members.items() if os.path.basename(os.path.dirname(k)) == "factors"}
    top = next((v for k, v in members.items() if os.path.basename(k) == "dfg.json"), None)
    if fg is None:
        sp = SolverParams()
        if top and isinstance(top.get("solverParams"), dict):
            for k in sp.__dataclass_fields__:
                if k in top["solverParams"]:
                    setattr(sp, k, top["solverParams"][k])
        fg = initfg(sp)
    byl = {d["label"]: d for d in variables.values()}
    order = [l for l in (top or {}).get("addHistory", []) if l in byl]
    order += sorted((l for l in byl if l not in set(order)), key=_natural_key)
    for l in order:
        d = byl[l]
        if "solverData" in d:  # v0.25
            tname = d["variableType"].split(".")[-1]
            sds = {s["solveKey"]: s for s in d["solverData"]}
            ppes = {p["solveKey"]: p for p in _jsonfield(d.get("ppes"), [])}
        else:  # legacy
            sds = _jsonfield(d.get("solverDataDict"), {})
            tname = re.search(r"(\w+)\s*(?:\(|$)", next(iter(sds.values()))["softtype"]).group(1)  # `Pose2(3, String[], ...)`
            ppes = _jsonfield(d.get("ppeDict"), {})
        if tname not in _VARTYPES:
            raise ValueError(f"variable {l}: type {tname} is outside this path")
        v = addVariable(fg, l, _VARTYPES[tname], tags=[t for t in _jsonfield(d.get("tags"), []) if t != "VARIABLE"])
        for key, s in sds.items():
            pts = np.asarray(s["vecval"], dtype=np.float64).reshape(-1, int(s["dimval"]))
            if key == solveKey and s.get("initialized", False):
                v.val = pts
            elif key == "parametric" and key != solveKey:
                v.parametric = pts[0].copy()
        v.ppes = {k: {f: np.asarray(p[f], dtype=np.float64) for f in ("suggested", "max", "mean") if f in p}
                  for k, p in ppes.items()}
    for d in sorted(factors.values(), key=lambda d: (str(d.get("timestamp", "")), _natural_key(d["label"]))):
        fnc, data = _factor_from_json(d)
        f = addFactor(fg, _jsonfield(d["_variableOrderSymbols"], []), fnc, graphinit=False,
                      tags=[t for t in _jsonfield(d.get("tags"), []) if t != "FACTOR"])
        if f.label != d["label"]:  # keep the saved label
            del fg.factors[f.label]
            f.label = d["label"]
            fg.factors[f.label] = f
        mh = data.get("multihypo")
        if mh not in (None, "", []) or float(data.get("nullhypo") or 0.0) != 0.0:
            raise ValueError(f"factor {f.label}: multihypo / nullhypo are outside this path")
    return fg


# ---- writing (v0.25 layout) -----------------------------------------------------------------------------------------
def _solver_data(v, key, pts, initialized):
    d = v.variableType.dim
    return {"vecval": np.asarray(pts, dtype=np.float64).reshape(-1).tolist(), "dimval": d, "vecbw": [0.0] * d, "dimbw": d,
            "BayesNetOutVertIDs": [], "dimIDs": [], "dims": d, "eliminated": False, "BayesNetVertID": "NOTHING",
            "separator": [], "variableType": "RoME." + v.variableType.__name__, "initialized": bool(initialized),
            "infoPerCoord": [0.0] * d, "ismargin": False, "dontmargin": False, "solveInProgress": 0, "solvedCount": 0,
            "solveKey": key, "covar": [], "_version": DFG_VERSION}


def saveDFG(fg: FactorGraph, path: str) -> str:
    """saveDFG(fg, path): DFG v0.25 tar.gz (the extension is appended when missing, like the reference does)"""
    if not path.endswith(".tar.gz"):
        path += ".tar.gz"
    files = {}
    sp = fg.solverParams
    files["dfg.json"] = {"description": "", "addHistory": list(fg.variables),
                         "solverParams": {k: getattr(sp, k) for k in sp.__dataclass_fields__},
                         "solverParams_type": "SolverParams", "graphLabel": "factorgraph_rome_b200"}
    for l, v in fg.variables.items():
        sd = []
        par = getattr(v, "parametric", None)
        if par is not None:
            sd.append(_solver_data(v, "parametric", par, True))
        N = fg.solverParams.N
        sd.append(_solver_data(v, "default", v.val if v.val is not None else np.zeros((N, v.variableType.dim)),
                               v.val is not None))
        ppes = [dict(solveKey=k, **{f: np.asarray(x).tolist() for f, x in p.items()})
                for k, p in getattr(v, "ppes", {}).items()]
        files[f"variables/{l}.json"] = {"label": l, "tags": ["VARIABLE"] + list(v.tags), "nstime": "0", "ppes": ppes,
                                        "blobEntries": [], "variableType": "RoME." + v.variableType.__name__,
                                        "_version": DFG_VERSION, "metadata": "e30=", "solvable": 1, "solverData": sd}
    t0 = datetime.datetime(2024, 1, 1)  # creation order is all the time stamps carry: one millisecond per factor
    for k, (l, f) in enumerate(fg.factors.items()):
        p = pack(f.fnc)
        fnc = {kk: vv for kk, vv in p.items() if kk != "_type"}
        data = {"eliminated": False, "potentialused": False, "edgeIDs": [], "fnc": fnc, "multihypo": [],
                "certainhypo": list(range(1, len(f.variableOrderSymbols) + 1)), "nullhypo": 0.0, "solveInProgress": 0,
                "inflation": fg.solverParams.inflation}
        files[f"factors/{l}.json"] = {"label": l, "tags": ["FACTOR"] + list(f.tags),
                                      "_variableOrderSymbols": list(f.variableOrderSymbols),
                                      "timestamp": (t0 + datetime.timedelta(milliseconds=k)).isoformat(timespec="milliseconds")
                                      + "+00:00", "nstime": "0",
                                      "fnctype": type(f.fnc).__name__, "solvable": 1,
                                      "data": json.dumps(data, separators=(",", ":")), "metadata": "e30=",
                                      "_version": DFG_VERSION}
    with tarfile.open(path, "w:gz") as tf:
        for name, obj in files.items():
            raw = json.dumps(obj, separators=(",", ":")).encode("utf-8")
            ti = tarfile.TarInfo(name)
            ti.size = len(raw)
            tf.addfile(ti, io.BytesIO(raw))
    return path


def setPPE(fg: FactorGraph, labels=None, solveKey: str = "default"):
    """setPPE!.(fg, ls(fg), solveKey) (test/testG2oExportSE3.jl:23): point estimates from the stored solver data --
    `suggested` / `mean` = the on-manifold mean of the particles (circular for the Pose2 heading), `max` = the particle
    nearest to it; for "parametric" the parametric solution itself."""
    for l in (labels or list(fg.variables)):
        v = fg.variables[str(l)]
        if solveKey == "parametric" and getattr(v, "parametric", None) is not None:
            x = np.asarray(v.parametric, dtype=np.float64)
            est = {"suggested": x.copy(), "max": x.copy(), "mean": x.copy()}
        else:
            if v.val is None:
                raise ValueError(f"variable {l} has no solver data for {solveKey}")
            mean = v.val.mean(0)
            if v.variableType is Pose2:
                mean[2] = np.arctan2(np.sin(v.val[:, 2]).mean(), np.cos(v.val[:, 2]).mean())
            est = {"suggested": mean, "max": v.val[np.argmin(((v.val - mean) ** 2).sum(1))].copy(), "mean": v.val.mean(0)}
        if not hasattr(v, "ppes"):
            v.ppes = {}
        v.ppes[solveKey] = est
    return fg
