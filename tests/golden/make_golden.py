#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

Run in the BUILD container only (needs /root/reference, which does not exist on the GPU
box):  python tests/golden/make_golden.py

Outputs
  known_answers.json      known-answer vectors transcribed from the reference's own tests
                          (each entry cites file:line); SURVEY.md Appendix B
  manhattan_g2o.npz       parsed EDGE_SE2 records of examples/manhattan.g2o (ids, mu, upper-tri info)
  octagon_g2o.npz         same for test/octagon.g2o
  manhattan500_fixture.npz  subset of examples/manhattan-batch-500-fg.tar.gz: solved Pose2 particles
                          (first 120 poses x 100) + every factor among them (mu, Sigma) + PPE means
  pose3_clouds.npz        test/X1ptst.csv, test/X2ptst.csv (Pose3 particle coordinates)

The reference cannot be executed here (no Julia), so the known answers are the numbers the
reference's tests ASSERT, not outputs of a run; the data fixtures are parsed reference data.
"""
import io
import json
import math
import os
import re
import tarfile

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
pi = math.pi


def known_answers():
    ka = {"pose2pose2": [], "bearingrange": [], "pose3pose3": [], "pose2pose2_parametric": []}
    P = ka["pose2pose2"]
    src = "test/testParametricSimulated.jl"
    P.append(dict(src=src + ":37-40", X=[0, 0, -pi], p=[0, 0, 0], q=[0, 0, 0], expect_abs=[0, 0, pi], atol=1e-12))
    P.append(dict(src=src + ":42-43", X=[0, 0, -pi], p=[0, 0, 0], q=[0, 0, -pi], expect=[0, 0, 0], atol=1e-14))
    P.append(dict(src=src + ":45-46", X=[0, 0, -pi], p=[0, 0, 0], q=[0, 0, pi], expect=[0, 0, 0], atol=1e-14))
    meas = [10.0, 0.0, 1.0471975511965976]
    X1 = [15.000000000016204, 8.660254037814505, 2.0943951023931953]
    P.append(dict(src=src + ":105-130", X=meas, p=X1, q=[10.00004891350537, 17.320479835550103, 4.498439149584132e-6],
                  expect_abs=[0, 0, pi], atol=1e-4))
    P.append(dict(src=src + ":133-137", X=meas, p=X1, q=[10.00004891350537, 17.320479835550103, pi],
                  expect=[0, 0, 0], atol=1e-4))
    P.append(dict(src=src + ":140-144", X=meas, p=X1, q=[10.00004891350537, 17.320479835550103, -pi],
                  expect=[0, 0, 0], atol=1e-4))
    # optimizer trace recorded in the comments: residual value at a given q
    P.append(dict(src=src + ":118-119 (recorded trace)", X=meas, p=X1,
                  q=[9.999965307450084, 17.32052810543953, -3.864597114787283e-6],
                  expect=[3.469264352442895e-5, -2.002964655985515e-5, -3.1415887889926792], atol=1e-9))
    P.append(dict(src=src + ":120-121 (recorded trace)", X=meas, p=X1,
                  q=[9.999965307450083, 17.32052810543953, -3.864597114889992e-6],
                  expect=[3.469264352798169e-5, -2.0029646563407863e-5, -3.141588788992679], atol=1e-9))

    B = ka["bearingrange"]
    src = "test/testBearingRange2D.jl"
    B.append(dict(src=src + ":57-66", meas=[0, 20.0], p=[0, 0, 0], l=[20.0, 0], expect=[0, 0], atol=1e-14))
    B.append(dict(src=src + ":70-82", meas=[pi / 2, 20.0], p=[0, 0, 0], l=[0, 20.0], expect=[0, 0], atol=1e-14))
    B.append(dict(src=src + ":86-96", meas=[0, 20.0], p=[0, 0, pi / 2], l=[0, 20.0], expect=[0, 0], atol=1e-14))
    B.append(dict(src=src + ":100-116", meas=[pi / 2, 20.0], p=[0, 0, -pi / 2], l=[20.0, 0], expect=[0, 0], atol=1e-14))
    x1, x2 = [0, 0, 0], [0, 0, pi / 2]
    for (b, la, lb, ln) in [(0.0, [10, 0], [0, 10], "122-135"), (pi / 2, [0, 10], [-10, 0], "137-150"),
                            (pi, [-10, 0], [0, -10], "152-165"), (-pi / 2, [0, -10], [10, 0], "167-180")]:
        B.append(dict(src=src + ":" + ln, meas=[b, 10.0], p=x1, l=la, expect=[0, 0], atol=1e-9))
        B.append(dict(src=src + ":" + ln, meas=[b, 10.0], p=x2, l=lb, expect=[0, 0], atol=1e-9))
    B.append(dict(src=src + ":186-198", meas=[0, 10.0], p=x1, l=[11, 0], expect=[0, -1], atol=1e-9))
    B.append(dict(src=src + ":186-198", meas=[0, 10.0], p=x2, l=[0, 11], expect=[0, -1], atol=1e-9))
    B.append(dict(src=src + ":200-213", meas=[0, 10.0], p=x1, l=[9, 0], expect=[0, 1], atol=1e-9))
    B.append(dict(src=src + ":200-213", meas=[0, 10.0], p=x2, l=[0, 9], expect=[0, 1], atol=1e-9))
    s, c = 10 * math.sin(0.001), 10 * math.cos(0.001)
    B.append(dict(src=src + ":216-226", meas=[0, 10.0], p=x1, l=[c, s], expect=[-0.001, 0], atol=1e-9))
    B.append(dict(src=src + ":228-233", meas=[0, 10.0], p=x2, l=[s, c], expect=[0.001, 0], atol=1e-9))
    r2 = 10 / math.sqrt(2)
    B.append(dict(src=src + ":238-245", meas=[0, 10.0], p=x1, l=[r2, r2], expect=[-pi / 4, 0], atol=1e-9))
    B.append(dict(src=src + ":247-250", meas=[0, 10.0], p=x2, l=[r2, r2], expect=[pi / 4, 0], atol=1e-9))

    # Pose3Pose3: p = identity, q = (xyz, RotXYZ(rpy)); X = log of the relative pose -> residual 0.
    # Rotations.RotXYZ(a,b,c) = Rx(a) Ry(b) Rz(c); one angle non-zero per row so it is Exp(angle * axis).
    T = ka["pose3pose3"]
    src = "test/testPartialPose3.jl:398-436"
    a = 0.1
    rows = [[10., 0, 0, 0, 0, 0], [0., 10, 0, 0, 0, 0], [0., 0, 10, 0, 0, 0], [10., 0, 0, a, 0, 0],
            [0., 10, 0, a, 0, 0], [0., 0, 10, a, 0, 0], [10., 0, 0, 0, a, 0], [0., 10, 0, 0, a, 0],
            [0., 0, 10, 0, a, 0], [0., 15, 10, 0, a, 0], [10., 0, 0, 0, 0, a], [0., 10, 0, 0, 0, a],
            [0., 0, 10, 0, 0, a], [0., 15, 10, 0, 0, a]]
    for r in rows:
        T.append(dict(src=src, X=r, p=[0] * 6, q=r, expect_norm_below=1e-10))
    src = "test/threeDimLinearProductTest.jl"
    T.append(dict(src=src + ":150-156", X=[10., 0, 0, 0, 0, 0], p=[0] * 6, q=[10., 0, 0, 0, 0, 0], expect_norm_below=1e-10))
    T.append(dict(src=src + ":162-167", X=[10., 0, 0, pi, pi, pi], p=[0] * 6, q=[10., 0, 0, pi, pi, pi],
                  expect_norm_below=1e-10))

    # Pose2Point2Bearing (next row N1): test/testBearing2D.jl:13-47 (q = (5,5), measurement pi/4) and :59-66
    Bg = ka.setdefault("pose2point2bearing", [])
    ps = [[0., 0, 0], [5., 0, 0], [10., 0, 0], [10., 5, 0], [10., 10, 0], [5., 10, 0], [0., 10, 0], [0., 5, 0],
          [0., 0, pi / 4], [0., 0, -pi / 4], [1., 2, 0]]
    rs = [0, -pi / 4, -pi / 2, -3 * pi / 4, -pi, 3 * pi / 4, pi / 2, pi / 4, pi / 4, -pi / 4, pi / 4 - math.atan2(3, 4)]
    for pp, rr in zip(ps, rs):
        Bg.append(dict(src="test/testBearing2D.jl:13-47", b=pi / 4, p=pp, l=[5., 5.], expect=[rr], atol=1e-3,
                       modulo_2pi=True))
    Bg.append(dict(src="test/testBearing2D.jl:59-66", b=pi, p=[0., 0, 0], l=[-1., -0.001], expect=[-0.001], atol=1e-3))

    # ---- next-row 3-D families: test/testPartialPose3.jl ------------------------------------
    # poses are given there as (t, R); coordinates = (t, rotation vector of R) computed with SciPy here
    from scipy.spatial.transform import Rotation as Rot

    def coords(t, R):
        return [float(v) for v in t] + [float(v) for v in Rot.from_matrix(np.asarray(R, dtype=float)).as_rotvec()]

    def RZ(a): return Rot.from_euler("z", a).as_matrix()
    def RY(a): return Rot.from_euler("y", a).as_matrix()
    def RX(a): return Rot.from_euler("x", a).as_matrix()
    Y = ka.setdefault("pose3pose3xyyaw", [])
    src = "test/testPartialPose3.jl"
    mu2 = [20.0, 5.0, pi / 2]  # :91-92, evaluated at the belief mean (the reference samples: atol 0.15)
    cases = [
        (":172-177", ([0, 0, 0.], RZ(0)), ([20, 5, 0.], RZ(pi / 2))),
        (":179-183", ([0, 0, 0.], RZ(pi / 2)), ([-5, 20, 0.], np.diag([-1., -1, 1]))),
        (":185-189", ([0, 0, 100.], RZ(pi / 2)), ([-5, 20, -100.], np.diag([-1., -1, 1]))),
        (":191-196", ([0, 0, 0.], RY(pi / 4)), ([20, 5, 0.], RZ(pi / 2) @ RY(pi / 4))),
        (":198-202", ([0, 0, 10.], RY(pi / 4)), ([20, 5, -10.], RZ(pi / 2) @ RY(pi / 4))),
        (":204-208", ([0, 0, 0.], RX(pi / 4)), ([20, 5, 0.], RZ(pi / 2) @ RX(pi / 4))),
        (":210-214", ([0, 0, 10.], RX(pi / 4)), ([20, 5, -10.], RZ(pi / 2) @ RX(pi / 4))),
        (":216-220", ([10, 0, 0.], RX(pi / 4)), ([30, 5, 0.], RZ(pi / 2) @ RX(pi / 4))),
        (":222-226", ([0, 0, 0.], RY(pi / 4) @ RX(pi / 6)), ([20, 5, 0.], RZ(pi / 2) @ RY(pi / 4) @ RX(pi / 6))),
        (":228-232", ([10, 0, 10.], RY(pi / 4) @ RX(pi / 6)), ([30, 5, -10.], RZ(pi / 2) @ RY(pi / 4) @ RX(pi / 6))),
    ]
    for ln, (tp, Rp), (tq, Rq) in cases:
        Y.append(dict(src=src + ln, X=mu2, p=coords(tp, Rp), q=coords(tq, Rq), expect=[0, 0, 0], atol=0.15,
                      exact_atol=1e-9))
    # sign / frame checks :272-285 (RotZYX(0, -pi/4, pi/4) = Rz(0) Ry(-pi/4) Rx(pi/4))
    Rj = RZ(0.0) @ RY(-pi / 4) @ RX(pi / 4)
    Y.append(dict(src=src + ":272-277", X=[20.0, 5.0, 0.0], p=coords([10, 0, 10.], RZ(0.0)), q=coords([20, 10, -10.], Rj),
                  expect=[10, -5, 0], atol=0.15, exact_atol=1e-9))
    Y.append(dict(src=src + ":280-285", X=[20.0, 5.0, pi / 4], p=coords([10, 0, 10.], RZ(pi / 2)),
                  q=coords([20, 10, -10.], Rj), expect=[-15, 10, 3 * pi / 4], atol=0.15, exact_atol=1e-9))
    # Pose3Pose3Rotation :504-548: p = identity, q = (0, RotXYZ(rpy)) with one non-zero angle, measurement = rpy
    Rr = ka.setdefault("pose3pose3rotation", [])
    for rpy in [[0, 0, 0], [0.1, 0, 0], [0, 0.1, 0], [0, 0, 0.1], [-0.1, 0, 0], [0, -0.1, 0], [0, 0, -0.1]]:
        Rq = RX(rpy[0]) @ RY(rpy[1]) @ RZ(rpy[2])  # Rotations.RotXYZ
        Rr.append(dict(src=src + ":504-548", m=[float(v) for v in rpy], p=[0.0] * 6, q=coords([0, 0, 0.], Rq),
                       expect_norm_below=1e-10))

    # parametric square loop: the posterior means the reference asserts are exact roots of the
    # chain x_{k+1} = x_k o Exp(m) (pins the Hybrid exp: translation NOT coupled through V(theta))
    ka["pose2pose2_parametric"].append(dict(
        src="test/testParametric.jl:22-53", prior=[10, 10, -pi + 1e-5], X=[10, 0, pi / 2],
        chain=[[0, 10, -pi / 2], [0, 0, 0], [10, 0, pi / 2], [10, 10, -pi]], atol=1e-3))
    # hexagon ground truth of generateGraph_Hexagonal: test/testHexagonal2D_CliqByCliq.jl:37-79
    ka["hexagon_truth"] = dict(
        src="test/testHexagonal2D_CliqByCliq.jl:37-79; src/canonical/GenerateCircular.jl:57-90",
        X=[10.0, 0.0, pi / 3],
        poses=[[0, 0, 0], [10, 0, pi / 3], [15, 8.66, 2 * pi / 3], [10, 17.32, pi], [0, 17.32, -2 * pi / 3],
               [-5, 8.66, -pi / 3], [0, 0, 0]], landmark=[20, 0], atol=0.01)
    # data association + extra legs that generateGraph_Honeycomb! forces (parsed, not transcribed)
    txt = open(os.path.join(REF, "src/canonical/GenerateHoneycomb.jl")).read()
    ka["honeycomb_recipe"] = dict(
        src="src/canonical/GenerateHoneycomb.jl:3-52",
        landmarks=dict(re.findall(r":(l\d+)\s*=>\s*:(l\d+)", txt)),
        legs={k: v for k, v in re.findall(r":(x\d+)\s*=>\s*:(left|right)", txt)})
    with open(os.path.join(OUT, "known_answers.json"), "w") as fh:
        json.dump(ka, fh, indent=1)
    return ka


def parse_g2o_se2(path):
    """EDGE_SE2 reader following src/services/g2oParser.jl:39-49,91-105 (whitespace split)."""
    ids, mu, info = [], [], []
    with open(path) as fh:
        for ln in fh:
            t = ln.split()
            if t and t[0] == "EDGE_SE2":
                ids.append([int(t[1]), int(t[2])])
                mu.append([float(v) for v in t[3:6]])
                info.append([float(v) for v in t[6:12]])
    return np.array(ids, np.int32), np.array(mu), np.array(info)


def parse_fullnormal(txt):
    mu = [float(v) for v in re.search(r"μ: \[([^\]]*)\]", txt).group(1).replace(",", " ").split()]
    sg = re.search(r"Σ: \[([^\]]*)\]", txt).group(1)
    S = [[float(v) for v in row.split()] for row in sg.split(";")]
    return np.array(mu), np.array(S)


def manhattan500(npose=120):
    tf = tarfile.open(os.path.join(REF, "examples/manhattan-batch-500-fg.tar.gz"))
    variables, factors = {}, []
    for m in tf.getmembers():
        if not m.name.endswith(".json"):
            continue
        d = json.load(io.TextIOWrapper(tf.extractfile(m), encoding="utf-8"))
        if "/variables/" in m.name:
            sd = json.loads(d["solverDataDict"])["default"]
            ppe = json.loads(d["ppeDict"])["default"]
            variables[d["label"]] = (np.array(sd["vecval"], dtype=np.float64).reshape(-1, sd["dimval"]),
                                     np.array(ppe["mean"], dtype=np.float64))
        elif "/factors/" in m.name:
            vo = d["_variableOrderSymbols"]
            vo = json.loads(vo) if isinstance(vo, str) else vo
            data = json.loads(d["data"]) if isinstance(d["data"], str) else d["data"]
            fnc = data["fnc"]
            txt = fnc.get("datastr", fnc.get("str"))
            mu, S = parse_fullnormal(txt)
            factors.append((d["fnctype"], vo, mu, S))
    keep = ["x%d" % i for i in range(npose)]
    assert all(k in variables for k in keep)
    idx = {k: i for i, k in enumerate(keep)}
    parts = np.stack([variables[k][0] for k in keep])  # [npose][100][3]
    ppe = np.stack([variables[k][1] for k in keep])
    ip, iq, mu, Sg = [], [], [], []
    prior = None
    for typ, vo, m, S in factors:
        if typ == "Pose2Pose2" and all(v in idx for v in vo):
            ip.append(idx[vo[0]]); iq.append(idx[vo[1]]); mu.append(m); Sg.append(S)
        elif typ == "PriorPose2":
            prior = (idx[vo[0]], m, S)
    np.savez_compressed(os.path.join(OUT, "manhattan500_fixture.npz"), particles=parts, ppe_mean=ppe,
                        ip=np.array(ip, np.int32), iq=np.array(iq, np.int32), mu=np.array(mu), Sigma=np.array(Sg),
                        prior_var=np.int32(prior[0]), prior_mu=prior[1], prior_Sigma=prior[2],
                        n_variables_total=np.int32(len(variables)), n_factors_total=np.int32(len(factors)))
    return len(ip), parts.shape


def main():
    ka = known_answers()
    ids, mu, info = parse_g2o_se2(os.path.join(REF, "examples/manhattan.g2o"))
    np.savez_compressed(os.path.join(OUT, "manhattan_g2o.npz"), ids=ids, mu=mu, info=info)
    i2, m2, f2 = parse_g2o_se2(os.path.join(REF, "test/octagon.g2o"))
    np.savez_compressed(os.path.join(OUT, "octagon_g2o.npz"), ids=i2, mu=m2, info=f2)
    nf, shp = manhattan500()
    x1 = np.loadtxt(os.path.join(REF, "test/X1ptst.csv"), delimiter=",")
    x2 = np.loadtxt(os.path.join(REF, "test/X2ptst.csv"), delimiter=",")
    np.savez_compressed(os.path.join(OUT, "pose3_clouds.npz"), X1=x1, X2=x2)
    print("known answers:", {k: (len(v) if isinstance(v, list) else 1) for k, v in ka.items()})
    print("manhattan edges", ids.shape, "octagon", i2.shape, "m500 factors", nf, shp, "clouds", x1.shape, x2.shape)


if __name__ == "__main__":
    main()
