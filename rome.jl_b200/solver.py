"""On-device belief propagation sweeps: convolutions (the hot path) followed by the product of every variable's
proposal densities, with particles resident on the GPU between sweeps.

This is NOT the reference's Bayes-tree solve (`solveTree!`): there is no tree, no clique scheduling and no
message passing between cliques (all out of scope).  It is the inner operation IIF repeats inside every
clique -- for each variable, `propagateBelief` = product over its factors of `approxConvBelief` -- applied to ALL
variables of the graph at once (a synchronous sweep), `gibbsIters` times.  All arithmetic runs in librome_b200.so.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .engine import FAMILY, VAR_DIM, Context, npad
from .graph import DeviceGraph, FactorGraph


def build_product_plans(fg: FactorGraph, families=None):
    """Host logic: for every variable type the CSR product plan over proposal buffers.

    Buffer numbering: buffer 2*k is prop_fwd and 2*k+1 is prop_bwd of the k-th family (ascending family id) that has
    factors.  A factor contributes prop_fwd to its LAST variable (or to the prior's variable) when the family has a
    closed-form forward root, and prop_bwd to its FIRST variable when it has a backward one.
    Returns (plans, buffers): plans[vartype] = (var_offsets, src_buf, src_row); buffers = [(family, "fwd"|"bwd")]."""
    fams = sorted(f for f in FAMILY if any(x.fnc.family == f for x in fg.factors.values()))
    if families is not None:
        fams = [f for f in fams if f in families]
    buffers = []
    for f in fams:
        buffers += [(f, "fwd"), (f, "bwd")]
    per_var = {t: {} for t in VAR_DIM}
    for v in fg.variables.values():
        per_var[v.variableType.vartype][v.index] = []
    for x in sorted(fg.factors.values(), key=lambda x: (x.fnc.family, x.index)):
        fam = x.fnc.family
        if fam not in fams:
            continue
        vt0, vt1, _, _, _, _, dfwd, dbwd = FAMILY[fam]
        k = fams.index(fam)
        vs = [fg.variables[l] for l in x.variableOrderSymbols]
        if dfwd:
            tgt = vs[-1]
            per_var[tgt.variableType.vartype][tgt.index].append((2 * k, x.index))
        if dbwd and vt1 is not None:
            per_var[vs[0].variableType.vartype][vs[0].index].append((2 * k + 1, x.index))
    plans = {}
    for t, d in per_var.items():
        if not d:
            continue
        n = max(d) + 1
        off = np.zeros(n + 1, np.int32)
        sb, sr = [], []
        for i in range(n):
            src = d.get(i, [])
            if len(src) > L.MAX_PRODUCT_SOURCES:
                raise ValueError(f"variable {i} of type {t} has {len(src)} proposals (limit {L.MAX_PRODUCT_SOURCES})")
            sb += [s[0] for s in src]
            sr += [s[1] for s in src]
            off[i + 1] = len(sb)
        plans[t] = (off, np.asarray(sb, np.int32), np.asarray(sr, np.int32))
    return plans, buffers


class GibbsSolver:
    """Device-resident sweeps over a DeviceGraph.

    distributed=True (inside an initialised torch.distributed job, one rank per GPU, the same graph and particles on
    every rank -- SURVEY.md 8e): every rank convolves only its contiguous share of each family's factor list, the proposal
    rows are all-gathered (one collective per proposal buffer per sweep), and every rank takes the products of ALL
    variables -- the product sampler is keyed by (seed, sweep, variable, particle), so the replicated particle stores
    stay bit-identical without a second exchange."""

    def __init__(self, dg: DeviceGraph, gibbs_inner: int = 2, distributed: bool = False, group=None):
        self.dg, self.ctx, self.N = dg, dg.ctx, dg.N
        self.gibbs_inner = gibbs_inner
        self.plans, self.buffers = build_product_plans(dg.fg)
        self.distributed, self.group = distributed, group
        self.world, self.rank = 1, 0
        if distributed:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            import torch
            # collectives and kernels must be ordered on ONE stream: a dedicated torch stream that the ctx adopts
            # (torch's default stream is handle 0, which rome_b200_set_stream reads as "use the ctx's own stream")
            self.ctx.synchronize()  # uploads issued on the ctx's own stream land before the stream changes
            self._tstream = torch.cuda.Stream(device=self.ctx.device)
            self.ctx.set_stream(self._tstream.cuda_stream)
        Np = npad(self.N)
        self._dev, self._tensors = [], []
        for fam, which in self.buffers:
            d = FAMILY[fam][6] if which == "fwd" else FAMILY[fam][7]
            nF = self.ctx.num_factors(fam)
            if not d:
                self._dev.append(0)
            elif distributed:
                import torch
                from .sharding import shard_size
                t = torch.zeros((self.world * shard_size(nF, self.world), Np, d), dtype=torch.float32,
                                device=torch.device("cuda", self.ctx.device))
                self._tensors.append((len(self._dev), t, nF))
                self._dev.append(t.data_ptr())
            else:
                self._dev.append(self.ctx.malloc_device(max(1, nF * Np * d * 4)))
        for t, (off, sb, sr) in self.plans.items():
            self.ctx.set_product_plan(t, off, sb, sr)
        self.families = sorted({f for f, _ in self.buffers})
        self.sweeps_done = 0
        if not any(self._dev):
            raise ValueError("no factor of this graph has a closed-form proposal: nothing to sweep")
        if distributed:
            import torch
            torch.cuda.synchronize(self.ctx.device)  # the zero-fill of the proposal tensors ran on another stream

    def close(self):
        if not self.distributed:
            for p in self._dev:
                if p:
                    self.ctx.free_device(p)
        else:
            self.ctx.synchronize()
            self.ctx.set_stream(None)
        self._dev, self._tensors = [], []

    def sweep(self, seed: int = 0):
        """one synchronous sweep: every factor convolves (fused getSample + closed-form roots), then every variable
        takes the product of its proposals; 1 launch per family + 1-2 per variable type, nothing leaves the device"""
        c = self.ctx
        for k, fam in enumerate(self.families):
            dfwd, dbwd = FAMILY[fam][6], FAMILY[fam][7]
            flags = L.SAMPLE | (L.PROPOSAL_FWD if dfwd else 0) | (L.PROPOSAL_BWD if dbwd and FAMILY[fam][1] is not None else 0)
            if not flags & (L.PROPOSAL_FWD | L.PROPOSAL_BWD):
                continue
            if k > 0:
                flags |= L.INDEPENDENT  # the family kernels of one sweep read the same particles, write different buffers
            first, count = 0, -1
            if self.distributed:
                from .sharding import shard_range
                first, count = shard_range(c.num_factors(fam), self.rank, self.world)
            c.eval(fam, flags, seed=seed, stream_id=self.sweeps_done, first=first, count=count,
                   prop_fwd=self._dev[2 * k] or None, prop_bwd=self._dev[2 * k + 1] or None)
        if self.distributed:
            import torch
            from .sharding import allgather_rows
            with torch.cuda.stream(self._tstream):
                for _, t, nF in self._tensors:  # the one exchange of the path: every rank ends up with all proposal rows
                    allgather_rows(t, nF, self.group)
        some = next(p for p in self._dev if p)
        ptrs = [p or some for p in self._dev]  # placeholders for absent directions are never indexed by the plan
        for t in self.plans:
            c.product(t, ptrs, seed=seed, stream_id=self.sweeps_done, gibbs_iters=self.gibbs_inner, reanchor=True)
        self.sweeps_done += 1

    def solve(self, sweeps: int | None = None, seed: int = 0):
        for s in range(sweeps if sweeps is not None else self.dg.fg.solverParams.gibbsIters):
            self.sweep(seed + s)
        self.ctx.synchronize()


def solveGraphGibbs(fg: FactorGraph, sweeps: int | None = None, seed: int = 0, ctx: Context | None = None,
                    N: int | None = None) -> FactorGraph:
    """Run `sweeps` (default SolverParams.gibbsIters) device-resident sweeps on an initialised graph and write the new
    particles back into `fg` (the caller of the hot path in the reference is solveTree!; see the module docstring for
    what this is and is not)."""
    dg = DeviceGraph(fg, ctx=ctx, N=N)
    gs = GibbsSolver(dg)
    try:
        gs.solve(sweeps, seed)
        dg.download_particles()
    finally:
        gs.close()
    return fg
