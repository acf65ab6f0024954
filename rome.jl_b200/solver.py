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

    Buffer numbering: dense, in ascending family id, one buffer per (family, direction) that actually HAS a closed-form
    proposal (prop_fwd of families with a forward root, prop_bwd of binary families with a backward root) -- a 2-D SLAM
    graph with every family present stays far below the library's limit of ROME_B200_MAX_PRODUCT_BUFFERS.  A factor
    contributes prop_fwd to its LAST variable (or to the prior's variable) and prop_bwd to its FIRST variable.
    Returns (plans, buffers): plans[vartype] = (var_offsets, src_buf, src_row); buffers = [(family, "fwd"|"bwd")]."""
    fams = sorted(f for f in FAMILY if any(x.fnc.family == f for x in fg.factors.values()))
    if families is not None:
        fams = [f for f in fams if f in families]
    buffers = []
    for f in fams:
        if FAMILY[f][6]:
            buffers.append((f, "fwd"))
        if FAMILY[f][7] and FAMILY[f][1] is not None:
            buffers.append((f, "bwd"))
    if len(buffers) > L.MAX_PRODUCT_BUFFERS:
        raise ValueError(f"the graph needs {len(buffers)} proposal buffers, the library takes {L.MAX_PRODUCT_BUFFERS} per "
                         f"product call (ROME_B200_MAX_PRODUCT_BUFFERS)")
    index = {b: i for i, b in enumerate(buffers)}
    per_var = {t: {} for t in VAR_DIM}
    for v in fg.variables.values():
        per_var[v.variableType.vartype][v.index] = []
    for x in sorted(fg.factors.values(), key=lambda x: (x.fnc.family, x.index)):
        fam = x.fnc.family
        if fam not in fams:
            continue
        vt0, vt1, _, _, _, _, dfwd, dbwd = FAMILY[fam]
        vs = [fg.variables[l] for l in x.variableOrderSymbols]
        if (fam, "fwd") in index:
            tgt = vs[-1]
            per_var[tgt.variableType.vartype][tgt.index].append((index[(fam, "fwd")], x.index))
        if (fam, "bwd") in index:
            per_var[vs[0].variableType.vartype][vs[0].index].append((index[(fam, "bwd")], x.index))
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
            if distributed:
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
        self.convolve(seed)
        self.update(seed)
        self.sweeps_done += 1

    def convolve(self, seed: int = 0):
        """first half of a sweep: one launch per family writes every factor's proposal rows (+ the exchange when distributed)"""
        c = self.ctx
        index = {b: i for i, b in enumerate(self.buffers)}
        launched = False  # has a kernel of THIS sweep been launched yet?
        for fam in self.families:
            fwd, bwd = index.get((fam, "fwd")), index.get((fam, "bwd"))
            flags = L.SAMPLE | (L.PROPOSAL_FWD if fwd is not None else 0) | (L.PROPOSAL_BWD if bwd is not None else 0)
            first, count = 0, c.num_factors(fam)
            if self.distributed:
                from .sharding import shard_range
                first, count = shard_range(c.num_factors(fam), self.rank, self.world)
            if count == 0:
                continue  # nothing is launched for an empty share
            if launched:
                # the family kernels of one sweep read the same particles and write different buffers; the FIRST launch of
                # a sweep is not flagged: its predecessor is the previous sweep's product / re-anchor kernel, which
                # wrote the particle store
                flags |= L.INDEPENDENT
            c.eval(fam, flags, seed=seed, stream_id=self.sweeps_done, first=first, count=count,
                   prop_fwd=self._dev[fwd] if fwd is not None else None, prop_bwd=self._dev[bwd] if bwd is not None else None)
            launched = True
        if self.distributed:
            import torch
            from .sharding import allgather_rows
            with torch.cuda.stream(self._tstream):
                for _, t, nF in self._tensors:  # the one exchange of the path: every rank ends up with all proposal rows
                    allgather_rows(t, nF, self.group)

    def update(self, seed: int = 0):
        """second half of a sweep: every variable takes the product of its proposal densities (+ re-anchoring)"""
        ptrs = list(self._dev)
        for t in self.plans:
            self.ctx.product(t, ptrs, seed=seed, stream_id=self.sweeps_done, gibbs_iters=self.gibbs_inner, reanchor=True)

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


class OwnerShardedSolver:
    """Device-resident sweeps with the graph PARTITIONED over the ranks of a torch.distributed job (one rank per GPU):
    every rank owns a contiguous range of the variables of each type (sharding.OwnerSharding), holds only those plus
    the halo copies its factors read, evaluates the factors placed on it, and takes the products of its own variables.
    Per sweep only cut-edge data crosses NVLink, all of it written by the GPUs themselves into CUDA-IPC-mapped peer
    memory:

        eval      every local factor: fused getSample + forward / backward roots; the forward rows of CUT factors go
                  straight into the receive buffer of the target variable's owner (rome_b200_set_proposal_destinations)
        barrier A (rome_b200_peer_signal / _wait): all rows have landed
        product   of every OWNED variable's proposals (local forward / backward rows + received rows), re-anchor
        push      the new particle blocks of variables that are halos elsewhere (rome_b200_push_halo)
        barrier B all halo blocks have landed -- also what keeps the next sweep's rows out of receive buffers a slower
                  peer is still reading

    `w` is a workload (rome_b200.workloads: `graph_arrays(fg)` of an initialised host graph, or `manhattan_arrays`),
    identical on every rank; nothing but the CUDA IPC handles is exchanged through torch.distributed.  The samplers are
    keyed by LOCAL factor / variable numbers, so a sharded solve is a different (equally valid) random realisation than
    a single-GPU one: tests compare beliefs statistically and the exchanged rows / blocks bit for bit."""

    def __init__(self, w, ctx: Context, group=None, gibbs_inner: int = 2):
        import torch
        import torch.distributed as dist
        from . import workloads as W
        self.ctx, self.N, self.gibbs_inner, self.group = ctx, w["N"], gibbs_inner, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        G, rank, Np = self.world, self.rank, npad(w["N"])
        self.sh = W.sharding_of(w, G)
        self.lv = lv = W.local_view(w, self.sh, rank, fill_halo=True)
        c = ctx
        c.synchronize()
        self._tstream = torch.cuda.Stream(device=c.device)
        c.set_stream(self._tstream.cuda_stream)
        for vt, p in lv["particles"].items():
            c.set_particles(vt, p)
            c.set_owned_variables(vt, lv["loc"]["own"][vt][1] - lv["loc"]["own"][vt][0])
        self.fams = sorted(f for f in lv["families"] if FAMILY[f][6] or (FAMILY[f][7] and FAMILY[f][1] is not None))
        self.bufs, names = {}, []          # (family, "fwd" | "bwd" | "recv") -> device pointer
        for fam in self.fams:
            f = lv["families"][fam]
            W.upload_family(c, fam, f["i0"], f["i1"], f["a"], f["b"])
            nF, dfwd, dbwd = len(f["i0"]), FAMILY[fam][6], FAMILY[fam][7] if FAMILY[fam][1] is not None else 0
            if dfwd:
                self.bufs[(fam, "fwd")] = c.malloc_device(max(1, nF) * Np * dfwd * 4)
                self.bufs[(fam, "recv")] = c.malloc_device(max(1, len(f["recv"])) * Np * dfwd * 4)
            if dbwd:
                self.bufs[(fam, "bwd")] = c.malloc_device(max(1, nF) * Np * dbwd * 4)
        names = sorted(self.bufs)
        if len(names) > L.MAX_PRODUCT_BUFFERS:
            raise ValueError(f"{len(names)} proposal buffers exceed the limit of {L.MAX_PRODUCT_BUFFERS}")
        self.names = names
        # product plans over the OWNED variables: local forward rows of interior factors, backward rows of all local
        # factors (their first variable is owned here), received forward rows of the peers' cut factors
        per_var = {vt: [[] for _ in range(lv["loc"]["own"][vt][1] - lv["loc"]["own"][vt][0])] for vt in lv["particles"]}
        for fam in self.fams:
            f = lv["families"][fam]
            vt0, vt1 = FAMILY[fam][0], FAMILY[fam][1]
            tgt_vt = vt1 if vt1 is not None else vt0
            tgt = f["i1"] if f["i1"] is not None else f["i0"]
            if (fam, "fwd") in self.bufs:
                b = names.index((fam, "fwd"))
                for fl in range(len(f["i0"])):
                    if not (f["cut_first"] <= fl < f["cut_first"] + f["n_cut"]):   # interior: the target is owned here
                        per_var[tgt_vt][int(tgt[fl])].append((b, fl))
                b = names.index((fam, "recv"))
                F = self.sh.families[fam]
                lo = lv["loc"]["own"][tgt_vt][0]
                for k, g in enumerate(f["recv"]):
                    per_var[tgt_vt][int(F["i1"][g]) - lo].append((b, k))
            if (fam, "bwd") in self.bufs:
                b = names.index((fam, "bwd"))
                for fl in range(len(f["i0"])):
                    per_var[vt0][int(f["i0"][fl])].append((b, fl))
        self.plans = []
        for vt, lists in per_var.items():
            off = np.zeros(len(lists) + 1, np.int32)
            sb, sr = [], []
            for i, src in enumerate(lists):
                if len(src) > L.MAX_PRODUCT_SOURCES:
                    raise ValueError(f"variable {i} of type {vt} has {len(src)} proposals (limit {L.MAX_PRODUCT_SOURCES})")
                sb += [s[0] for s in src]
                sr += [s[1] for s in src]
                off[i + 1] = len(sb)
            if len(lists):
                c.set_product_plan(vt, off, np.asarray(sb, np.int32), np.asarray(sr, np.int32))
                self.plans.append(vt)
        # CUDA IPC: receive buffers, particle stores, barrier state
        self.state = c.peer_state_alloc()
        mine = {"recv": {fam: c.ipc_export(self.bufs[(fam, "recv")]) for fam in self.fams if (fam, "recv") in self.bufs},
                "store": {vt: c.ipc_export(c.particles_device(vt)[0]) for vt in lv["particles"]},
                "state": c.ipc_export(self.state)}
        c.synchronize()
        everyone = [None] * G
        dist.all_gather_object(everyone, mine, group=group)
        self.peer_slots = [c.ipc_import(everyone[p]["state"]) + 4 * (rank if rank < p else rank - 1)
                           for p in range(G) if p != rank]
        recv_ptr = {p: {fam: c.ipc_import(h) for fam, h in everyone[p]["recv"].items()} for p in range(G) if p != rank}
        store_ptr = {p: {vt: c.ipc_import(h) for vt, h in everyone[p]["store"].items()} for p in range(G) if p != rank}
        for fam in self.fams:
            f = lv["families"][fam]
            if (fam, "recv") in self.bufs and f["n_cut"]:
                rowb = Np * FAMILY[fam][6] * 4
                c.set_proposal_destinations(fam, 0, [0] * f["cut_first"] +
                                            [recv_ptr[int(d)][fam] + int(r) * rowb for d, r in zip(f["dst_rank"], f["dst_row"])] +
                                            [0] * (len(f["i0"]) - f["cut_first"] - f["n_cut"]))
        for vt, pushes in lv["loc"]["push"].items():
            src, dst, bb = [], [], c.particles_device(vt)[1]
            for reader, local_vars, slots in pushes:
                src += [int(v) for v in local_vars]
                dst += [store_ptr[reader][vt] + int(sl) * bb for sl in slots]
            if src:
                c.set_halo_plan(vt, src, dst)
        self.sweeps_done = 0
        c.synchronize()
        dist.barrier(group=group)

    def _barrier(self):
        self.ctx.peer_barrier(self.state, self.peer_slots)

    def sweep(self, seed: int = 0):
        c, lv = self.ctx, self.lv
        first = True
        for fam in self.fams:
            nF = len(lv["families"][fam]["i0"])
            if nF == 0:
                continue
            flags = L.SAMPLE | (L.PROPOSAL_FWD if (fam, "fwd") in self.bufs else 0) | \
                (L.PROPOSAL_BWD if (fam, "bwd") in self.bufs else 0) | (0 if first else L.INDEPENDENT)
            c.eval(fam, flags, seed=seed, stream_id=self.sweeps_done, prop_fwd=self.bufs.get((fam, "fwd")),
                   prop_bwd=self.bufs.get((fam, "bwd")))
            first = False
        self._barrier()                                   # A: every cut factor's row sits in its owner's receive buffer
        ptrs = [self.bufs[n] for n in self.names]
        for vt in self.plans:
            c.product(vt, ptrs, seed=seed, stream_id=self.sweeps_done, gibbs_iters=self.gibbs_inner, reanchor=True)
        for vt in lv["particles"]:
            c.push_halo(vt)
        self._barrier()                                   # B: every halo copy is the owner's new block
        self.sweeps_done += 1

    def solve(self, sweeps: int = 3, seed: int = 0):
        for s in range(sweeps):
            self.sweep(seed + s)
        self.ctx.synchronize()
        if self.ctx.peer_gave_up(self.state):
            raise RuntimeError("a peer never reached the barrier (rome_b200_peer_wait gave up)")

    def owned_particles(self):
        """{vartype: (global ids, Float64 [n_owned][N][d])} of the variables this rank owns"""
        out = {}
        for vt in self.lv["particles"]:
            lo, hi = self.lv["loc"]["own"][vt]
            out[vt] = (np.arange(lo, hi), self.ctx.get_particles(vt)[:hi - lo])
        return out

    def close(self):
        self.ctx.synchronize()
        for p in self.bufs.values():
            self.ctx.free_device(p)
        self.bufs = {}
        self.ctx.set_stream(None)
