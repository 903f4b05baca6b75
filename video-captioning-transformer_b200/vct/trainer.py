"""Native training step: the loop body of the reference's ``train_epoch`` (train.py:119-131 --
forward, zero_grad, backward, optimizer.step, loss all-reduce) as ONE launch sequence over the flat
arenas, optionally captured in a CUDA graph.

    per step:  H2D stage inputs -> [tick | zero dE | forward plan (fused SCE grad) | backward plan |
               gradient all-reduce (NCCL, bucketed, overlapped with backward) | vct_adam] -> loss

Data parallelism is the reference's: one process per GPU, batch sharded by the caller
(DistributedSampler, dataloader.py:524), gradients averaged across ranks (DDP, train.py:218).  Here the
"bucket" is a slice of the flat gradient arena, so no flatten/copy is needed; slices are all-reduced
(SUM; in bf16 when the engine computes in bf16) on a side stream as soon as the backward plan has produced
them and the 1/world_size factor is folded into the Adam kernel.  The embedding-table gradient -- 94 MB
dense, at most B*S non-zero rows -- is exchanged in its sparse form (all-gather of rows + ids, local scatter).
"""


from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch


def gather_embedding_rows(rows: torch.Tensor, ids: torch.Tensor, group=None):
    """Sparse exchange of the embedding-table gradient: every rank contributes its [B*S, d] masked gradient rows and
    its [B, S+1] token ids; returns (all_rows [world*B*S, d], all_ids [world*B, S+1]).  Scattering all_rows by
    all_ids[:, :-1] (skipping pad ids) gives the SUM over ranks of the dense table gradients -- what DDP's all-reduce of
    the dense gradient produces (train.py:218) before its 1/world averaging.  Any backend (nccl / gloo)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    all_rows = rows.new_empty((world * rows.shape[0],) + tuple(rows.shape[1:]))
    all_ids = ids.new_empty((world * ids.shape[0],) + tuple(ids.shape[1:]))
    dist.all_gather_into_tensor(all_rows, rows.contiguous(), group=group)
    dist.all_gather_into_tensor(all_ids, ids.contiguous(), group=group)
    return all_rows, all_ids


def scatter_embedding_rows(all_rows: torch.Tensor, all_ids: torch.Tensor, V: int, pad_id: int) -> torch.Tensor:
    """Dense [V, d] table gradient from gathered rows (host-side restatement of vct_embed_bwd with drop_p = 0;
    used by the CPU tests of the exchange scheme)."""
    d = all_rows.shape[-1]
    tgt = all_ids[:, :-1].reshape(-1)
    keep = tgt != pad_id
    out = torch.zeros(V, d, dtype=all_rows.dtype, device=all_rows.device)
    out.index_add_(0, tgt[keep], all_rows.reshape(-1, d)[keep])
    return out


def gradient_buckets(arena, order: List[str], max_bytes: int = 64 << 20) -> List[Tuple[int, int]]:
    """Partition the arena into contiguous [start, end) element ranges, in the order the backward plan
    finishes them (``order`` = parameter-name prefixes, last-finished last).  Every element of the arena
    belongs to exactly one bucket."""
    names = arena.names
    spans = []
    for pre in order:
        idx = [i for i, n in enumerate(names) if n.startswith(pre)]
        if not idx:
            continue
        lo = arena.offset[names[idx[0]]]
        last = names[idx[-1]]
        hi = arena.offset[names[idx[-1] + 1]] if idx[-1] + 1 < len(names) else arena.numel
        spans.append((lo, hi))
    covered = sorted(spans)
    pos = 0
    for lo, hi in covered:
        if lo != pos:
            raise ValueError(f"bucket order leaves a gap or overlap at element {pos} (next span starts at {lo})")
        pos = hi
    if pos != arena.numel:
        raise ValueError("bucket order does not cover the arena")
    out = []
    max_el = max_bytes // 4
    for lo, hi in spans:
        while hi - lo > max_el:
            out.append((lo, lo + max_el))
            lo += max_el
        out.append((lo, hi))
    return out


def split_segments(calls) -> List[Tuple[list, list]]:
    """Split a backward plan (``Plan.calls`` entries ``(name, fn, args, lane)``) at every run of optimizer-lane (lane 2)
    calls: returns ``[(gpu_calls, optimizer_calls), ...]`` in issue order, where ``gpu_calls`` are the main-lane /
    weight-gradient-lane entries that precede the run (possibly empty) and the last pair may have no optimizer calls."""
    segs, cur = [], []
    in_opt = False
    for c in calls:
        if c[3] == 2:
            if not in_opt:
                segs.append((cur, []))
                cur = []
                in_opt = True
            segs[-1][1].append(c)
        else:
            cur.append(c)
            in_opt = False
    if cur:
        segs.append((cur, []))
    return segs


def all_reduce_flat(flat: torch.Tensor, buckets: List[Tuple[int, int]], group=None) -> None:
    """SUM all-reduce of each bucket slice (any backend: nccl on GPU, gloo in the CPU tests)."""
    import torch.distributed as dist
    for lo, hi in buckets:
        dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group)


class CaptionTrainer:
    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 use_graph: bool = True, process_group=None, world_size: Optional[int] = None, uniform_shapes: bool = False):
        """uniform_shapes (data parallel only): the caller promises that every rank calls ``step`` with the same batch size
        and the same caption length in a given step (synthetic batches; loaders that pad to a fixed length).  The
        embedding-table gradient is then exchanged in its sparse form (<= B*S rows per rank instead of the dense [V, d]
        table).  Without the promise -- ragged batches whose longest caption differs from rank to rank, as the reference's
        loader produces them -- the dense gradient is exchanged, which needs no agreement on shapes."""
        import torch.distributed as dist
        self.model = model
        self.engine = model._engine()
        if model.f_type != "caption":
            raise ValueError("CaptionTrainer drives the caption task: call model.mode('caption') first")
        self.engine.uniform_shapes = bool(uniform_shapes)
        self.engine.set_adam(lr, betas, eps, weight_decay)
        self.engine.refresh_shadow(force=True)
        # invariant of the native step: the embedding-table gradient is all-zero when a step starts (each step re-zeroes
        # only the rows it scattered into, vct_embed_zero); establish it once
        self.engine.zero_scatter_grads()
        self.use_graph = use_graph
        self.fuse_adam = os.environ.get("VCT_FUSE_ADAM", "1") != "0" and self.engine.side_streams is not None
        self.group = process_group
        self.world = world_size if world_size is not None else (dist.get_world_size(process_group)
                                                                 if dist.is_available() and dist.is_initialized() else 1)
        self._graphs_global = {}       # graphs that touch no per-shape workspace (the stand-alone Adam launch)
        self._graphs, self._warm = self._graphs_global, {}
        # N > 1: backward as CUDA-graph segments between the optimizer slices (VCT_SEGMENTED=0: fully eager backward).
        # (Capturing the NCCL all-reduces themselves inside one step graph hung on this stack -- torch 2.11 + NCCL 2.28.9,
        # 2 x B200, both ranks stall in the first replay -- so the collectives stay outside the graphs.)
        self.segmented = os.environ.get("VCT_SEGMENTED", "1") != "0"
        # The step is captured on a HIGH-priority stream: its kernels (the latency-bound dependent chain of forward and
        # backward) are the critical path, while the side lanes (weight gradients, column sums, Adam slices -- default,
        # i.e. lowest, priority) only need to finish by the end of the step.  The block scheduler then hands freed SM
        # slots to main-lane CTAs first instead of letting a 1184-CTA Adam launch stall the next GEMM of backward.
        self._cap_stream = torch.cuda.Stream(device=self.engine.device, priority=-1) \
            if os.environ.get("VCT_MAIN_PRIORITY", "0") == "1" else None
        self.buckets = None
        self.peer = None
        if self.world > 1:
            # backward finishes the decoder (generator first, embedding last) before the encoder
            self.buckets = gradient_buckets(self.engine.arena, ["video_encoder.", "cap_decoder."])
            self.comm_stream = torch.cuda.Stream(device=self.engine.device)
            self._attach_peer_comm()

    def _attach_peer_comm(self) -> None:
        """Gradient exchange over NVLink peer memory (csrc/peer_comm.cu) when every rank of the group can map every other
        rank's memory; otherwise (VCT_COMM=nccl, fp32 exchange, no P2P) the NCCL path below stays in charge.  The decision
        is made collectively: one rank that cannot take part switches all of them."""
        import torch.distributed as dist
        from .engine import BF16
        from .peer import PeerComm, peer_comm_supported
        eng = self.engine
        rank = dist.get_rank(self.process_group_or_default())
        ok = self.fuse_adam and eng.cdt == BF16 and eng.grad_comm_dtype == BF16 and peer_comm_supported(eng.device, self.world)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=eng.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            return
        try:
            comm = PeerComm(rank, self.world, eng.peer_region_bytes(eng.arena.numel, eng.dims.d), eng.device, group=self.group)
            good = 1
        except Exception as e:                                   # e.g. cudaIpc refused inside this container
            import sys
            print(f"[vct] peer-memory communication unavailable on rank {rank} ({e}); using NCCL", file=sys.stderr, flush=True)
            comm, good = None, 0
        flag.fill_(good)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            if comm is not None:
                comm.close()
            return
        eng.attach_peer(comm)
        self.peer = comm

    def process_group_or_default(self):
        return self.group

    def set_lr(self, lr: float) -> None:
        self.engine.set_lr(lr)

    # ---- one step -----------------------------------------------------------------------------------
    def _compute(self, ws, fuse_adam: bool = False, allreduce=None) -> None:
        eng = self.engine
        eng.tick()
        world = allreduce[1] if allreduce is not None else 1
        # (the embedding-table gradient is all-zero here: the previous step cleared the rows it had scattered into)
        if fuse_adam and eng.emb_mode(ws, world) in ("local", "sparse"):
            eng.run(eng.plan_embed_early(ws, world))    # optimizer lane, un-joined: overlaps the whole forward
        eng.run(eng.plan_forward(ws, fused_grad=True, part="all"))
        eng.run(eng.plan_backward(ws, sce_first=False, part="all", fuse_adam=fuse_adam, allreduce=allreduce))

    def _forward(self, ws, tick: bool = True) -> None:
        eng = self.engine
        early = tick and self.fuse_adam and eng.emb_mode(ws, 1) == "local"
        if tick:
            eng.tick()
        if early:
            eng.run(eng.plan_embed_early(ws, 1))
        eng.run(eng.plan_forward(ws, fused_grad=True, part="all"))
        if early and eng.side_streams:
            # this graph ends with the forward: the optimizer lane forked above must re-join before capture ends
            torch.cuda.current_stream(eng.device).wait_stream(eng.side_streams[1])

    def _early_embedding(self, ws) -> None:
        """Data parallel, start of a step (eager, outside the graphs): all-gather the token ids of every rank on the comm
        stream, then -- optimizer stream -- stamp the touched table rows, sort the gathered tokens for the deterministic
        scatter and run Adam on all untouched rows of the embedding table.  Everything here overlaps the forward graph."""
        import torch.distributed as dist
        from . import lib as L
        eng = self.engine
        if eng.emb_mode(ws, self.world) != "sparse":
            return
        main = torch.cuda.current_stream(eng.device)
        opt, comm = eng.side_streams[1], self.comm_stream
        all_ids = eng.emb_buffers(ws, self.world)[0]
        ev = ws.graphs.setdefault("_early_events", (torch.cuda.Event(), torch.cuda.Event()))
        ev[0].record(main)                                   # ids staged, step ticked
        comm.wait_event(ev[0])
        opt.wait_event(ev[0])
        with torch.cuda.stream(comm):
            dist.all_gather_into_tensor(all_ids, ws.ids, group=self.group)
        ev[1].record(comm)
        opt.wait_event(ev[1])
        for name, fn, args, _lane in eng.plan_embed_early(ws, self.world).calls:
            L.check(fn(*args, opt.cuda_stream), name)
            eng.launches += 1

    def _update(self, ws=None) -> None:
        self.engine.adam(grad_scale=1.0 / self.world)
        if ws is not None:
            self.engine.embed_zero(ws)

    def _graphed(self, key, fn, ws=None):
        """Run ``fn`` eagerly twice (plan building, kernel attributes), then capture it once and replay.  Graphs over a
        workspace's pointers are stored IN that workspace (``ws.graphs``), so evicting the workspace from the engine's
        LRU cache drops them with it."""
        eng = self.engine
        if not self.use_graph:
            fn()
            return
        store = ws.graphs if ws is not None else self._graphs_global
        self._graphs, self._warm = store, store.setdefault("_warm", {})
        if key not in self._graphs:
            if self._warm.get(key, 0) < 2:
                self._warm[key] = self._warm.get(key, 0) + 1
                fn()
                return
            g = torch.cuda.CUDAGraph()
            before = eng.launches
            if self._cap_stream is not None:
                self._cap_stream.wait_stream(torch.cuda.current_stream(eng.device))
                with torch.cuda.graph(g, stream=self._cap_stream):
                    fn()
            else:
                with torch.cuda.graph(g):
                    fn()
            self._graphs[key] = (g, eng.launches - before)
            eng.launches = before          # capture issued nothing; the replay below is this step
        g, n = self._graphs[key]
        g.replay()
        eng.launches += n

    def _backward_segments(self, ws, key) -> None:
        """Backward + per-slice (all-reduce, Adam) as a chain of CUDA-graph segments (see ``step``).  The backward plan is
        split at every run of optimizer-lane calls; segment k holds the main-lane and weight-gradient-lane launches that
        precede slice k (the side lane re-joins the main lane at the end of a segment, as graph capture requires)."""
        from .engine import Plan
        eng = self.engine
        main = torch.cuda.current_stream(eng.device)
        opt = eng.side_streams[1]
        ar = (self.group, self.world) if self.world > 1 else None
        plan = eng.plan_backward(ws, sce_first=False, part="all", fuse_adam=True, allreduce=ar)
        warm = ws.graphs.setdefault("_warm", {})
        segments = ws.graphs.setdefault("_segments", {})
        if warm.get(key + ("segments",), 0) < 2 or not self.use_graph:
            warm[key + ("segments",)] = warm.get(key + ("segments",), 0) + 1
            eng.run(plan)
            return
        if key not in segments:
            built = []
            for gpu_calls, opt_calls in split_segments(plan.calls):      # optimizer lane: py:all_reduce / vct_adam
                g, n = None, 0
                real = [c for c in gpu_calls if c[0] != "join"]
                if real:
                    sub = Plan()
                    sub.calls, sub.keep = list(gpu_calls), plan.keep
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        n = sub.run(torch.cuda.current_stream(eng.device), eng.side_streams)
                built.append((g, n, opt_calls, torch.cuda.Event(), sub if real else None,
                              [torch.cuda.Event() for _ in opt_calls]))
            segments[key] = (built, torch.cuda.Event(), torch.cuda.Event())
        built, done, done_comm = segments[key]
        from . import lib as L
        # Two optimizer-side streams: the collectives (and the fp32 -> bf16 casts that feed them) queue on the COMM stream
        # back to back, the Adam / scatter kernels on the OPT stream, each waiting only for the collective it consumes.
        # A slice's all-reduce therefore overlaps the previous slice's Adam instead of waiting behind it.
        comm = self.comm_stream if self.world > 1 else opt
        for g, n, opt_calls, ev, _sub, evs in built:
            if g is not None:
                g.replay()
                eng.launches += n
            if opt_calls:
                ev.record(main)
                opt.wait_event(ev)
                if comm is not opt:
                    comm.wait_event(ev)
                pending = None                       # event of the last collective the OPT stream has not waited for yet
                for (name, fn, args, _lane), e in zip(opt_calls, evs):
                    on_comm = comm is not opt and (name.startswith("py:") or name.startswith("vct_cast"))
                    st = comm if on_comm else opt
                    if not on_comm and pending is not None:
                        opt.wait_event(pending)
                        pending = None
                    rc = fn(st) if name.startswith("py:") else fn(*args, st.cuda_stream)
                    if rc:
                        L.check(rc, name)
                    if not name.startswith("py:"):
                        eng.launches += 1
                    if on_comm:
                        e.record(comm)
                        pending = e
        done.record(opt)
        main.wait_event(done)
        if comm is not opt:
            done_comm.record(comm)
            main.wait_event(done_comm)

    # ---- input prefetch ---------------------------------------------------------------------------------
    def prefetch(self, feats: torch.Tensor, vid_pad: Optional[torch.Tensor], ids: torch.Tensor) -> None:
        """Start the host -> device copy of the NEXT batch on a copy stream while the current step is still running (the
        reference's loop copies inside the step, train.py:120-121: ``.to(device)`` per modality + a tokenizer pass).  The
        next ``step()`` called with the same three tensors (or ``step_prefetched()``) consumes the staged copy with a
        device-to-device move instead of crossing PCIe on the critical path.  Pinned host tensors make the copy truly
        asynchronous."""
        eng = self.engine
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=eng.device)
            self._staged = None
        B, T, Din = feats.shape
        key = (B, T, Din, ids.shape[1], vid_pad is not None)
        bufs = getattr(self, "_stage_bufs", {})
        if key not in bufs:
            bufs[key] = (torch.empty((B, T, Din), dtype=torch.float32, device=eng.device),
                         torch.empty((B, T), dtype=torch.bool, device=eng.device) if vid_pad is not None else None,
                         torch.empty(tuple(ids.shape), dtype=torch.int64, device=eng.device), torch.cuda.Event())
            self._stage_bufs = bufs
        f, v, i, ev = bufs[key]
        cs = self._copy_stream
        free = getattr(self, "_stage_free", None)
        if free is not None:
            cs.wait_event(free)          # the previous consumer (a device-to-device move at the START of a step) is done;
                                         # waiting for the whole stream would push this copy behind the running step
        with torch.cuda.stream(cs):
            f.copy_(feats, non_blocking=True)
            if v is not None:
                v.copy_(vid_pad, non_blocking=True)
            i.copy_(ids, non_blocking=True)
            ev.record(cs)
        self._staged = (feats, vid_pad, ids, f, v, i, ev)

    def step_prefetched(self) -> torch.Tensor:
        if getattr(self, "_staged", None) is None:
            raise RuntimeError("step_prefetched() without a preceding prefetch()")
        feats, vid_pad, ids = self._staged[:3]
        return self.step(feats, vid_pad, ids)

    def _take_staged(self, feats, vid_pad, ids):
        st = getattr(self, "_staged", None)
        if st is None or st[0] is not feats or st[1] is not vid_pad or st[2] is not ids:
            return feats, vid_pad, ids
        self._staged = None
        torch.cuda.current_stream(self.engine.device).wait_event(st[6])
        return st[3], st[4], st[5]

    def step(self, feats: torch.Tensor, vid_pad: Optional[torch.Tensor], ids: torch.Tensor) -> torch.Tensor:
        """feats fp32 [B,T,Din], vid_pad bool [B,T] | None, ids int64 [B,S+1] (host -- ideally pinned -- or
        device tensors).  Returns the step's loss as a device scalar (no host sync).

        world == 1: [tick, forward, backward, Adam] is ONE CUDA graph.
        world  > 1: graph [tick, forward]; backward as graph segments whose optimizer lane all-reduces (NCCL SUM) and
        then updates each arena slice as soon as its gradient is final (VCT_SEGMENTED=0: eager backward;
        VCT_FUSE_ADAM=0: graph [tick, forward, backward] -> bucketed all-reduce -> graph [Adam])."""
        eng = self.engine
        B, T, _ = feats.shape
        S = ids.shape[1] - 1
        ws = eng.workspace(B, T, S, True)
        self._n_steps = getattr(self, "_n_steps", 0) + 1
        eng.check_arena(quick=(self._n_steps % 32 != 1))      # 79 data_ptr() calls cost ~20 us of host time per step
        eng.refresh_shadow()               # no-op unless the masters were edited outside vct_adam
        staged = getattr(self, "_staged", None) is not None
        feats, vid_pad, ids = self._take_staged(feats, vid_pad, ids)
        eng.stage_inputs(ws, feats, vid_pad, ids)
        if staged:
            if getattr(self, "_stage_free", None) is None:
                self._stage_free = torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream(eng.device))
        if self.world == 1 and self.fuse_adam and os.environ.get("VCT_FORCE_SEGMENTED") == "1":
            # test hook: the N > 1 execution scheme (forward graph + backward graph segments + eager optimizer lane) on one GPU
            self._graphed((B, T, S, "forward"), lambda: self._forward(ws), ws=ws)
            self._backward_segments(ws, (B, T, S))
        elif self.world == 1 and self.fuse_adam:
            # optimizer-in-backward: vct_adam runs slice by slice on a side lane while backward continues
            self._graphed((B, T, S, "step+adam"), lambda: self._compute(ws, fuse_adam=True), ws=ws)
        elif self.world == 1:
            self._graphed((B, T, S, "step"), lambda: (self._compute(ws), self._update(ws)), ws=ws)
        elif self.peer is not None and self.fuse_adam:
            # data parallel over NVLink peer memory: the collectives are kernels of the library, so [tick, forward, backward,
            # per-slice all-reduce + Adam] is ONE CUDA graph exactly like the single-GPU step
            if eng.emb_mode(ws, self.world) == "sparse":
                all_ids = eng.emb_buffers(ws, self.world)[0]
                r = self.peer.rank
                all_ids[r * B:(r + 1) * B].copy_(ws.ids, non_blocking=True)
            self._graphed((B, T, S, "step+adam+peer"),
                          lambda: self._compute(ws, fuse_adam=True, allreduce=(self.group, self.world)), ws=ws)
        elif self.fuse_adam and self.segmented:
            # data parallel: forward is one CUDA graph and backward a chain of graph SEGMENTS -- the launches between two
            # optimizer slices are captured together -- with the per-slice NCCL all-reduce + Adam issued eagerly on the
            # optimizer lane after each segment.  ~25 host operations per step instead of ~110 ctypes launches, so the
            # N-GPU step is no longer bound by the host.
            eng.tick()
            self._early_embedding(ws)
            self._graphed((B, T, S, "forward-only"), lambda: self._forward(ws, tick=False), ws=ws)
            self._backward_segments(ws, (B, T, S))
        elif self.fuse_adam:
            # data parallel with overlap: forward is a CUDA graph; backward runs eagerly (NCCL collectives outside graph
            # capture) -- each arena slice is all-reduced and then updated on the optimizer lane while the rest of
            # backward continues on the main lane
            eng.tick()
            self._early_embedding(ws)
            self._graphed((B, T, S, "forward-only"), lambda: self._forward(ws, tick=False), ws=ws)
            eng.run(eng.plan_backward(ws, sce_first=False, part="all", fuse_adam=True, allreduce=(self.group, self.world)))
            if self.world > 1:
                torch.cuda.current_stream(eng.device).wait_stream(self.comm_stream)
        else:
            # (dense exchange of every gradient incl. the embedding table: the reference's DDP scheme, kept for comparison)
            self._graphed((B, T, S, "compute"), lambda: self._compute(ws), ws=ws)
            all_reduce_flat(eng.arena.grad, self.buckets, self.group)
            self._graphed((B, T, S, "update"), lambda: self._update(ws), ws=ws)
        return ws.loss[0]
