"""Host-side sequencing of the hot path over libvct_b200.so.

``CaptionEngine`` owns the parameter arena, the per-shape activation workspaces and the launch
plans (lists of pre-built C-ABI calls) for

  * caption forward   (model/MMT4Caption.py:114-121 -> MMEncoder.py:244-276 -> CapDecoder.py:34-60
                       -> loss.py:78-92)
  * its backward      (train.py:125 ``loss.backward()``)
  * Adam              (train.py:126)
  * greedy decoding   (model/MMT4Caption.py:146-184, KV-cached; SURVEY Q18 shows equivalence)

PyTorch is used for device memory and streams only; every arithmetic step is a kernel of the
shared library.  There is no CPU path: constructing an engine without CUDA raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch

from . import lib as L
from .arena import ParamArena

F32, BF16 = L.F32, L.BF16
_ESIZE = {F32: 4, BF16: 2}
_TDT = {F32: torch.float32, BF16: torch.bfloat16}

# dropout call sites (unique per layer / position; see include/vct.h "dropout")
SITE_EMBED = 1
def _enc_site(layer: int, k: int) -> int: return 100 + 16 * layer + k      # 0 probs, 1 dropout1, 2 ffn, 3 dropout2
def _dec_site(layer: int, k: int) -> int: return 1000 + 16 * layer + k     # 0 self probs, 1 dropout1, 2 cross probs, 3 dropout2, 4 ffn, 5 dropout3


class Plan:
    """A pre-built sequence of C-ABI calls.  ``run`` issues them in order on the main stream; calls added
    with ``lane=1`` (weight-gradient GEMMs and bias column sums, which nothing on the critical path of
    backward consumes) go to a side stream that forks from / joins the main stream through events, so that
    under CUDA-graph capture they become parallel branches of the graph."""

    def __init__(self):
        self.calls: List[Tuple] = []
        self.keep: List = []      # keeps ctypes structs / tensors alive
        self.lane = 0             # lane given to calls added while it is set (see CaptionEngine._side)
        self.no_final_join = False   # leave the side lanes un-joined: a later plan of the same step joins them
        self._events: List = []

    def add(self, name: str, fn, *args):
        self.calls.append((name, fn, args, self.lane))

    def add_join(self):
        """main waits here for everything the side lanes have been given so far."""
        self.calls.append(("join", None, (), 0))

    def run(self, main, sides=None) -> int:
        """main: torch.cuda.Stream; sides: list of side streams (lane k runs on sides[k-1]) or None (everything
        runs on main).  A side-lane call waits for everything main has issued so far; lane 2 (the optimizer lane)
        additionally waits for lane 1 (whose kernels produce the gradients it consumes).  main joins every used
        side lane at the end."""
        mp = main.cuda_stream
        n_kernels = sum(1 for c in self.calls if not c[0].startswith("py:") and c[0] != "join")
        if not sides or not any(c[3] for c in self.calls):
            for name, fn, args, _ in self.calls:
                if name == "join":
                    continue
                rc = fn(main) if name.startswith("py:") else fn(*args, mp)
                if rc:
                    L.check(rc, name)
            return n_kernels
        ev_i = 0

        def event():
            nonlocal ev_i
            if ev_i == len(self._events):
                self._events.append(torch.cuda.Event())
            ev_i += 1
            return self._events[ev_i - 1]

        nl = len(sides) + 1
        main_seen = [False] * nl      # has lane k already waited for main's latest work?
        lane1_seen2 = True            # has lane 2 already waited for lane 1's latest work?
        used = [False] * nl
        for name, fn, args, lane in self.calls:
            lane = min(lane, nl - 1)
            if name == "join":
                for k in range(1, nl):
                    if used[k]:
                        e = event()
                        e.record(sides[k - 1])
                        main.wait_event(e)
                        used[k] = False
                continue
            if lane == 0:
                rc = fn(main) if name.startswith("py:") else fn(*args, mp)
                main_seen = [False] * nl
            else:
                st = sides[lane - 1]
                if not main_seen[lane]:
                    e = event()
                    e.record(main)
                    st.wait_event(e)
                    main_seen[lane] = True
                if lane == 2 and not lane1_seen2 and used[1]:
                    e = event()
                    e.record(sides[0])
                    st.wait_event(e)
                    lane1_seen2 = True
                rc = fn(st) if name.startswith("py:") else fn(*args, st.cuda_stream)
                used[lane] = True
                if lane == 1:
                    lane1_seen2 = False
            if rc:
                L.check(rc, name)
        for k in range(1, nl):
            if used[k] and not self.no_final_join:
                e = event()
                e.record(sides[k - 1])
                main.wait_event(e)
        return n_kernels

    def __len__(self):
        return len(self.calls)


class CaptionEngine:
    def __init__(self, video_encoder, cap_decoder, *, dims: dict, device: torch.device, precision: str = "bf16",
                 gemm_impl: Optional[str] = None, seed: int = 666):
        if device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("vct_b200 has no CPU path: a CUDA (sm_100a) device is required")
        self.lib = L.load()
        sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
        with torch.cuda.device(device):
            L.check(self.lib.vct_device_info(C.byref(sm), C.byref(maj), C.byref(mnr)), "vct_device_info")
        self.device = device
        self.dims = SimpleNamespace(**dims)   # Din d H_enc H_dec F_enc F_dec L_enc L_dec V pad_id alpha dropout
        if precision not in ("bf16", "fp32", "bf16x3", "bf16x6"):
            raise ValueError("precision must be 'bf16', 'fp32', 'bf16x3' or 'bf16x6'")
        self.precision = precision
        self.cdt = BF16 if precision == "bf16" else F32
        # bf16x3 / bf16x6: fp32 storage like 'fp32', but every dense contraction runs on tcgen05 with the fp32 operands split
        # into 2 / 3 bf16 pieces and 3 / 6 cross terms accumulated in fp32 (csrc/gemm_split.cu) -- the reference-precision
        # tensor-core mode (token-id argmax exactness without the SIMT FFMA GEMM)
        self.split_terms = {"bf16x3": 3, "bf16x6": 6}.get(precision, 0)
        # VCT_GEMM only selects between the two bf16 implementations; fp32 storage runs the SIMT kernel or the split path
        impl = gemm_impl or (os.environ.get("VCT_GEMM", "tcgen05") if precision == "bf16" else "simt")
        if impl not in ("simt", "tcgen05"):
            raise ValueError("gemm_impl must be 'simt' or 'tcgen05'")
        if impl == "tcgen05" and precision == "fp32":
            raise ValueError("the tcgen05 GEMM takes bf16 operands (precision 'bf16x3' / 'bf16x6' runs fp32 data on it)")
        self.gemm_impl = L.GEMM_TCGEN05 if impl == "tcgen05" else L.GEMM_SIMT
        if self.split_terms:
            self.gemm_impl = L.GEMM_TCGEN05_X3 if self.split_terms == 3 else L.GEMM_TCGEN05_X6
        # dtype of the data-parallel gradient exchange (dense buckets): bf16 halves the NVLink bytes; fp32 storage keeps fp32
        comm = os.environ.get("VCT_GRAD_COMM", "bf16" if precision == "bf16" else "fp32")
        if comm not in ("bf16", "fp32"):
            raise ValueError("VCT_GRAD_COMM must be 'bf16' or 'fp32'")
        self.grad_comm_dtype = BF16 if comm == "bf16" else F32
        self.video_encoder, self.cap_decoder = video_encoder, cap_decoder
        named = []
        if video_encoder is not None:
            named += [("video_encoder." + n, p) for n, p in video_encoder.named_parameters()]
        if cap_decoder is not None:
            named += [("cap_decoder." + n, p) for n, p in cap_decoder.named_parameters()]
        self.arena = ParamArena(named, device)
        # constant sinusoid buffers (state_dict entries of the reference)
        self.pos = cap_decoder.positional_encoding.pos_embedding if cap_decoder is not None else None   # [5000, d]
        self.pe = video_encoder.temp_emb.pe if video_encoder is not None else None                      # [1, 512, d]
        # per-step device state
        # dropout stream: {seed, step}.  Every data-parallel rank draws different masks (the reference's ranks do too:
        # each process owns its own torch RNG stream), so the rank is mixed into the seed.
        self.seed = int(seed)
        self.rng_state = torch.tensor([self._rank_seed(seed), 0], dtype=torch.int64, device=device)
        self._host_step = 0          # last step value handed out on the autograd path (next_rng_step)
        self._dev_step = 0           # value known to be in rng_state[1]; None after a device-side tick
        self._pinned_step = None     # set by MMT4Caption.caption_forward: encoder and decoder share one step
        self.hyper = torch.zeros(8, dtype=torch.float32, device=device)
        self.counters = torch.zeros(1024, dtype=torch.int32, device=device)
        self.upstream = torch.ones(1, dtype=torch.float32, device=device)
        # step stamp per embedding row (vct_embed_mark): which rows of the table the current step touches
        self.emb_stamp = torch.zeros(max(1, int(self.dims.V)), dtype=torch.int32, device=device)
        self.side_streams = [torch.cuda.Stream(device=device) for _ in range(3)] \
            if os.environ.get("VCT_SIDE_STREAM", "1") != "0" else None
        # one grouped persistent launch per layer for the weight-gradient GEMMs (csrc/gemm_tc.cu gemm_tc_grouped_kernel)
        self.group_wgrads = self.gemm_impl == L.GEMM_TCGEN05 and os.environ.get("VCT_GROUP_WGRADS", "1") != "0"
        self.peer = None             # vct.peer.PeerComm once the data-parallel trainer has attached one (attach_peer)
        self._tempo: Dict[int, torch.Tensor] = {}
        self._ws: Dict[Tuple, SimpleNamespace] = {}
        self._shadow_version = None
        self.launches = 0

    # ------------------------------------------------------------------------------------------
    # dropout RNG bookkeeping
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _rank_seed(seed: int) -> int:
        rank = 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank = dist.get_rank()
        except Exception:
            rank = 0
        return (int(seed) + 0x9E3779B97F4A7C15 * rank) & 0x7FFFFFFFFFFFFFFF

    def set_seed(self, seed: int, rank_mix: bool = True) -> None:
        self.seed = int(seed)
        self.rng_state[0:1].fill_(self._rank_seed(seed) if rank_mix else int(seed))

    def set_rng_step(self, step: int) -> None:
        """Make ``step`` the training-step component of every dropout mask drawn from now on (autograd path: the
        backward of a forward must regenerate that forward's masks, whatever ran in between)."""
        if self._dev_step != step:
            self.rng_state[1:2].fill_(int(step))
            self._dev_step = int(step)

    def next_rng_step(self) -> int:
        """A fresh step value for one training forward on the autograd path (the native trainer advances the device
        counter itself through vct_step_tick inside its CUDA graph)."""
        if self._pinned_step is not None:
            return self._pinned_step
        self._host_step += 1
        self.set_rng_step(self._host_step)
        return self._host_step

    # ------------------------------------------------------------------------------------------
    # parameter access
    # ------------------------------------------------------------------------------------------
    def _w(self, name: str, row_off: int = 0) -> int:
        """pointer to a GEMM weight in the compute dtype (bf16 shadow or fp32 master)."""
        a = self.arena
        cols = a.shape[name][1] if len(a.shape[name]) > 1 else 1
        off = a.offset[name] + row_off * cols
        base = a.ensure_shadow() if self.cdt == BF16 else a.p32
        return base.data_ptr() + off * _ESIZE[self.cdt]

    def _p(self, name: str, off: int = 0) -> int:
        return self.arena.p32.data_ptr() + 4 * (self.arena.offset[name] + off)

    def _g(self, name: str, off: int = 0) -> int:
        return self.arena.grad.data_ptr() + 4 * (self.arena.offset[name] + off)

    def refresh_shadow(self, force: bool = False) -> None:
        """Re-derive the bf16 shadow weights if anything outside vct_adam touched the masters
        (torch optimizer step, load_state_dict, manual edits)."""
        if self.cdt != BF16:
            return
        a = self.arena
        ver = a.version()
        if force or a.shadow is None or self._shadow_version != ver:
            sh = a.ensure_shadow()
            L.check(self.lib.vct_cast(a.p32.data_ptr(), sh.data_ptr(), BF16, a.numel, self._stream()), "vct_cast")
            self.launches += 1
            self._shadow_version = ver

    def check_arena(self, quick: bool = False) -> None:
        if not self.arena.is_current(quick):
            raise RuntimeError("model parameters were re-allocated after the vct engine was built "
                               "(.to()/.half()?); rebuild the engine (model.reset_engine())")

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    class _Side:
        def __init__(self, plan, lane): self.plan, self.lane = plan, lane
        def __enter__(self): self.prev, self.plan.lane = self.plan.lane, self.lane
        def __exit__(self, *a): self.plan.lane = self.prev

    def _side(self, plan: Plan, lane: int = 1):
        """``with self._side(plan):`` -- calls added inside run on a side stream (off the critical path).
        lane 1: weight-gradient GEMMs, bias column sums, LayerNorm partial reductions; lane 2: optimizer slices."""
        return CaptionEngine._Side(plan, lane)

    def _adam_slice(self, plan: Plan, first: str, last: str, reduce: bool = True):
        """Optimizer-in-backward: update the arena slice [first .. last] (parameter names, arena order) on lane 2
        as soon as backward has produced its gradients and no longer reads its weights.  Data parallel (plan.allreduce):
        the slice's gradient is first SUM all-reduced over NCCL on the same lane -- in bf16 when the engine computes in
        bf16 (VCT_GRAD_COMM=fp32 keeps the exchange in fp32): the fp32 slice is cast into the bf16 gradient arena, that is
        all-reduced (half the bytes over NVLink) and vct_adam reads the bf16 result.  reduce=False: the caller has already
        made the slice's gradient global (sparse exchange of the embedding rows)."""
        if not getattr(plan, "fuse_adam", False):
            return
        a = self.arena
        lo = a.offset[first]
        i = a.names.index(last)
        hi = a.offset[a.names[i + 1]] if i + 1 < len(a.names) else a.numel
        a.ensure_optimizer_state()
        shadow = a.ensure_shadow().data_ptr() + 2 * lo if self.cdt == BF16 else None
        grad_scale = 1.0
        g_ptr, g_dt = a.grad.data_ptr() + 4 * lo, F32
        with self._side(plan, 2):
            ar = getattr(plan, "allreduce", None)
            if ar is not None:
                # data parallel: 1/world is folded into the Adam kernel (DDP averages, train.py:218)
                group, world = ar
                grad_scale = 1.0 / world
                if reduce and self.peer is not None:
                    # NVLink peer-memory exchange: cast the slice into the communication region, two-shot all-reduce
                    # kernel (graph-capturable), Adam reads the bf16 sums
                    g16 = a.ensure_grad16()
                    n8 = (hi - lo + 7) // 8 * 8
                    plan.add(f"vct_cast:{first}..{last}", self.lib.vct_cast, a.grad.data_ptr() + 4 * lo, g16.data_ptr() + 2 * lo,
                             BF16, hi - lo)
                    plan.add(f"vct_peer_allreduce:{first}..{last}", self.lib.vct_peer_allreduce_bf16, self.peer.handle,
                             self._peer_g16_off + 2 * lo, n8, self._peer_stage_off, 0)
                    g_ptr, g_dt = g16.data_ptr() + 2 * lo, BF16
                elif reduce:
                    if self.grad_comm_dtype == BF16:
                        g16 = a.ensure_grad16()
                        plan.add(f"vct_cast:{first}..{last}", self.lib.vct_cast, a.grad.data_ptr() + 4 * lo, g16.data_ptr() + 2 * lo,
                                 BF16, hi - lo)
                        flat = g16[lo:hi]
                        g_ptr, g_dt = g16.data_ptr() + 2 * lo, BF16
                    else:
                        flat = a.grad[lo:hi]

                    def all_reduce(stream, flat=flat, group=group):
                        import torch.distributed as dist
                        with torch.cuda.stream(stream):
                            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
                        return 0
                    plan.add(f"py:all_reduce:{first}..{last}", all_reduce)
            plan.add(f"vct_adam:{first}..{last}", self.lib.vct_adam, a.p32.data_ptr() + 4 * lo, g_ptr, g_dt,
                     a.exp_avg.data_ptr() + 4 * lo, a.exp_avg_sq.data_ptr() + 4 * lo, shadow, hi - lo,
                     self.hyper.data_ptr(), grad_scale)
        plan.adam_covered = getattr(plan, "adam_covered", 0) + (hi - lo)

    PEER_MAX_TOKENS = 16384          # tokens of all ranks the sparse embedding exchange covers (vct_embed_sort limit)

    @staticmethod
    def peer_region_bytes(numel: int, d: int) -> int:
        """Size of one rank's communication region: bf16 gradient arena + gathered embedding rows + gathered ids."""
        r = lambda v: (v + 255) // 256 * 256
        return 2 * r(numel * 2 + 4096) + r(CaptionEngine.PEER_MAX_TOKENS * d * 4) + r(2 * CaptionEngine.PEER_MAX_TOKENS * 8) + 4096

    def attach_peer(self, comm) -> None:
        """Route the data-parallel gradient exchange through the NVLink peer-memory kernels (csrc/peer_comm.cu): the bf16
        gradient arena, the gathered embedding rows and the gathered ids live in the communication region every rank maps."""
        if self.grad_comm_dtype != BF16:
            raise RuntimeError("the peer-memory exchange carries bf16 gradients (VCT_GRAD_COMM=bf16)")
        a = self.arena
        self.peer = comm
        self._peer_g16_off = comm.reserve(a.numel * 2 + 256)
        a.grad16 = comm.tensor(self._peer_g16_off, (a.numel,), torch.bfloat16)
        a.grad16.zero_()
        self._peer_stage_off = comm.reserve(a.numel * 2 + 4096)     # staging of the two-shot all-reduce (any slice fits)
        self._peer_rows_off = comm.reserve(self.PEER_MAX_TOKENS * self.dims.d * 4)
        self._peer_ids_off = comm.reserve(2 * self.PEER_MAX_TOKENS * 8)
        for ws in self._ws.values():                      # plans built before the attachment hold stale pointers
            if hasattr(ws, "plans"):
                ws.plans.clear()
                getattr(ws, "graphs", {}).clear()
                getattr(ws, "scratch", {}).pop(("emb.buffers", comm.world), None)

    def emb_mode(self, ws, world: int) -> str:
        """How the native trainer treats the embedding-table gradient of this workspace:
        'local'  one GPU: atomic scatter, Adam split into untouched rows (start of the step) and touched rows (end)
        'sparse' data parallel: ids gathered and sorted at the start of the step, gradient rows all-gathered and summed by
                 the deterministic segment kernel, same Adam split
        'dense'  data parallel with more tokens / a larger vocabulary than the one-CTA sort covers: the dense table gradient
                 is all-reduced like every other bucket"""
        if os.environ.get("VCT_SPLIT_EMB_ADAM", "1") == "0":
            return "dense" if world > 1 else "local-unsplit"
        if world == 1:
            return "local"
        # the sparse exchange moves [B*S, d] rows and [B, S+1] ids per rank: every rank must present the SAME (B, S) in a step.
        # The trainer only enables it when its caller promises that (CaptionTrainer(uniform_shapes=True): synthetic batches,
        # loaders that pad captions to a fixed length); otherwise -- ragged caption lengths differ from rank to rank -- the
        # dense table gradient is exchanged, which is shape-independent like every other slice.
        ok = world * ws.B * ws.S <= self.PEER_MAX_TOKENS and self.dims.V <= 32768 and os.environ.get("VCT_SPARSE_EMB", "1") != "0" \
            and getattr(self, "uniform_shapes", False)
        if self.peer is not None and (ws.B * (ws.S + 1)) % 2:
            ok = False                                    # the peer all-gather moves 16-byte vectors: ids slots must be even
        return "sparse" if ok else "dense"

    def emb_buffers(self, ws, world: int):
        """(all_ids [world*B, S+1] int64, all_rows [world*B*S, d] fp32, rows [B*S, d] fp32, keys int32 [world*B*S])."""
        key = ("emb.buffers", world)
        if key not in ws.scratch:
            Rd = ws.B * ws.S
            keys = torch.zeros(world * Rd, dtype=torch.int32, device=self.device)
            if self.peer is not None and world > 1:
                # views of the communication region: this rank's rows / ids are written straight into its own slot
                r = self.peer.rank
                all_ids = self.peer.tensor(self._peer_ids_off, (world * ws.B, ws.S + 1), torch.int64)
                all_rows = self.peer.tensor(self._peer_rows_off, (world * Rd, self.dims.d), torch.float32)
                all_ids.zero_()
                ws.scratch[key] = (all_ids, all_rows, all_rows[r * Rd:(r + 1) * Rd], keys)
            else:
                ws.scratch[key] = (torch.zeros((world * ws.B, ws.S + 1), dtype=torch.int64, device=self.device),
                                   torch.empty((world * Rd, self.dims.d), dtype=torch.float32, device=self.device),
                                   torch.empty((Rd, self.dims.d), dtype=torch.float32, device=self.device), keys)
        return ws.scratch[key]

    def _emb_range(self):
        a = self.arena
        emb = "cap_decoder.tgt_to_emb.weight"
        lo = a.offset[emb]
        i = a.names.index(emb)
        hi = a.offset[a.names[i + 1]] if i + 1 < len(a.names) else a.numel
        return emb, lo, hi

    def plan_embed_early(self, ws, world: int) -> Plan:
        """Start-of-step half of the embedding update (optimizer lane, left un-joined): stamp the rows this step touches
        (for world > 1 from the all-gathered ids, which the trainer has placed in emb_buffers()[0]), sort the gathered
        tokens for the deterministic scatter, and run Adam on every row the step does NOT touch."""
        key = ("embed_early", world)
        if key not in ws.plans:
            D, lib, a = self.dims, self.lib, self.arena
            emb, lo, _hi = self._emb_range()
            a.ensure_optimizer_state()
            p = Plan()
            p.ws = ws
            p.no_final_join = True
            ids_ptr, nB = ws.ids.data_ptr(), ws.B
            if world > 1:
                all_ids, _, _, keys = self.emb_buffers(ws, world)
                ids_ptr, nB = all_ids.data_ptr(), world * ws.B
            shadow = a.ensure_shadow().data_ptr() + 2 * lo if self.cdt == BF16 else None
            with self._side(p, 2):
                if world > 1 and self.peer is not None:
                    # token ids of every rank (the trainer has staged this rank's ids into its slot of the region)
                    p.add("vct_peer_allgather:ids", lib.vct_peer_allgather, self.peer.handle, self._peer_ids_off,
                          ws.B * (ws.S + 1) * 8, 0)
                p.add("vct_embed_mark", lib.vct_embed_mark, ids_ptr, ws.S + 1, nB, ws.S, D.V, D.pad_id, self.emb_stamp.data_ptr(),
                      self.rng_state.data_ptr())
                if world > 1:
                    p.add("vct_embed_sort", lib.vct_embed_sort, ids_ptr, ws.S + 1, nB, ws.S, D.V, D.pad_id, keys.data_ptr())
                p.add("vct_adam_rows:untouched", lib.vct_adam_rows, a.p32.data_ptr() + 4 * lo, None, a.exp_avg.data_ptr() + 4 * lo,
                      a.exp_avg_sq.data_ptr() + 4 * lo, shadow, D.V, D.d, self.hyper.data_ptr(), 1.0, self.emb_stamp.data_ptr(),
                      self.rng_state.data_ptr(), 0)
            ws.plans[key] = p
        return ws.plans[key]

    def _scratch(self, ws, tag: str, rows: int, cols: int, dtype) -> torch.Tensor:
        """Per-use gradient scratch (never shared between call sites, so side-stream readers cannot race
        with later writers on the main stream)."""
        if tag not in ws.scratch:
            ws.scratch[tag] = torch.empty((rows, cols), dtype=dtype, device=self.device)
        return ws.scratch[tag]

    def tempo_table(self, T: int) -> torch.Tensor:
        """[T+1, d] temporal-encoding rows: row 0 zeros (global token), row i = pe[i-1]
        (model/MMEncoder.py:89-104 for one modality: indices = linspace(0, T-1, T) = 0..T-1).
        Constant per T, so it is built once instead of per forward."""
        if T not in self._tempo:
            if T > self.pe.shape[1]:
                raise ValueError(f"T={T} exceeds the temporal table ({self.pe.shape[1]})")
            t = torch.zeros(T + 1, self.dims.d, dtype=torch.float32, device=self.device)
            t[1:] = self.pe[0, :T].to(self.device, torch.float32)
            self._tempo[T] = t
        return self._tempo[T]

    # ------------------------------------------------------------------------------------------
    # small call builders
    # ------------------------------------------------------------------------------------------
    def _gemm(self, plan: Plan, tag: str, M, N, K, A, lda, a_trans, B, ldb, b_trans, Cp, c_dtype, ldc, *, bias=None,
              C2=None, c2_dtype=F32, ldc2=0, row_table=None, row_period=0, addend=None, ld_addend=0, act=L.ACT_NONE,
              aux=None, ld_aux=0, drop_p=0.0, site=0, in_dtype=None, defer=False):
        """defer: a weight-gradient GEMM that only has to be finished by the end of its layer's backward -- collected
        and issued by _flush_group as ONE grouped persistent launch (bf16 tcgen05 engines; issued at once otherwise)."""
        g = L.GemmArgs()
        g.M, g.N, g.K = M, N, K
        dt = self.cdt if in_dtype is None else in_dtype
        g.A, g.a_dtype, g.lda, g.a_trans = A, dt, lda, a_trans
        g.B, g.b_dtype, g.ldb, g.b_trans = B, dt, ldb, b_trans
        g.C, g.c_dtype, g.ldc = Cp, c_dtype, ldc
        g.C2, g.c2_dtype, g.ldc2 = C2, c2_dtype, ldc2
        g.bias = bias
        g.row_table, g.row_period = row_table, row_period
        g.addend, g.ld_addend = addend, ld_addend
        g.act = act
        g.aux, g.aux_dtype, g.ld_aux = aux, self.cdt, ld_aux
        g.drop_p, g.rng_state, g.site = drop_p, self.rng_state.data_ptr(), site
        g.impl = self.gemm_impl
        ws = getattr(plan, "ws", None)
        if ws is not None and getattr(ws, "splitk", None) is not None:
            sk = ws.splitk[1 if plan.lane else 0]           # one workspace per lane: the lanes run concurrently
            g.splitk_ws, g.splitk_ws_floats = sk.data_ptr(), sk.numel()
        if self.split_terms:
            g.split_ws, g.split_ws_bytes = self._split_ws(ws, 1 if plan.lane else 0)
        plan.keep.append(g)
        if defer and self.group_wgrads:
            plan.__dict__.setdefault("_group", []).append((tag, g))
            return
        plan.add("vct_gemm:" + tag, self.lib.vct_gemm, C.byref(g))

    def _flush_group(self, plan: Plan, tag: str):
        """Issue the deferred weight-gradient GEMMs of the layer just finished as one vct_gemm_grouped call (side lane)."""
        group = plan.__dict__.get("_group") or []
        if not group:
            return
        plan._group = []
        for i in range(0, len(group), 8):
            chunk = group[i:i + 8]
            arr = (L.GemmArgs * len(chunk))()
            for k, (_t, g) in enumerate(chunk):
                C.memmove(C.byref(arr, k * C.sizeof(L.GemmArgs)), C.byref(g), C.sizeof(L.GemmArgs))
            plan.keep.append(arr)
            with self._side(plan):
                plan.add(f"vct_gemm_grouped:{tag}" + (f".{i // 8}" if len(group) > 8 else ""), self.lib.vct_gemm_grouped, arr, len(chunk))

    def _split_ws(self, ws, lane: int):
        """(pointer, bytes) of the lane's operand-split scratch (precision bf16x3 / bf16x6)."""
        t = ws.split_ws[lane]
        return t.data_ptr(), t.numel()

    def _alloc_split_ws(self, ws, rows_max: int, with_grad: bool):
        """Scratch for the bf16 pieces of both operands of the largest GEMM a plan over ``ws`` issues, one per lane."""
        if not self.split_terms:
            ws.split_ws = None
            return
        D, t = self.dims, self.split_terms
        Fm = max(D.F_enc, D.F_dec)
        shapes = [(rows_max, D.V, D.d, 0, 0), (rows_max, 3 * D.d, D.d, 0, 0), (rows_max, Fm, D.d, 0, 0),
                  (rows_max, D.d, Fm, 0, 0), (rows_max, D.d, D.Din, 0, 0)]
        if with_grad:
            shapes += [(rows_max, D.d, D.V, 0, 1), (D.V, D.d, rows_max, 1, 1), (rows_max, Fm, D.d, 0, 1), (rows_max, D.d, Fm, 0, 1),
                       (Fm, D.d, rows_max, 1, 1), (D.d, Fm, rows_max, 1, 1), (3 * D.d, D.d, rows_max, 1, 1),
                       (rows_max, D.d, 3 * D.d, 0, 1), (D.d, D.Din, rows_max, 1, 1)]
        need = max(int(self.lib.vct_gemm_split_workspace_bytes(M, N, K, at, bt, t)) for M, N, K, at, bt in shapes)
        ws.split_ws = [torch.empty(need + 256, dtype=torch.uint8, device=self.device) for _ in range(2 if with_grad else 1)]
        if not with_grad:
            ws.split_ws.append(ws.split_ws[0])

    def _colsum(self, plan: Plan, tag: str, X, ld, M, N, out, ws):
        # column sums run on the side lane: they get their own partials buffer (LN backward uses ws.partials)
        plan.add("vct_colsum:" + tag, self.lib.vct_colsum, X, self.cdt, ld, M, N, out, ws.partials_side.data_ptr(),
                 self.counters.data_ptr() + 4 * 8)

    def _ln_fwd(self, plan: Plan, tag, x, r, gname, bname, y, y_c, s_out, mean, rstd, R, p, site):
        plan.add("vct_ln_residual_fwd:" + tag, self.lib.vct_ln_residual_fwd, x, r, self._p(gname), self._p(bname), y, y_c,
                 self.cdt, s_out, mean, rstd, R, self.dims.d, p, self.rng_state.data_ptr(), site)

    def _ln_bwd(self, plan: Plan, tag, dy, s, mean, rstd, gname, bname, ds, dr_c, dbias, R, p, site, ws):
        # main lane: ds / dr and the per-CTA column partials; side lane: the partials -> dgamma, dbeta, dbias
        # reduction, which nothing before Adam consumes (own partials buffer per call site)
        n = int(self.lib.vct_ln_bwd_workspace_floats(R, self.dims.d))
        part = self._scratch(ws, "ln_partials:" + tag, 1, n, torch.float32).data_ptr()
        plan.add("vct_ln_residual_bwd:" + tag, self.lib.vct_ln_residual_bwd, dy, s, mean, rstd, self._p(gname), ds, dr_c,
                 self.cdt, None, None, None, part, None, R, self.dims.d, p, self.rng_state.data_ptr(), site)
        with self._side(plan):
            plan.add("vct_ln_bwd_reduce:" + tag, self.lib.vct_ln_bwd_reduce, part, R, self.dims.d, self._g(gname),
                     self._g(bname), dbias)

    def _attn(self, plan: Plan, tag, bwd, *, B, H, Lq, Lk, q, q_ld, k, k_ld, v, v_ld, o, o_ld, key_pad=None, causal=0,
              p=0.0, site=0, probs=None, d_o=None, do_ld=0, dq=None, dq_ld=0, dk=None, dk_ld=0, dv=None, dv_ld=0,
              q_bs=0, k_bs=0, v_bs=0, o_bs=0, dbias=None, ws=None):
        a = L.AttnArgs()
        a.B, a.H, a.Lq, a.Lk, a.dh = B, H, Lq, Lk, self.dims.d // H
        a.dtype = self.cdt
        a.q, a.q_ld, a.k, a.k_ld, a.v, a.v_ld, a.o, a.o_ld = q, q_ld, k, k_ld, v, v_ld, o, o_ld
        a.key_pad, a.causal, a.scale = key_pad, causal, 1.0 / math.sqrt(self.dims.d // H)
        a.drop_p, a.rng_state, a.site = p, self.rng_state.data_ptr(), site
        a.probs = probs
        a.d_o, a.do_ld, a.dq, a.dq_ld, a.dk, a.dk_ld, a.dv, a.dv_ld = d_o, do_ld, dq, dq_ld, dk, dk_ld, dv, dv_ld
        a.q_bs, a.k_bs, a.v_bs, a.o_bs = q_bs, k_bs, v_bs, o_bs
        long_seq = max(Lq, Lk) > 64          # tiled kernels (csrc/attn_core.cu "Long sequences")
        if dbias is not None and not long_seq:
            # in-projection bias gradient (column sums of dq | dk | dv) produced by the backward kernel itself; the
            # partials buffer is per call site, the counters are self-resetting
            part = self._scratch(ws, "attn_dbias:" + tag, B, 3 * self.dims.d, torch.float32)
            a.dbias, a.dbias_partials, a.dbias_counters = dbias, part.data_ptr(), self.counters.data_ptr() + 4 * 512
        if bwd and long_seq:
            # per-row softmax statistics handed from the query-tile kernel to the key-tile kernel
            a.row_stats = self._scratch(ws, "attn_row_stats:" + tag, B * H * Lq, 4, torch.float32).data_ptr()
        plan.keep.append(a)
        plan.add(("vct_attn_bwd:" if bwd else "vct_attn_fwd:") + tag,
                 self.lib.vct_attn_bwd if bwd else self.lib.vct_attn_fwd, C.byref(a))
        if bwd and long_seq and dbias is not None:
            # the tiled backward leaves the in-projection bias gradient to column sums over dq | dk | dv (same lane as the
            # attention call; own partials and counters, so the side lane's column sums cannot interfere)
            d = self.dims.d
            n = int(self.lib.vct_colsum_workspace_floats(B * max(Lq, Lk), d))
            part = self._scratch(ws, "attn_dbias_long:" + tag, 1, n, torch.float32).data_ptr()
            cnt = self.counters.data_ptr() + 4 * 768
            for sec, (X, ld, rows) in enumerate(((dq, dq_ld, B * Lq), (dk, dk_ld, B * Lk), (dv, dv_ld, B * Lk))):
                plan.add(f"vct_colsum:{tag}.in_proj_bias[{sec}]", self.lib.vct_colsum, X, self.cdt, ld, rows, d,
                         dbias + 4 * sec * d, part, cnt)

    # ------------------------------------------------------------------------------------------
    # workspaces
    # ------------------------------------------------------------------------------------------
    def _buf(self, ws, name, shape, dtype):
        t = torch.empty(shape, dtype=dtype, device=self.device)
        setattr(ws, name, t)
        return t

    def workspace(self, B: int, T: int, S: int, training: bool) -> SimpleNamespace:
        key = (B, T, S, training)
        if key in self._ws:
            ws = self._ws.pop(key)
            self._ws[key] = ws            # most recently used last
            return ws
        self._check_lengths(T + 1, S)
        self._evict_workspaces()
        D = self.dims
        ws = SimpleNamespace(B=B, T=T, S=S, M=T + 1, training=training)
        d, M = D.d, T + 1
        Re, Rd = B * M, B * S
        cdt = _TDT[self.cdt]
        f32 = torch.float32
        two = self.cdt == BF16      # separate compute-dtype copies of fp32 activations?
        ws.Vp = (D.V + 7) // 8 * 8

        def act(name, rows, cols):
            """fp32 activation + (in bf16 mode) its bf16 twin; returns (fp32 tensor, compute-dtype ptr)."""
            t = self._buf(ws, name, (rows, cols), f32)
            c = self._buf(ws, name + "_c", (rows, cols), cdt) if two else t
            setattr(ws, name + "_c", c)
            return t

        # inputs (device-resident staging so that plans have static pointers)
        self._buf(ws, "feats", (B, T, D.Din), f32)
        self._buf(ws, "vid_pad", (B, M), torch.uint8)
        self._buf(ws, "ids", (B, S + 1), torch.int64)
        self._buf(ws, "tok_pad", (B, S), torch.uint8)
        # encoder
        self._buf(ws, "a0", (Re, D.Din), cdt)
        act("x0", Re, d)
        ws.enc = []
        for l in range(D.L_enc):
            e = SimpleNamespace()
            e.qkv = torch.empty((Re, 3 * d), dtype=cdt, device=self.device)
            e.ao = torch.empty((Re, d), dtype=cdt, device=self.device)
            e.s1 = torch.empty((Re, d), dtype=f32, device=self.device)
            e.x1 = torch.empty((Re, d), dtype=f32, device=self.device)
            e.x1_c = torch.empty((Re, d), dtype=cdt, device=self.device) if two else e.x1
            e.z = torch.empty((Re, D.F_enc), dtype=cdt, device=self.device)
            e.h = torch.empty((Re, D.F_enc), dtype=cdt, device=self.device)
            e.s2 = torch.empty((Re, d), dtype=f32, device=self.device)
            e.x2 = torch.empty((Re, d), dtype=f32, device=self.device)
            e.x2_c = torch.empty((Re, d), dtype=cdt, device=self.device) if two else e.x2
            e.stats = torch.empty((4, Re), dtype=f32, device=self.device)   # mean1 rstd1 mean2 rstd2
            ws.enc.append(e)
        act("mem", Re, d)
        self._buf(ws, "mem_stats", (2, Re), f32)
        # decoder
        act("e0", Rd, d)
        ws.dec = []
        for l in range(D.L_dec):
            e = SimpleNamespace()
            e.qkv = torch.empty((Rd, 3 * d), dtype=cdt, device=self.device)
            e.ao = torch.empty((Rd, d), dtype=cdt, device=self.device)
            e.s1 = torch.empty((Rd, d), dtype=f32, device=self.device)
            e.x1 = torch.empty((Rd, d), dtype=f32, device=self.device)
            e.x1_c = torch.empty((Rd, d), dtype=cdt, device=self.device) if two else e.x1
            e.q = torch.empty((Rd, d), dtype=cdt, device=self.device)
            e.kv = torch.empty((Re, 2 * d), dtype=cdt, device=self.device)
            e.ao2 = torch.empty((Rd, d), dtype=cdt, device=self.device)
            e.s2 = torch.empty((Rd, d), dtype=f32, device=self.device)
            e.x2 = torch.empty((Rd, d), dtype=f32, device=self.device)
            e.x2_c = torch.empty((Rd, d), dtype=cdt, device=self.device) if two else e.x2
            e.z = torch.empty((Rd, D.F_dec), dtype=cdt, device=self.device)
            e.h = torch.empty((Rd, D.F_dec), dtype=cdt, device=self.device)
            e.s3 = torch.empty((Rd, d), dtype=f32, device=self.device)
            e.x3 = torch.empty((Rd, d), dtype=f32, device=self.device)
            e.x3_c = torch.empty((Rd, d), dtype=cdt, device=self.device) if two else e.x3
            e.stats = torch.empty((6, Rd), dtype=f32, device=self.device)
            ws.dec.append(e)
        act("hfin", Rd, d)
        self._buf(ws, "hfin_stats", (2, Rd), f32)
        # teacher-forced logits: the training plans keep them in bf16 when the engine computes in bf16 (the caption loss is
        # all the reference keeps of them, model/MMT4Caption.py:120-121; VCT_LOGITS=fp32 restores the fp32 buffer); eval
        # workspaces stay fp32 (decode_word / argmax read them)
        ws.logits_dt = BF16 if (self.cdt == BF16 and training and os.environ.get("VCT_LOGITS", "bf16") != "fp32") else F32
        self._buf(ws, "logits", (Rd, ws.Vp), _TDT[ws.logits_dt])
        self._buf(ws, "loss", (1,), f32)
        self._buf(ws, "row_parts", (Rd, 2), f32)
        if training:
            # zero-filled once: columns V..Vp-1 stay zero, they are K-padding of the generator dgrad
            ws.dlogits = torch.zeros((Rd, ws.Vp), dtype=cdt, device=self.device)
            Fm = max(D.F_enc, D.F_dec)
            rmax = max(Re, Rd)
            self._buf(ws, "g_a", (rmax, d), f32)          # fp32 activation-gradient ping/pong
            self._buf(ws, "g_b", (rmax, d), f32)
            self._buf(ws, "g_s", (rmax, d), f32)          # ds of the LN being processed
            self._buf(ws, "g_o_c", (rmax, d), cdt)        # gradient wrt attention output
            self._buf(ws, "g_mem", (Re, d), f32)
            self._buf(ws, "g_x0_c", (Re, d), cdt)
            nws = max(int(self.lib.vct_ln_bwd_workspace_floats(rmax, d)),
                      int(self.lib.vct_colsum_workspace_floats(rmax, max(3 * d, Fm))),
                      int(self.lib.vct_colsum_workspace_floats(Rd, D.V)))
            self._buf(ws, "partials", (nws,), f32)
            self._buf(ws, "partials_side", (nws,), f32)
            ws.scratch = {}
        # split-K workspaces (fp32 partial tiles) for long-K GEMMs, one per lane
        ws.splitk = [torch.empty(8 * max(Re, Rd) * max(d, 8), dtype=f32, device=self.device) for _ in range(2)] \
            if self.gemm_impl != L.GEMM_SIMT else None
        # (the eval encoder / cross-K/V GEMMs of a forward-only workspace also run on the side lane: two buffers always)
        self._alloc_split_ws(ws, max(Re, Rd), True)
        ws.plans = {}
        ws.graphs = {}            # CUDA graphs captured over this workspace's pointers (vct.trainer): same lifetime
        ws.enc_version = 0        # bumped by every encoder / decoder forward over this workspace: backward checks that
        ws.dec_version = 0        # the activations it is about to read are still those of its own forward
        ws.mem_token = None
        self._ws[key] = ws
        return ws

    MAX_LEN = 1024      # longest sequence the attention kernels cover (memory rows T + 1, decoder positions S)

    def _check_lengths(self, M: int, S: int) -> None:
        """Sequences up to 64 rows (every shipped config: 13 / 33 memory rows, captions below 40 word pieces) run on the
        fused tcgen05 attention kernels, which hold a whole (batch, head) in one tile; longer ones -- the reference
        truncates neither captions nor frames for training -- run the tiled SIMT kernels of csrc/attn_core.cu, whose
        score tiles fit shared memory up to 1024 rows (the reference's own tables end at 512 frames,
        model/MMEncoder.py:65, and 5000 positions, model/Embedding.py:11).  Beyond that: fail BEFORE the first launch,
        with the remedy, instead of in the middle of an epoch."""
        if M > self.MAX_LEN or S > self.MAX_LEN:
            raise ValueError(
                f"vct_b200: sequence too long for the attention kernels (memory length T+1 = {M}, decoder positions = {S}; "
                f"limit {self.MAX_LEN}).  Sample at most {self.MAX_LEN - 1} frames per video and truncate captions to "
                f"{self.MAX_LEN} word pieces (+1 for the shifted target) in the data pipeline (see INTEGRATION.md).")

    def _evict_workspaces(self) -> None:
        """Bound the per-shape workspace cache (batches built by caption length have a new S almost every step): keep the
        VCT_WS_CACHE most recently used training / eval workspaces.  An evicted workspace (with its plans and CUDA
        graphs) is freed once nothing else -- e.g. a pending autograd graph -- references it."""
        limit = max(1, int(os.environ.get("VCT_WS_CACHE", "6")))
        keys = [k for k in self._ws if k and k[0] != "decode"]
        while len(keys) >= limit:
            self._ws.pop(keys.pop(0))

    # ------------------------------------------------------------------------------------------
    # forward plan
    # ------------------------------------------------------------------------------------------
    def _build_encoder(self, plan: Plan, ws, p_drop: float, zero_pad: bool = False):
        D, lib = self.dims, self.lib
        d, B, T, M = D.d, ws.B, ws.T, ws.M
        Re = B * M
        cd = self.cdt
        plan.add("vct_prep_frames", lib.vct_prep_frames, ws.feats.data_ptr(), ws.a0.data_ptr(), cd, B, T, D.Din)
        tempo = self.tempo_table(T)
        plan.keep.append(tempo)
        self._gemm(plan, "unify", Re, d, D.Din, ws.a0.data_ptr(), D.Din, 0, self._w("video_encoder.unify.0.weight"),
                   D.Din, 0, ws.x0.data_ptr(), F32, d, bias=self._p("video_encoder.unify.0.bias"),
                   C2=ws.x0_c.data_ptr() if cd == BF16 else None, c2_dtype=cd, ldc2=d,
                   row_table=tempo.data_ptr(), row_period=M)
        x, x_c = ws.x0, ws.x0_c
        for l, e in enumerate(ws.enc):
            pre = f"video_encoder.transformer_encoder.layers.{l}."
            m = L.MhaArgs()
            m.B, m.L, m.Lk, m.d, m.H, m.dtype = B, M, M, d, D.H_enc, cd
            m.x, m.w_in, m.b_in = x_c.data_ptr(), self._w(pre + "self_attn.in_proj_weight"), self._p(pre + "self_attn.in_proj_bias")
            m.qkv, m.o, m.key_pad = e.qkv.data_ptr(), e.ao.data_ptr(), ws.vid_pad.data_ptr()
            m.drop_p, m.rng_state, m.site = p_drop, self.rng_state.data_ptr(), _enc_site(l, 0)
            m.gemm_impl = self.gemm_impl
            if self.split_terms:
                m.split_ws, m.split_ws_bytes = self._split_ws(ws, 1 if plan.lane else 0)
            plan.keep.append(m)
            plan.add(f"vct_attn_enc_self_fwd:{l}", lib.vct_attn_enc_self_fwd, C.byref(m))
            self._gemm(plan, f"enc{l}.out_proj", Re, d, d, e.ao.data_ptr(), d, 0, self._w(pre + "self_attn.out_proj.weight"),
                       d, 0, e.s1.data_ptr(), F32, d, bias=self._p(pre + "self_attn.out_proj.bias"))
            self._ln_fwd(plan, f"enc{l}.norm1", x.data_ptr(), e.s1.data_ptr(), pre + "norm1.weight", pre + "norm1.bias",
                         e.x1.data_ptr(), e.x1_c.data_ptr() if cd == BF16 else None, e.s1.data_ptr(),
                         e.stats[0].data_ptr(), e.stats[1].data_ptr(), Re, p_drop, _enc_site(l, 1))
            self._gemm(plan, f"enc{l}.linear1", Re, D.F_enc, d, e.x1_c.data_ptr(), d, 0, self._w(pre + "linear1.weight"), d, 0,
                       e.z.data_ptr(), cd, D.F_enc, bias=self._p(pre + "linear1.bias"), C2=e.h.data_ptr(), c2_dtype=cd,
                       ldc2=D.F_enc, act=L.ACT_GELU_FWD_F if ws.training else L.ACT_GELU_FWD, drop_p=p_drop, site=_enc_site(l, 2))
            self._gemm(plan, f"enc{l}.linear2", Re, d, D.F_enc, e.h.data_ptr(), D.F_enc, 0, self._w(pre + "linear2.weight"),
                       D.F_enc, 0, e.s2.data_ptr(), F32, d, bias=self._p(pre + "linear2.bias"))
            self._ln_fwd(plan, f"enc{l}.norm2", e.x1.data_ptr(), e.s2.data_ptr(), pre + "norm2.weight", pre + "norm2.bias",
                         e.x2.data_ptr(), e.x2_c.data_ptr() if cd == BF16 else None, e.s2.data_ptr(),
                         e.stats[2].data_ptr(), e.stats[3].data_ptr(), Re, p_drop, _enc_site(l, 3))
            x, x_c = e.x2, e.x2_c
        ws.enc_out = x
        if zero_pad:
            # eval() + no_grad + key-padding mask: torch's nested-tensor fast path zero-fills padded positions before the
            # final norm -- reproduced so validation loss / greedy captions of padded batches match the
            # reference as it is actually run (train.py:151-168, eval.py:140)
            plan.add("vct_zero_rows", lib.vct_zero_rows, x.data_ptr(), ws.vid_pad.data_ptr(), Re, d)
        self._ln_fwd(plan, "enc.norm", None, x.data_ptr(), "video_encoder.transformer_encoder.norm.weight",
                     "video_encoder.transformer_encoder.norm.bias", ws.mem.data_ptr(),
                     ws.mem_c.data_ptr() if cd == BF16 else None, None, ws.mem_stats[0].data_ptr(),
                     ws.mem_stats[1].data_ptr(), Re, 0.0, 0)

    def _build_cross_kv(self, plan: Plan, ws):
        """K/V projections of the encoder memory for every decoder layer (rows [d:3d] of multihead_attn.in_proj)."""
        D = self.dims
        d = D.d
        for l, e in enumerate(ws.dec):
            pre = f"cap_decoder.decoder.layers.{l}.multihead_attn."
            self._gemm(plan, f"dec{l}.cross.kv", ws.B * ws.M, 2 * d, d, ws.mem_c.data_ptr(), d, 0,
                       self._w(pre + "in_proj_weight", d), d, 0, e.kv.data_ptr(), self.cdt, 2 * d,
                       bias=self._p(pre + "in_proj_bias", d))

    def _build_decoder(self, plan: Plan, ws, p_drop: float, with_loss: bool, with_grad: bool, kv_ready: bool = False):
        D, lib = self.dims, self.lib
        d, B, S, M = D.d, ws.B, ws.S, ws.M
        Re, Rd = B * M, B * S
        cd = self.cdt
        plan.add("vct_embed_fwd", lib.vct_embed_fwd, ws.ids.data_ptr(), S + 1, self._p("cap_decoder.tgt_to_emb.weight"),
                 self.pos.data_ptr(), ws.e0.data_ptr(), ws.e0_c.data_ptr() if cd == BF16 else None, cd, B, S, d, D.V, 0,
                 p_drop, self.rng_state.data_ptr(), SITE_EMBED)
        x, x_c = ws.e0, ws.e0_c
        for l, e in enumerate(ws.dec):
            pre = f"cap_decoder.decoder.layers.{l}."
            m = L.MhaArgs()
            m.B, m.L, m.Lk, m.d, m.H, m.dtype = B, S, S, d, D.H_dec, cd
            m.x, m.w_in, m.b_in = x_c.data_ptr(), self._w(pre + "self_attn.in_proj_weight"), self._p(pre + "self_attn.in_proj_bias")
            m.qkv, m.o, m.key_pad = e.qkv.data_ptr(), e.ao.data_ptr(), ws.tok_pad.data_ptr()
            m.drop_p, m.rng_state, m.site = p_drop, self.rng_state.data_ptr(), _dec_site(l, 0)
            m.gemm_impl = self.gemm_impl
            if self.split_terms:
                m.split_ws, m.split_ws_bytes = self._split_ws(ws, 1 if plan.lane else 0)
            plan.keep.append(m)
            plan.add(f"vct_attn_dec_self_fwd:{l}", lib.vct_attn_dec_self_fwd, C.byref(m))
            self._gemm(plan, f"dec{l}.self.out_proj", Rd, d, d, e.ao.data_ptr(), d, 0, self._w(pre + "self_attn.out_proj.weight"),
                       d, 0, e.s1.data_ptr(), F32, d, bias=self._p(pre + "self_attn.out_proj.bias"))
            self._ln_fwd(plan, f"dec{l}.norm1", x.data_ptr(), e.s1.data_ptr(), pre + "norm1.weight", pre + "norm1.bias",
                         e.x1.data_ptr(), e.x1_c.data_ptr() if cd == BF16 else None, e.s1.data_ptr(),
                         e.stats[0].data_ptr(), e.stats[1].data_ptr(), Rd, p_drop, _dec_site(l, 1))
            c = L.MhaArgs()
            c.B, c.L, c.Lk, c.d, c.H, c.dtype = B, S, M, d, D.H_dec, cd
            c.x, c.mem = e.x1_c.data_ptr(), ws.mem_c.data_ptr()
            c.w_in, c.b_in = self._w(pre + "multihead_attn.in_proj_weight"), self._p(pre + "multihead_attn.in_proj_bias")
            c.qkv, c.kv, c.kv_ready, c.o = e.q.data_ptr(), e.kv.data_ptr(), int(kv_ready), e.ao2.data_ptr()
            if kv_ready and l == 0:
                plan.add_join()                  # encoder memory + K/V projections (lane 1) are needed from here on
            c.drop_p, c.rng_state, c.site = p_drop, self.rng_state.data_ptr(), _dec_site(l, 2)
            c.gemm_impl = self.gemm_impl
            if self.split_terms:
                c.split_ws, c.split_ws_bytes = self._split_ws(ws, 1 if plan.lane else 0)
            plan.keep.append(c)
            plan.add(f"vct_attn_dec_cross_fwd:{l}", lib.vct_attn_dec_cross_fwd, C.byref(c))
            self._gemm(plan, f"dec{l}.cross.out_proj", Rd, d, d, e.ao2.data_ptr(), d, 0,
                       self._w(pre + "multihead_attn.out_proj.weight"), d, 0, e.s2.data_ptr(), F32, d,
                       bias=self._p(pre + "multihead_attn.out_proj.bias"))
            self._ln_fwd(plan, f"dec{l}.norm2", e.x1.data_ptr(), e.s2.data_ptr(), pre + "norm2.weight", pre + "norm2.bias",
                         e.x2.data_ptr(), e.x2_c.data_ptr() if cd == BF16 else None, e.s2.data_ptr(),
                         e.stats[2].data_ptr(), e.stats[3].data_ptr(), Rd, p_drop, _dec_site(l, 3))
            self._gemm(plan, f"dec{l}.linear1", Rd, D.F_dec, d, e.x2_c.data_ptr(), d, 0, self._w(pre + "linear1.weight"), d, 0,
                       e.z.data_ptr(), cd, D.F_dec, bias=self._p(pre + "linear1.bias"), C2=e.h.data_ptr(), c2_dtype=cd,
                       ldc2=D.F_dec, act=L.ACT_GELU_FWD_F if ws.training else L.ACT_GELU_FWD, drop_p=p_drop, site=_dec_site(l, 4))
            self._gemm(plan, f"dec{l}.linear2", Rd, d, D.F_dec, e.h.data_ptr(), D.F_dec, 0, self._w(pre + "linear2.weight"),
                       D.F_dec, 0, e.s3.data_ptr(), F32, d, bias=self._p(pre + "linear2.bias"))
            self._ln_fwd(plan, f"dec{l}.norm3", e.x2.data_ptr(), e.s3.data_ptr(), pre + "norm3.weight", pre + "norm3.bias",
                         e.x3.data_ptr(), e.x3_c.data_ptr() if cd == BF16 else None, e.s3.data_ptr(),
                         e.stats[4].data_ptr(), e.stats[5].data_ptr(), Rd, p_drop, _dec_site(l, 5))
            x, x_c = e.x3, e.x3_c
        ws.dec_out = x
        self._ln_fwd(plan, "dec.norm", None, x.data_ptr(), "cap_decoder.decoder.norm.weight", "cap_decoder.decoder.norm.bias",
                     ws.hfin.data_ptr(), ws.hfin_c.data_ptr() if cd == BF16 else None, None, ws.hfin_stats[0].data_ptr(),
                     ws.hfin_stats[1].data_ptr(), Rd, 0.0, 0)
        self._gemm(plan, "generator", Rd, D.V, d, ws.hfin_c.data_ptr(), d, 0, self._w("cap_decoder.generator.weight"), d, 0,
                   ws.logits.data_ptr(), ws.logits_dt, ws.Vp, bias=self._p("cap_decoder.generator.bias"))
        if with_loss or with_grad:
            self._sce(plan, ws, with_loss, with_grad)

    def _sce(self, plan: Plan, ws, with_loss: bool, with_grad: bool):
        D = self.dims
        plan.add("vct_sce_typed", self.lib.vct_sce_typed, ws.logits.data_ptr(), ws.logits_dt, ws.Vp, ws.ids.data_ptr(), ws.S + 1, ws.B, ws.S, D.V,
                 float(D.alpha), float(1.0 - D.alpha), D.pad_id,
                 ws.loss.data_ptr() if with_loss else None, ws.row_parts.data_ptr(), self.counters.data_ptr() + 4 * 4,
                 ws.dlogits.data_ptr() if with_grad else None, self.cdt, ws.Vp, self.upstream.data_ptr())

    def plan_forward(self, ws, *, fused_grad: bool, part: str = "all", with_loss: bool = True) -> Plan:
        """part 'all': encoder + decoder + generator + SCE loss; 'dec': decoder side only (memory already
        in ws.mem / ws.mem_c).  fused_grad: the SCE kernel also writes d loss / d logits in the same pass
        (native trainer); otherwise vct_sce runs again in backward with the upstream gradient (autograd)."""
        key = ("fwd", fused_grad, part, with_loss)
        if key not in ws.plans:
            p = Plan()
            p.ws = ws
            pd = float(self.dims.dropout) if ws.training else 0.0
            if part == "all":
                # the encoder and the cross-attention K/V projections of every decoder layer run on lane 1 while the
                # main lane embeds the tokens and runs the first decoder self-attention block; they join before the
                # first cross-attention
                with self._side(p, 1):
                    self._build_encoder(p, ws, pd)
                    self._build_cross_kv(p, ws)
            self._build_decoder(p, ws, pd, with_loss=with_loss, with_grad=fused_grad, kv_ready=(part == "all"))
            ws.plans[key] = p
        return ws.plans[key]

    def plan_encode(self, ws, zero_pad: bool = False) -> Plan:
        """zero_pad: eval-mode fast-path semantics for padded frames (see _build_encoder); never used in training."""
        key = ("encode", bool(zero_pad))
        if key not in ws.plans:
            if zero_pad and ws.training:
                raise RuntimeError("zero_pad is the eval()/no_grad behaviour of the reference's encoder")
            p = Plan()
            p.ws = ws
            self._build_encoder(p, ws, float(self.dims.dropout) if ws.training else 0.0, zero_pad=zero_pad)
            ws.plans[key] = p
        return ws.plans[key]

    # ------------------------------------------------------------------------------------------
    # backward plan
    # ------------------------------------------------------------------------------------------
    def plan_backward(self, ws, *, sce_first: bool, part: str = "all", fuse_adam: bool = False, allreduce=None) -> Plan:
        """part: 'all' (loss -> every gradient), 'dec' (loss -> decoder grads + d memory in ws.g_mem),
        'enc' (ws.g_mem -> encoder grads).  fuse_adam (single-GPU native trainer): every arena slice is updated
        by vct_adam on lane 2 as soon as its gradient is final, overlapping the optimizer's HBM traffic with the
        latency-bound remainder of backward."""
        key = ("bwd", sce_first, part, fuse_adam, allreduce is not None)
        if key in ws.plans:
            return ws.plans[key]
        if not ws.training:
            raise RuntimeError("backward needs a training workspace")
        p = Plan()
        p.ws = ws
        p.fuse_adam = fuse_adam
        p.allreduce = allreduce           # (process group, world size) or None
        if part in ("all", "dec"):
            self._build_decoder_bwd(p, ws, sce_first)
        if part in ("all", "enc"):
            self._build_encoder_bwd(p, ws)
        if fuse_adam and p.adam_covered != self.arena.numel:
            raise RuntimeError(f"optimizer slices cover {p.adam_covered} of {self.arena.numel} arena elements")
        ws.plans[key] = p
        return p

    def _build_decoder_bwd(self, p: Plan, ws, sce_first: bool):
        D, lib = self.dims, self.lib
        d, B, S, M = D.d, ws.B, ws.S, ws.M
        Re, Rd = B * M, B * S
        cd, cdt, es = self.cdt, _TDT[self.cdt], _ESIZE[self.cdt]
        pd = float(D.dropout)
        side = self._side
        if sce_first:
            self._sce(p, ws, with_loss=False, with_grad=True)
        dl = ws.dlogits.data_ptr()
        # ---- generator -------------------------------------------------------------------------
        with side(p):
            self._gemm(p, "generator.wgrad", D.V, d, Rd, dl, ws.Vp, 1, ws.hfin_c.data_ptr(), d, 1,
                       self._g("cap_decoder.generator.weight"), F32, d)
            self._colsum(p, "generator.bias", dl, ws.Vp, Rd, D.V, self._g("cap_decoder.generator.bias"), ws)
        self._gemm(p, "generator.dgrad", Rd, d, D.V, dl, ws.Vp, 0, self._w("cap_decoder.generator.weight"), d, 1,
                   ws.g_a.data_ptr(), F32, d)
        self._ln_bwd(p, "dec.norm", ws.g_a.data_ptr(), ws.dec_out.data_ptr(), ws.hfin_stats[0].data_ptr(),
                     ws.hfin_stats[1].data_ptr(), "cap_decoder.decoder.norm.weight", "cap_decoder.decoder.norm.bias",
                     ws.g_b.data_ptr(), None, None, Rd, 0.0, 0, ws)
        self._adam_slice(p, "cap_decoder.decoder.norm.weight", "cap_decoder.generator.bias")
        dx, other = ws.g_b, ws.g_a          # dx: gradient wrt the current layer's output
        first_mem = True
        for l in reversed(range(D.L_dec)):
            e = ws.dec[l]
            pre = f"cap_decoder.decoder.layers.{l}."
            xin_c = (ws.dec[l - 1].x3_c if l > 0 else ws.e0_c)
            g_r3 = self._scratch(ws, f"dec{l}.dr3", Rd, d, cdt).data_ptr()
            g_z = self._scratch(ws, f"dec{l}.dz", Rd, D.F_dec, cdt).data_ptr()
            g_r2 = self._scratch(ws, f"dec{l}.dr2", Rd, d, cdt).data_ptr()
            g_q = self._scratch(ws, f"dec{l}.dq", Rd, d, cdt).data_ptr()
            g_kv = self._scratch(ws, f"dec{l}.dkv", Re, 2 * d, cdt).data_ptr()
            g_r1 = self._scratch(ws, f"dec{l}.dr1", Rd, d, cdt).data_ptr()
            gq = self._scratch(ws, f"dec{l}.dqkv", Rd, 3 * d, cdt).data_ptr()
            g_o = ws.g_o_c.data_ptr()        # consumed on the main lane only
            # norm3 / FFN
            self._ln_bwd(p, f"dec{l}.norm3", dx.data_ptr(), e.s3.data_ptr(), e.stats[4].data_ptr(), e.stats[5].data_ptr(),
                         pre + "norm3.weight", pre + "norm3.bias", ws.g_s.data_ptr(), g_r3,
                         self._g(pre + "linear2.bias"), Rd, pd, _dec_site(l, 5), ws)
            with side(p):
                self._gemm(p, f"dec{l}.linear2.wgrad", d, D.F_dec, Rd, g_r3, d, 1, e.h.data_ptr(), D.F_dec, 1,
                           self._g(pre + "linear2.weight"), F32, D.F_dec, defer=True)
            self._gemm(p, f"dec{l}.linear2.dgrad", Rd, D.F_dec, d, g_r3, d, 0, self._w(pre + "linear2.weight"),
                       D.F_dec, 1, g_z, cd, D.F_dec, act=L.ACT_MUL_AUX, aux=e.z.data_ptr(), ld_aux=D.F_dec,
                       drop_p=pd, site=_dec_site(l, 4))
            with side(p):
                self._gemm(p, f"dec{l}.linear1.wgrad", D.F_dec, d, Rd, g_z, D.F_dec, 1, e.x2_c.data_ptr(), d, 1,
                           self._g(pre + "linear1.weight"), F32, d, defer=True)
                self._colsum(p, f"dec{l}.linear1.bias", g_z, D.F_dec, Rd, D.F_dec, self._g(pre + "linear1.bias"), ws)
            self._gemm(p, f"dec{l}.linear1.dgrad", Rd, d, D.F_dec, g_z, D.F_dec, 0, self._w(pre + "linear1.weight"),
                       d, 1, other.data_ptr(), F32, d, addend=ws.g_s.data_ptr(), ld_addend=d)
            dx, other = other, dx           # dx = grad wrt x2
            # norm2 / cross attention
            self._ln_bwd(p, f"dec{l}.norm2", dx.data_ptr(), e.s2.data_ptr(), e.stats[2].data_ptr(), e.stats[3].data_ptr(),
                         pre + "norm2.weight", pre + "norm2.bias", ws.g_s.data_ptr(), g_r2,
                         self._g(pre + "multihead_attn.out_proj.bias"), Rd, pd, _dec_site(l, 3), ws)
            with side(p):
                self._gemm(p, f"dec{l}.cross.out_proj.wgrad", d, d, Rd, g_r2, d, 1, e.ao2.data_ptr(), d, 1,
                           self._g(pre + "multihead_attn.out_proj.weight"), F32, d, defer=True)
            self._gemm(p, f"dec{l}.cross.out_proj.dgrad", Rd, d, d, g_r2, d, 0,
                       self._w(pre + "multihead_attn.out_proj.weight"), d, 1, g_o, cd, d)
            self._attn(p, f"dec{l}.cross", True, B=B, H=D.H_dec, Lq=S, Lk=M, q=e.q.data_ptr(), q_ld=d,
                       k=e.kv.data_ptr(), k_ld=2 * d, v=e.kv.data_ptr() + d * es, v_ld=2 * d, o=None, o_ld=d,
                       p=pd, site=_dec_site(l, 2), d_o=g_o, do_ld=d, dq=g_q, dq_ld=d,
                       dk=g_kv, dk_ld=2 * d, dv=g_kv + d * es, dv_ld=2 * d,
                       dbias=self._g(pre + "multihead_attn.in_proj_bias"), ws=ws)
            wname, bname = pre + "multihead_attn.in_proj_weight", pre + "multihead_attn.in_proj_bias"
            with side(p):
                self._gemm(p, f"dec{l}.cross.q.wgrad", d, d, Rd, g_q, d, 1, e.x1_c.data_ptr(), d, 1, self._g(wname), F32, d, defer=True)
                self._gemm(p, f"dec{l}.cross.kv.wgrad", 2 * d, d, Re, g_kv, 2 * d, 1, ws.mem_c.data_ptr(), d, 1,
                           self._g(wname, d * d), F32, d, defer=True)
            self._gemm(p, f"dec{l}.cross.q.dgrad", Rd, d, d, g_q, d, 0, self._w(wname), d, 1,
                       other.data_ptr(), F32, d, addend=ws.g_s.data_ptr(), ld_addend=d)
            self._gemm(p, f"dec{l}.cross.kv.dgrad", Re, d, 2 * d, g_kv, 2 * d, 0, self._w(wname, d), d, 1,
                       ws.g_mem.data_ptr(), F32, d, addend=None if first_mem else ws.g_mem.data_ptr(), ld_addend=d)
            first_mem = False
            dx, other = other, dx           # dx = grad wrt x1
            # norm1 / self attention
            self._ln_bwd(p, f"dec{l}.norm1", dx.data_ptr(), e.s1.data_ptr(), e.stats[0].data_ptr(), e.stats[1].data_ptr(),
                         pre + "norm1.weight", pre + "norm1.bias", ws.g_s.data_ptr(), g_r1,
                         self._g(pre + "self_attn.out_proj.bias"), Rd, pd, _dec_site(l, 1), ws)
            with side(p):
                self._gemm(p, f"dec{l}.self.out_proj.wgrad", d, d, Rd, g_r1, d, 1, e.ao.data_ptr(), d, 1,
                           self._g(pre + "self_attn.out_proj.weight"), F32, d, defer=True)
            self._gemm(p, f"dec{l}.self.out_proj.dgrad", Rd, d, d, g_r1, d, 0,
                       self._w(pre + "self_attn.out_proj.weight"), d, 1, g_o, cd, d)
            qkv = e.qkv.data_ptr()
            self._attn(p, f"dec{l}.self", True, B=B, H=D.H_dec, Lq=S, Lk=S, q=qkv, q_ld=3 * d, k=qkv + d * es, k_ld=3 * d,
                       v=qkv + 2 * d * es, v_ld=3 * d, o=None, o_ld=d, key_pad=ws.tok_pad.data_ptr(), causal=1, p=pd,
                       site=_dec_site(l, 0), d_o=g_o, do_ld=d, dq=gq, dq_ld=3 * d, dk=gq + d * es,
                       dk_ld=3 * d, dv=gq + 2 * d * es, dv_ld=3 * d,
                       dbias=self._g(pre + "self_attn.in_proj_bias"), ws=ws)
            wname, bname = pre + "self_attn.in_proj_weight", pre + "self_attn.in_proj_bias"
            with side(p):
                self._gemm(p, f"dec{l}.self.in_proj.wgrad", 3 * d, d, Rd, gq, 3 * d, 1, xin_c.data_ptr(), d, 1,
                           self._g(wname), F32, d, defer=True)
            self._gemm(p, f"dec{l}.self.in_proj.dgrad", Rd, d, 3 * d, gq, 3 * d, 0, self._w(wname), d, 1, other.data_ptr(), F32, d,
                       addend=ws.g_s.data_ptr(), ld_addend=d)
            dx, other = other, dx           # dx = grad wrt the layer input
            self._flush_group(p, f"dec{l}.wgrads")
            self._adam_slice(p, pre + "self_attn.in_proj_weight", pre + "norm3.bias")
        # ---- embedding ---------------------------------------------------------------------------
        emb, e_lo, e_hi = self._emb_range()
        fused = getattr(p, "fuse_adam", False)
        world = p.allreduce[1] if getattr(p, "allreduce", None) is not None else 1
        mode = self.emb_mode(ws, world) if fused else "plain"
        a = self.arena
        if fused:
            a.ensure_optimizer_state()
        shadow = a.ensure_shadow().data_ptr() + 2 * e_lo if (fused and cd == BF16) else None

        def touched_adam(grad_scale):
            # end-of-step half of the embedding update: only the rows that received gradient (see plan_embed_early)
            p.add("vct_adam_rows:touched", lib.vct_adam_rows, a.p32.data_ptr() + 4 * e_lo, self._g(emb), a.exp_avg.data_ptr() + 4 * e_lo,
                  a.exp_avg_sq.data_ptr() + 4 * e_lo, shadow, D.V, d, self.hyper.data_ptr(), grad_scale, self.emb_stamp.data_ptr(),
                  self.rng_state.data_ptr(), 1)
            p.adam_covered = getattr(p, "adam_covered", 0) + (e_hi - e_lo)

        if mode in ("plain", "local-unsplit", "dense"):
            p.add("vct_embed_bwd", lib.vct_embed_bwd, ws.ids.data_ptr(), S + 1, dx.data_ptr(), self._g(emb), B, S, d, D.V,
                  D.pad_id, pd, self.rng_state.data_ptr(), SITE_EMBED)
            self._adam_slice(p, emb, emb)
            if fused:
                # the table gradient is all-zero at the start of a native-trainer step: once Adam has consumed it, only the
                # rows this step touched are cleared again (no 94 MB memset per step)
                with self._side(p, 2):
                    p.add("vct_embed_zero", lib.vct_embed_zero, ws.ids.data_ptr(), S + 1, self._g(emb), B, S, d, D.V)
        elif mode == "local":
            p.add("vct_embed_bwd", lib.vct_embed_bwd, ws.ids.data_ptr(), S + 1, dx.data_ptr(), self._g(emb), B, S, d, D.V,
                  D.pad_id, pd, self.rng_state.data_ptr(), SITE_EMBED)
            with self._side(p, 2):
                touched_adam(1.0)
                p.add("vct_embed_zero", lib.vct_embed_zero, ws.ids.data_ptr(), S + 1, self._g(emb), B, S, d, D.V)
        else:
            # data parallel: the table gradient has at most B*S non-zero rows per rank.  Exchange THOSE (all-gather of the
            # masked rows; the ids were gathered and sorted at the start of the step) and sum every rank's rows locally in a
            # fixed order, instead of all-reducing the dense 94 MB table gradient as DDP does (train.py:218); 1/world is
            # folded into Adam like for the dense buckets.
            group = p.allreduce[0]
            all_ids, all_rows, rows, keys = self.emb_buffers(ws, world)
            p.add("vct_embed_bwd_rows", lib.vct_embed_bwd_rows, dx.data_ptr(), rows.data_ptr(), B, S, d, pd,
                  self.rng_state.data_ptr(), SITE_EMBED)
            # (peer-memory exchange: own lane and own channel, so that this chain and the per-slice all-reduce + Adam chain
            # of lane 2 do not queue behind each other at the end of the step; NCCL: everything on the optimizer lane)
            with self._side(p, 3 if self.peer is not None else 2):
                def gather(stream, rows=rows, all_rows=all_rows, group=group):
                    import torch.distributed as dist
                    with torch.cuda.stream(stream):
                        dist.all_gather_into_tensor(all_rows, rows, group=group)
                    return 0
                if self.peer is not None:
                    p.add("vct_peer_allgather:embedding_rows", lib.vct_peer_allgather, self.peer.handle, self._peer_rows_off,
                          Rd * d * 4, 1)
                else:
                    p.add("py:all_gather:embedding_rows", gather)
                # every rank sums the same rows in the same order: bit-identical table gradients, replicas cannot drift
                p.add("vct_embed_segment_sum", lib.vct_embed_segment_sum, keys.data_ptr(), all_rows.data_ptr(), self._g(emb),
                      world * Rd, d)
                touched_adam(1.0 / world)
                p.add("vct_embed_zero", lib.vct_embed_zero, all_ids.data_ptr(), S + 1, self._g(emb), world * B, S, d, D.V)

    def _build_encoder_bwd(self, p: Plan, ws):
        D, lib = self.dims, self.lib
        d, B, M = D.d, ws.B, ws.M
        Re = B * M
        cd, cdt, es = self.cdt, _TDT[self.cdt], _ESIZE[self.cdt]
        pd = float(D.dropout)
        side = self._side
        self._ln_bwd(p, "enc.norm", ws.g_mem.data_ptr(), ws.enc_out.data_ptr(), ws.mem_stats[0].data_ptr(),
                     ws.mem_stats[1].data_ptr(), "video_encoder.transformer_encoder.norm.weight",
                     "video_encoder.transformer_encoder.norm.bias", ws.g_a.data_ptr(), None, None, Re, 0.0, 0, ws)
        # Data parallel: the final norm travels with the last layer's slice and unify with the first layer's (contiguous in
        # the arena), because every extra exchange costs two cross-GPU barriers at the very end of the step.
        merge = getattr(p, "allreduce", None) is not None
        if not merge:
            self._adam_slice(p, "video_encoder.transformer_encoder.norm.weight", "video_encoder.transformer_encoder.norm.bias")
        dx, other = ws.g_a, ws.g_b
        g_o = ws.g_o_c.data_ptr()
        for l in reversed(range(D.L_enc)):
            e = ws.enc[l]
            pre = f"video_encoder.transformer_encoder.layers.{l}."
            xin_c = (ws.enc[l - 1].x2_c if l > 0 else ws.x0_c)
            g_r2 = self._scratch(ws, f"enc{l}.dr2", Re, d, cdt).data_ptr()
            g_z = self._scratch(ws, f"enc{l}.dz", Re, D.F_enc, cdt).data_ptr()
            g_r1 = self._scratch(ws, f"enc{l}.dr1", Re, d, cdt).data_ptr()
            gq = self._scratch(ws, f"enc{l}.dqkv", Re, 3 * d, cdt).data_ptr()
            self._ln_bwd(p, f"enc{l}.norm2", dx.data_ptr(), e.s2.data_ptr(), e.stats[2].data_ptr(), e.stats[3].data_ptr(),
                         pre + "norm2.weight", pre + "norm2.bias", ws.g_s.data_ptr(), g_r2,
                         self._g(pre + "linear2.bias"), Re, pd, _enc_site(l, 3), ws)
            with side(p):
                self._gemm(p, f"enc{l}.linear2.wgrad", d, D.F_enc, Re, g_r2, d, 1, e.h.data_ptr(), D.F_enc, 1,
                           self._g(pre + "linear2.weight"), F32, D.F_enc, defer=True)
            self._gemm(p, f"enc{l}.linear2.dgrad", Re, D.F_enc, d, g_r2, d, 0, self._w(pre + "linear2.weight"),
                       D.F_enc, 1, g_z, cd, D.F_enc, act=L.ACT_MUL_AUX, aux=e.z.data_ptr(), ld_aux=D.F_enc,
                       drop_p=pd, site=_enc_site(l, 2))
            with side(p):
                self._gemm(p, f"enc{l}.linear1.wgrad", D.F_enc, d, Re, g_z, D.F_enc, 1, e.x1_c.data_ptr(), d, 1,
                           self._g(pre + "linear1.weight"), F32, d, defer=True)
                self._colsum(p, f"enc{l}.linear1.bias", g_z, D.F_enc, Re, D.F_enc, self._g(pre + "linear1.bias"), ws)
            self._gemm(p, f"enc{l}.linear1.dgrad", Re, d, D.F_enc, g_z, D.F_enc, 0, self._w(pre + "linear1.weight"),
                       d, 1, other.data_ptr(), F32, d, addend=ws.g_s.data_ptr(), ld_addend=d)
            dx, other = other, dx
            self._ln_bwd(p, f"enc{l}.norm1", dx.data_ptr(), e.s1.data_ptr(), e.stats[0].data_ptr(), e.stats[1].data_ptr(),
                         pre + "norm1.weight", pre + "norm1.bias", ws.g_s.data_ptr(), g_r1,
                         self._g(pre + "self_attn.out_proj.bias"), Re, pd, _enc_site(l, 1), ws)
            with side(p):
                self._gemm(p, f"enc{l}.out_proj.wgrad", d, d, Re, g_r1, d, 1, e.ao.data_ptr(), d, 1,
                           self._g(pre + "self_attn.out_proj.weight"), F32, d, defer=True)
            self._gemm(p, f"enc{l}.out_proj.dgrad", Re, d, d, g_r1, d, 0, self._w(pre + "self_attn.out_proj.weight"),
                       d, 1, g_o, cd, d)
            qkv = e.qkv.data_ptr()
            self._attn(p, f"enc{l}.self", True, B=B, H=D.H_enc, Lq=M, Lk=M, q=qkv, q_ld=3 * d, k=qkv + d * es, k_ld=3 * d,
                       v=qkv + 2 * d * es, v_ld=3 * d, o=None, o_ld=d, key_pad=ws.vid_pad.data_ptr(), causal=0, p=pd,
                       site=_enc_site(l, 0), d_o=g_o, do_ld=d, dq=gq, dq_ld=3 * d, dk=gq + d * es,
                       dk_ld=3 * d, dv=gq + 2 * d * es, dv_ld=3 * d,
                       dbias=self._g(pre + "self_attn.in_proj_bias"), ws=ws)
            wname, bname = pre + "self_attn.in_proj_weight", pre + "self_attn.in_proj_bias"
            with side(p):
                self._gemm(p, f"enc{l}.in_proj.wgrad", 3 * d, d, Re, gq, 3 * d, 1, xin_c.data_ptr(), d, 1, self._g(wname), F32, d, defer=True)
            last = l == 0
            self._gemm(p, f"enc{l}.in_proj.dgrad", Re, d, 3 * d, gq, 3 * d, 0, self._w(wname), d, 1, other.data_ptr(), F32, d,
                       addend=ws.g_s.data_ptr(), ld_addend=d,
                       C2=ws.g_x0_c.data_ptr() if (last and cd == BF16) else None, c2_dtype=cd, ldc2=d)
            dx, other = other, dx
            if l > 0:                                        # (layer 0's weight gradients are grouped with unify's, below)
                self._flush_group(p, f"enc{l}.wgrads")
                if not merge:
                    self._adam_slice(p, pre + "self_attn.in_proj_weight", pre + "norm2.bias")
                else:
                    self._adam_slice(p, pre + "self_attn.in_proj_weight",
                                     "video_encoder.transformer_encoder.norm.bias" if l == D.L_enc - 1 else pre + "norm2.bias")
        g_x0_c = ws.g_x0_c.data_ptr() if cd == BF16 else dx.data_ptr()
        with side(p):
            self._gemm(p, "unify.wgrad", d, D.Din, Re, g_x0_c, d, 1, ws.a0.data_ptr(), D.Din, 1,
                       self._g("video_encoder.unify.0.weight"), F32, D.Din, defer=True)
            self._colsum(p, "unify.bias", g_x0_c, d, Re, d, self._g("video_encoder.unify.0.bias"), ws)
        self._flush_group(p, "enc0.wgrads+unify")
        if not merge:
            if D.L_enc > 0:
                pre0 = "video_encoder.transformer_encoder.layers.0."
                self._adam_slice(p, pre0 + "self_attn.in_proj_weight", pre0 + "norm2.bias")
            self._adam_slice(p, "video_encoder.unify.0.weight", "video_encoder.unify.0.bias")
        elif D.L_enc == 0:
            self._adam_slice(p, "video_encoder.unify.0.weight", "video_encoder.transformer_encoder.norm.bias")
        else:
            pre0 = "video_encoder.transformer_encoder.layers.0."
            self._adam_slice(p, "video_encoder.unify.0.weight",
                             "video_encoder.transformer_encoder.norm.bias" if D.L_enc == 1 else pre0 + "norm2.bias")

    # ------------------------------------------------------------------------------------------
    # running
    # ------------------------------------------------------------------------------------------
    def stage_inputs(self, ws, feats: torch.Tensor, vid_pad: Optional[torch.Tensor], ids: Optional[torch.Tensor],
                     tok_pad: Optional[torch.Tensor] = None):
        """Copy the step's inputs into the workspace (host or device sources; async on the current stream).
        tok_pad: optional bool [B, S] key-padding mask of the decoder inputs (default: ids[:, :-1] == pad_id, which is
        what the reference's CapPreprocessor produces, model/CapPreprocessor.py:35)."""
        if self._stage_native(ws, feats, vid_pad, ids, tok_pad):
            return
        ws.feats.copy_(feats.reshape(ws.feats.shape), non_blocking=True)
        ws.vid_pad[:, 0] = 0
        if vid_pad is None:
            ws.vid_pad[:, 1:] = 0
        else:
            ws.vid_pad[:, 1:].copy_(vid_pad, non_blocking=True)
        if ids is not None:
            self.stage_ids(ws, ids, tok_pad)

    def _stage_native(self, ws, feats, vid_pad, ids, tok_pad) -> bool:
        """Device-resident sources in the layouts the loaders / ``CaptionTrainer.prefetch`` produce: ONE vct_stage_inputs
        launch instead of five torch copy / fill / compare launches.  Anything else (host tensors, other dtypes, strided
        views) keeps the torch copies below, which convert as they go.

        OFF by default (VCT_STAGE_NATIVE=1 turns it on).  Measured A/B on one B200 box (profiles/r02f_stage_inputs_ab.txt):
        the device-resident loop gains 1 % (1.532 vs 1.546 ms/step), but the end-to-end loop -- where the host is on the
        critical path after every ``loss.item()`` -- dropped from 39.4-39.5k to 35.3-38.3k captions/s and became noisy,
        with and without programmatic dependent launch on the kernel; the cause was not found within the round's GPU
        budget, and the end-to-end number is the one users see."""
        def dev_ok(t, dtypes, shape):
            return (t.is_cuda and t.device == ws.feats.device and t.dtype in dtypes and t.is_contiguous()
                    and tuple(t.shape) == tuple(shape))
        if os.environ.get("VCT_STAGE_NATIVE", "0") != "1":
            return False
        B, T, S = ws.B, ws.T, ws.S
        byte = (torch.bool, torch.uint8)
        if feats is not None and not (dev_ok(feats, (torch.float32,), (B, T, self.dims.Din)) and feats.data_ptr() % 16 == 0
                                      and self.dims.Din % 4 == 0):
            return False
        if vid_pad is not None and not dev_ok(vid_pad, byte, (B, T)):
            return False
        if ids is not None and not dev_ok(ids, (torch.int64,), (B, S + 1)):
            return False
        if tok_pad is not None and (ids is None or not dev_ok(tok_pad, byte, (B, S))):
            return False
        if feats is None and ids is None:
            return False
        L.check(self.lib.vct_stage_inputs(
            feats.data_ptr() if feats is not None else None, ws.feats.data_ptr(),
            vid_pad.data_ptr() if (vid_pad is not None and feats is not None) else None, ws.vid_pad.data_ptr(),
            B, T, self.dims.Din, ids.data_ptr() if ids is not None else None, ws.ids.data_ptr(),
            tok_pad.data_ptr() if tok_pad is not None else None, ws.tok_pad.data_ptr(), S + 1, int(self.dims.pad_id),
            self._stream()), "vct_stage_inputs")
        return True

    def stage_ids(self, ws, ids: torch.Tensor, tok_pad: Optional[torch.Tensor] = None) -> None:
        if self._stage_native(ws, None, None, ids, tok_pad):
            return
        ws.ids.copy_(ids, non_blocking=True)
        if tok_pad is None:
            torch.eq(ws.ids[:, :-1], self.dims.pad_id, out=ws.tok_pad.view(torch.bool))
        else:
            ws.tok_pad.view(torch.bool).copy_(tok_pad.reshape(ws.tok_pad.shape), non_blocking=True)

    def run(self, plan: Plan) -> None:
        self.launches += plan.run(torch.cuda.current_stream(self.device), self.side_streams)

    def zero_scatter_grads(self) -> None:
        """Only the embedding gradient is accumulated with atomics; every other gradient is overwritten."""
        self.arena.grad_view("cap_decoder.tgt_to_emb.weight").zero_()

    # ---- Adam ------------------------------------------------------------------------------------
    def set_adam(self, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0) -> None:
        self.arena.ensure_optimizer_state()
        h = self.hyper.tolist()
        h[0:5] = [lr, betas[0], betas[1], eps, weight_decay]
        self.hyper.copy_(torch.tensor(h, dtype=torch.float32))

    def set_lr(self, lr: float) -> None:
        self.hyper[0:1].fill_(lr)

    def tick(self) -> None:
        L.check(self.lib.vct_step_tick(self.rng_state.data_ptr(), self.hyper.data_ptr(), self._stream()), "vct_step_tick")
        self.launches += 1
        self._dev_step = None        # advanced on the device (possibly inside a CUDA graph): host mirror unknown

    def adam(self, grad_scale: float = 1.0) -> None:
        a = self.arena
        shadow = a.ensure_shadow().data_ptr() if self.cdt == BF16 else None
        L.check(self.lib.vct_adam(a.p32.data_ptr(), a.grad.data_ptr(), F32, a.exp_avg.data_ptr(), a.exp_avg_sq.data_ptr(), shadow,
                                  a.numel, self.hyper.data_ptr(), grad_scale, self._stream()), "vct_adam")
        self.launches += 1

    def embed_zero(self, ws) -> None:
        """Clear the embedding-gradient rows the step of ``ws`` scattered into (non-fused optimizer path)."""
        D = self.dims
        L.check(self.lib.vct_embed_zero(ws.ids.data_ptr(), ws.S + 1, self._g("cap_decoder.tgt_to_emb.weight"), ws.B, ws.S, D.d, D.V,
                                        self._stream()), "vct_embed_zero")
        self.launches += 1

    # ------------------------------------------------------------------------------------------
    # greedy decoding with a K/V cache (model/MMT4Caption.py:146-172, model/CapDecoder.py:62-79)
    # ------------------------------------------------------------------------------------------
    def decode_workspace(self, B: int, T: int, max_len: int) -> SimpleNamespace:
        key = ("decode", B, T, max_len)
        if key in self._ws:
            return self._ws[key]
        self._check_lengths(T + 1, max_len - 1)
        D = self.dims
        d, M = D.d, T + 1
        cdt, f32 = _TDT[self.cdt], torch.float32
        two = self.cdt == BF16
        dev = self.device
        ws = SimpleNamespace(B=B, T=T, M=M, max_len=max_len, Vp=(D.V + 7) // 8 * 8)
        ws.ys = torch.zeros((B, max_len), dtype=torch.int64, device=dev)
        ws.ended = torch.zeros(B, dtype=torch.int32, device=dev)
        ws.n_ended = torch.zeros(1, dtype=torch.int32, device=dev)

        def pair(rows, cols):
            t = torch.empty((rows, cols), dtype=f32, device=dev)
            return t, (torch.empty((rows, cols), dtype=cdt, device=dev) if two else t)

        ws.x, ws.x_c = pair(B, d)
        ws.layers = []
        for _ in range(D.L_dec):
            e = SimpleNamespace()
            e.cache = torch.zeros((B, max_len, 3 * d), dtype=cdt, device=dev)    # q,k,v of every position so far
            e.ao = torch.empty((B, d), dtype=cdt, device=dev)
            e.s1 = torch.empty((B, d), dtype=f32, device=dev)
            e.x1, e.x1_c = pair(B, d)
            e.q = torch.empty((B, d), dtype=cdt, device=dev)
            e.kv = torch.empty((B * M, 2 * d), dtype=cdt, device=dev)            # cross K/V: projected once
            e.ao2 = torch.empty((B, d), dtype=cdt, device=dev)
            e.s2 = torch.empty((B, d), dtype=f32, device=dev)
            e.x2, e.x2_c = pair(B, d)
            e.z = torch.empty((B, D.F_dec), dtype=cdt, device=dev)
            e.h = torch.empty((B, D.F_dec), dtype=cdt, device=dev)
            e.s3 = torch.empty((B, d), dtype=f32, device=dev)
            e.x3, e.x3_c = pair(B, d)
            e.probs = None
            ws.layers.append(e)
        ws.hfin, ws.hfin_c = pair(B, d)
        ws.logits = torch.empty((B, ws.Vp), dtype=f32, device=dev)
        self._alloc_split_ws(ws, B * M, False)
        ws.plans = {}
        self._ws[key] = ws
        return ws

    def plan_decode_init(self, dws, enc_ws) -> Plan:
        """cross-attention K/V of every decoder layer from the encoder memory (once per batch)."""
        key = ("init",)
        if getattr(dws, "init_ws", None) is not enc_ws:
            dws.plans.pop(key, None)      # the encoder workspace was re-allocated (LRU eviction): its pointers changed
            dws.init_ws = enc_ws          # (the reference also keeps that workspace alive as long as this plan exists)
        if key not in dws.plans:
            D = self.dims
            d = D.d
            p = Plan()
            p.ws = dws
            for l, e in enumerate(dws.layers):
                pre = f"cap_decoder.decoder.layers.{l}.multihead_attn."
                self._gemm(p, f"dec{l}.cross.kv", dws.B * dws.M, 2 * d, d, enc_ws.mem_c.data_ptr(), d, 0,
                           self._w(pre + "in_proj_weight", d), d, 0, e.kv.data_ptr(), self.cdt, 2 * d,
                           bias=self._p(pre + "in_proj_bias", d))
            dws.plans[key] = p
        return dws.plans[key]

    def plan_decode_step(self, dws, t: int, want_probs: bool = False) -> Plan:
        """Feed token column t (ys[:, t]) and append ys[:, t+1] = argmax of the next-word logits."""
        key = ("step", t, want_probs)
        if key in dws.plans:
            return dws.plans[key]
        D, lib = self.dims, self.lib
        d, B, M, Lmax = D.d, dws.B, dws.M, dws.max_len
        cd, es = self.cdt, _ESIZE[self.cdt]
        p = Plan()
        p.ws = dws
        p.add("vct_embed_fwd", lib.vct_embed_fwd, dws.ys.data_ptr() + 8 * t, Lmax, self._p("cap_decoder.tgt_to_emb.weight"),
              self.pos.data_ptr(), dws.x.data_ptr(), dws.x_c.data_ptr() if cd == BF16 else None, cd, B, 1, d, D.V, t,
              0.0, self.rng_state.data_ptr(), SITE_EMBED)
        x, x_c = dws.x, dws.x_c
        for l, e in enumerate(dws.layers):
            pre = f"cap_decoder.decoder.layers.{l}."
            cache = e.cache.data_ptr()
            row = cache + t * 3 * d * es
            self._gemm(p, f"dec{l}.self.in_proj", B, 3 * d, d, x_c.data_ptr(), d, 0, self._w(pre + "self_attn.in_proj_weight"),
                       d, 0, row, cd, Lmax * 3 * d, bias=self._p(pre + "self_attn.in_proj_bias"))
            self._attn(p, f"dec{l}.self", False, B=B, H=D.H_dec, Lq=1, Lk=t + 1, q=row, q_ld=3 * d, q_bs=Lmax * 3 * d,
                       k=cache + d * es, k_ld=3 * d, k_bs=Lmax * 3 * d, v=cache + 2 * d * es, v_ld=3 * d, v_bs=Lmax * 3 * d,
                       o=e.ao.data_ptr(), o_ld=d, o_bs=d)
            self._gemm(p, f"dec{l}.self.out_proj", B, d, d, e.ao.data_ptr(), d, 0, self._w(pre + "self_attn.out_proj.weight"),
                       d, 0, e.s1.data_ptr(), F32, d, bias=self._p(pre + "self_attn.out_proj.bias"))
            self._ln_fwd(p, f"dec{l}.norm1", x.data_ptr(), e.s1.data_ptr(), pre + "norm1.weight", pre + "norm1.bias",
                         e.x1.data_ptr(), e.x1_c.data_ptr() if cd == BF16 else None, None, None, None, B, 0.0, 0)
            self._gemm(p, f"dec{l}.cross.q", B, d, d, e.x1_c.data_ptr(), d, 0, self._w(pre + "multihead_attn.in_proj_weight"),
                       d, 0, e.q.data_ptr(), cd, d, bias=self._p(pre + "multihead_attn.in_proj_bias"))
            if want_probs and e.probs is None:
                e.probs = torch.zeros((B, D.H_dec, 1, M), dtype=torch.float32, device=self.device)
            self._attn(p, f"dec{l}.cross", False, B=B, H=D.H_dec, Lq=1, Lk=M, q=e.q.data_ptr(), q_ld=d, q_bs=d,
                       k=e.kv.data_ptr(), k_ld=2 * d, v=e.kv.data_ptr() + d * es, v_ld=2 * d, o=e.ao2.data_ptr(), o_ld=d,
                       o_bs=d, probs=e.probs.data_ptr() if want_probs else None)
            self._gemm(p, f"dec{l}.cross.out_proj", B, d, d, e.ao2.data_ptr(), d, 0,
                       self._w(pre + "multihead_attn.out_proj.weight"), d, 0, e.s2.data_ptr(), F32, d,
                       bias=self._p(pre + "multihead_attn.out_proj.bias"))
            self._ln_fwd(p, f"dec{l}.norm2", e.x1.data_ptr(), e.s2.data_ptr(), pre + "norm2.weight", pre + "norm2.bias",
                         e.x2.data_ptr(), e.x2_c.data_ptr() if cd == BF16 else None, None, None, None, B, 0.0, 0)
            self._gemm(p, f"dec{l}.linear1", B, D.F_dec, d, e.x2_c.data_ptr(), d, 0, self._w(pre + "linear1.weight"), d, 0,
                       e.z.data_ptr(), cd, D.F_dec, bias=self._p(pre + "linear1.bias"), C2=e.h.data_ptr(), c2_dtype=cd,
                       ldc2=D.F_dec, act=L.ACT_GELU_FWD)
            self._gemm(p, f"dec{l}.linear2", B, d, D.F_dec, e.h.data_ptr(), D.F_dec, 0, self._w(pre + "linear2.weight"),
                       D.F_dec, 0, e.s3.data_ptr(), F32, d, bias=self._p(pre + "linear2.bias"))
            self._ln_fwd(p, f"dec{l}.norm3", e.x2.data_ptr(), e.s3.data_ptr(), pre + "norm3.weight", pre + "norm3.bias",
                         e.x3.data_ptr(), e.x3_c.data_ptr() if cd == BF16 else None, None, None, None, B, 0.0, 0)
            x, x_c = e.x3, e.x3_c
        self._ln_fwd(p, "dec.norm", None, x.data_ptr(), "cap_decoder.decoder.norm.weight", "cap_decoder.decoder.norm.bias",
                     dws.hfin.data_ptr(), dws.hfin_c.data_ptr() if cd == BF16 else None, None, None, None, B, 0.0, 0)
        self._gemm(p, "generator", B, D.V, d, dws.hfin_c.data_ptr(), d, 0, self._w("cap_decoder.generator.weight"), d, 0,
                   dws.logits.data_ptr(), F32, dws.Vp, bias=self._p("cap_decoder.generator.bias"))
        p.add("vct_argmax_append", lib.vct_argmax_append, dws.logits.data_ptr(), dws.Vp, B, D.V, dws.ys.data_ptr(), Lmax,
              t + 1, self.end_id, dws.ended.data_ptr(), dws.n_ended.data_ptr())
        dws.plans[key] = p
        return p

    end_id = 102

    @torch.no_grad()
    def greedy_decode(self, feats: torch.Tensor, vid_pad: Optional[torch.Tensor], max_len: int, start_id: int,
                      end_id: int, sync_every: int = 1, want_probs: bool = False, eval_fastpath: bool = True):
        """Returns ys [B, n] (device int64, incl. the start token).  Same stopping rule as the reference:
        stop once every row has produced end_id (checked every ``sync_every`` tokens; the reference checks
        every token through ``.tolist()``, model/MMT4Caption.py:168-172).  Checking less often only appends
        tokens after a row's first [SEP], which the caption cut discards."""
        self.check_arena()
        self.refresh_shadow()
        self.end_id = end_id
        B, T, _ = feats.shape
        enc_ws = self.workspace(B, T, 1, False)
        self.stage_inputs(enc_ws, feats, vid_pad, None)
        # greedy decoding runs under eval() + no_grad in the reference (eval.py:125-142): with masks the encoder takes
        # the nested-tensor fast path (padded memory rows = norm.bias); predict_video.py passes no masks
        self.run(self.plan_encode(enc_ws, zero_pad=(vid_pad is not None and eval_fastpath)))
        enc_ws.enc_version += 1
        dws = self.decode_workspace(B, T, max_len)
        dws.ys.zero_()
        dws.ys[:, 0] = start_id
        dws.ended.zero_()
        dws.n_ended.zero_()
        self.run(self.plan_decode_init(dws, enc_ws))
        n = 1
        probs = [[] for _ in dws.layers] if want_probs else None
        use_graph = os.environ.get("VCT_DECODE_GRAPH", "1") != "0" and not want_probs
        for t in range(max_len - 1):
            plan = self.plan_decode_step(dws, t, want_probs)
            if use_graph:
                # the step's launch sequence is static: first use runs eagerly, second use is captured, later uses replay
                seen = dws.plans.get(("seen", t), 0)
                g = dws.plans.get(("graph", t))
                if g is not None:
                    g.replay()
                    self.launches += len(plan)
                elif seen >= 1:
                    g = torch.cuda.CUDAGraph()
                    before = self.launches
                    with torch.cuda.graph(g):
                        self.run(plan)
                    self.launches = before + len(plan)
                    dws.plans[("graph", t)] = g
                    g.replay()
                else:
                    self.run(plan)
                    dws.plans[("seen", t)] = seen + 1
            else:
                self.run(plan)
            n = t + 2
            if want_probs:
                for l, e in enumerate(dws.layers):
                    probs[l].append(e.probs.mean(dim=1).clone())      # head average [B, 1, M]
            if (t + 1) % sync_every == 0 or t == max_len - 2:
                if int(dws.n_ended.item()) >= B:
                    break
        ys = dws.ys[:, :n].clone()
        if want_probs:
            return ys, [torch.cat(pl, dim=1) for pl in probs]
        return ys
