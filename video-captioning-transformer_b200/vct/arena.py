"""Flat parameter arena: every parameter of the video encoder and the caption decoder lives in ONE
contiguous fp32 device buffer (HBM layout chosen for B200: one Adam launch, one gradient
all-reduce payload, 16-byte aligned rows for TMA), with parallel arenas for gradients, Adam
moments and the bf16 shadow copy the tensor-core GEMMs read.  ``nn.Parameter`` objects keep their
names / shapes (state_dict keys are the reference's); only their storage moves.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

ALIGN = 64  # elements; keeps every tensor 128-byte (bf16) / 256-byte (fp32) aligned


class ParamArena:
    def __init__(self, named_params: List[Tuple[str, torch.nn.Parameter]], device: torch.device):
        self.device = device
        self.names: List[str] = []
        self.offset: Dict[str, int] = {}
        self.shape: Dict[str, torch.Size] = {}
        self.params: Dict[str, torch.nn.Parameter] = {}
        off = 0
        for name, p in named_params:
            if p.dtype != torch.float32:
                raise TypeError(f"{name}: parameters must be fp32 master weights (got {p.dtype}); SURVEY Q17")
            self.names.append(name)
            self.offset[name] = off
            self.shape[name] = p.shape
            self.params[name] = p
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.p32 = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros(off, dtype=torch.float32, device=device)
        self.shadow = None       # bf16 copy, created on demand
        self.exp_avg = None      # Adam moments, created on demand
        self.exp_avg_sq = None
        self._shadow_version = -1
        with torch.no_grad():
            for name in self.names:
                p = self.params[name]
                view = self.view(self.p32, name)
                view.copy_(p.detach())
                p.data = view

    # ---- views -------------------------------------------------------------------------------
    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        o = self.offset[name]
        shp = self.shape[name]
        return flat[o:o + shp.numel()].view(shp)

    def grad_view(self, name: str) -> torch.Tensor:
        return self.view(self.grad, name)

    def is_current(self, quick: bool = False) -> bool:
        """True while every parameter still aliases its arena slot (``.to()``, ``.half()`` or an
        optimizer that swaps ``p.data`` would break that).  quick: look at the first, middle and last parameter only
        (module-wide moves re-allocate all of them) -- the per-step check of the native trainer, which runs the full
        check every 32nd step."""
        base = self.p32.data_ptr()
        names = self.names if not quick else [self.names[0], self.names[len(self.names) // 2], self.names[-1]]
        for name in names:
            p = self.params[name]
            if p.data_ptr() != base + 4 * self.offset[name] or p.dtype != torch.float32:
                return False
        return True

    def version(self) -> int:
        return sum(self.params[n]._version for n in self.names)

    def ensure_optimizer_state(self) -> None:
        if self.exp_avg is None:
            self.exp_avg = torch.zeros_like(self.p32)
            self.exp_avg_sq = torch.zeros_like(self.p32)

    def ensure_grad16(self) -> torch.Tensor:
        """bf16 twin of the gradient arena: payload of the data-parallel exchange when gradients travel in bf16."""
        if getattr(self, "grad16", None) is None:
            self.grad16 = torch.zeros(self.numel, dtype=torch.bfloat16, device=self.device)
        return self.grad16

    def ensure_shadow(self) -> torch.Tensor:
        if self.shadow is None:
            self.shadow = torch.empty(self.numel, dtype=torch.bfloat16, device=self.device)
        return self.shadow
