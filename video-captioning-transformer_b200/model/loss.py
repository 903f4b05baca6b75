"""Losses (reference: model/loss.py).

``SCELoss`` -- the caption loss on the hot path -- runs on the ``vct_sce`` kernel (one pass over the
logits, closed form of SURVEY Q9).  The contrastive losses belong to the "match"/"cross" tasks, which
are outside the hot path (SURVEY section 2.1 #5b); they are kept as thin torch expressions so that
``Matching`` constructs and its parameters/state_dict keys exist."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor


class SCELoss(nn.Module):
    """alpha * CE(ignore_index) + beta * mean(RCE), model/loss.py:69-92."""

    def __init__(self, alpha, beta, ignore_index, num_classes=10, device=torch.device('cuda')):
        super().__init__()
        self.device = device
        self.alpha = alpha
        self.beta = beta
        self.num_classes = num_classes
        self.ignore_index = ignore_index

    def forward(self, pred: Tensor, labels: Tensor):
        from vct.functional import sce_loss
        return sce_loss(None, pred, labels, self.alpha, self.beta, self.ignore_index)


class ClipSymmetricalLoss(nn.Module):
    """Symmetric video<->text contrastive loss (model/loss.py:7-35); "match"/"cross" tasks only."""

    def __init__(self, enable_tem=False, tem: float = None, device=torch.device("cuda")):
        super().__init__()
        self.device = device
        self.enable_tem = enable_tem
        self.temperature = None
        if tem is not None:
            self.temperature = torch.tensor([tem], dtype=torch.float32, device=device)
        elif enable_tem is True:
            self.temperature = nn.Parameter(torch.tensor([1.0]), requires_grad=True)

    def _sim(self, batch_video: Tensor, batch_text: Tensor) -> Tensor:
        v = batch_video / torch.linalg.norm(batch_video, dim=-1, keepdim=True)
        t = batch_text / torch.linalg.norm(batch_text, dim=-1, keepdim=True)
        sim = v @ t.T
        return sim * torch.exp(self.temperature) if self.temperature is not None else sim

    def forward(self, batch_video: Tensor, batch_text: Tensor):
        sim = self._sim(batch_video, batch_text)
        target = torch.arange(len(batch_video), dtype=torch.long, device=sim.device)
        return (F.cross_entropy(sim, target) + F.cross_entropy(sim.T, target)) / 2


class ClipSymmetricalLoss_WithDualSoftmax(ClipSymmetricalLoss):
    """Dual-softmax variant (model/loss.py:38-66)."""

    def forward(self, batch_video: Tensor, batch_text: Tensor):
        v = batch_video / torch.linalg.norm(batch_video, dim=-1, keepdim=True)
        t = batch_text / torch.linalg.norm(batch_text, dim=-1, keepdim=True)
        sim = v @ t.T
        sim = sim * F.softmax(sim / self.temperature, dim=0) * len(sim)
        target = torch.arange(len(batch_video), dtype=torch.long, device=sim.device)
        return (F.cross_entropy(sim, target) + F.cross_entropy(sim.T, target)) / 2
