"""clip_grad_norm_ + Adam(amsgrad, L2 weight decay) as two kernels over flat buffers.

Replaces ``torch.nn.utils.clip_grad_norm_`` + ``torch.optim.Adam.step`` of the reference
(solver.py:75-77, 91-93).  ``FusedAdam`` subclasses ``torch.optim.Adam`` only to inherit
its ``state_dict`` / ``load_state_dict`` *format* (so ``.opt`` checkpoints written by the
reference load here and vice versa, solver.py:42,54); the update itself is
``avc_sqnorm`` + ``avc_adam_step`` on one flat parameter / gradient / moment buffer.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L

# device hyper-parameter vector layout (avc_b200.h)
HP_LREC, HP_LKL, HP_GSCALE, HP_LR, HP_B1, HP_B2, HP_EPS, HP_WD, HP_MAXNORM, HP_AMSGRAD = range(10)
HP_SIZE = 16


class FusedAdam(torch.optim.Adam):
    def __init__(self, model, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=True,
                 max_norm: float = 5.0, world_size: int = 1):
        flat = model._flat if getattr(model, "_flat", None) is not None else model.flatten_parameters()
        if not flat.is_cuda:
            raise L.AvcError("FusedAdam needs the model on a CUDA device (no CPU path)")
        params = list(model.parameters())
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
        self.lib = L.load()
        self.flat_p = flat
        n = flat.numel()
        dev = flat.device
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_vmax = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.scratch = torch.zeros(1024, dtype=torch.float32, device=dev)
        self.max_norm = float(max_norm)
        self.world_size = int(world_size)
        self.hp_host = torch.zeros(HP_SIZE, dtype=torch.float32)   # pageable on purpose, see sync_hparams
        self.hp = torch.zeros(HP_SIZE, dtype=torch.float32, device=dev)
        self._views()
        self.sync_hparams()

    # ---- views of the flat buffers in the torch.optim.Adam state layout
    def _views(self):
        off = 0
        self.grad_views = {}
        for p in self.param_groups[0]["params"]:
            n = p.numel()
            sl = slice(off, off + n)
            self.state[p] = {
                "step": torch.tensor(0.0),
                "exp_avg": self.flat_m[sl].view(p.shape),
                "exp_avg_sq": self.flat_v[sl].view(p.shape),
                "max_exp_avg_sq": self.flat_vmax[sl].view(p.shape),
            }
            self.grad_views[p] = self.flat_g[sl].view(p.shape)
            off += n

    def named_grad_views(self, model):
        return {name: self.grad_views[p] for name, p in model.named_parameters()}

    def sync_hparams(self, lambda_rec: Optional[float] = None, lambda_kl: Optional[float] = None):
        g = self.param_groups[0]
        h = self.hp_host
        if lambda_rec is not None:
            h[HP_LREC] = lambda_rec
        if lambda_kl is not None:
            h[HP_LKL] = lambda_kl
        h[HP_GSCALE] = 1.0 / self.world_size
        h[HP_LR], h[HP_B1], h[HP_B2] = g["lr"], g["betas"][0], g["betas"][1]
        h[HP_EPS], h[HP_WD], h[HP_MAXNORM] = g["eps"], g["weight_decay"], self.max_norm
        h[HP_AMSGRAD] = 1.0 if g["amsgrad"] else 0.0
        # 64 bytes from PAGEABLE memory: the driver stages the bytes before the call returns, so the host
        # vector may be rewritten for the next step (per-iteration KL annealing) while this copy is still
        # queued behind the previous step -- a pinned source with non_blocking=True would race
        self.hp.copy_(h)

    def zero_grad(self, set_to_none: bool = True):
        st = torch.cuda.current_stream(self.flat_g.device).cuda_stream
        L.check(self.lib.avc_fill_zero(self.flat_g.data_ptr(), self.flat_g.numel() * 4, st), "zero_grad")
        for p in self.param_groups[0]["params"]:
            p.grad = None

    def gather_autograd_grads(self):
        """Copy .grad tensors produced by autograd into the flat gradient buffer (only used
        when the step is driven through ``loss.backward()`` instead of the fused trainer)."""
        for p in self.param_groups[0]["params"]:
            if p.grad is not None and p.grad.data_ptr() != self.grad_views[p].data_ptr():
                self.grad_views[p].copy_(p.grad)

    @torch.no_grad()
    def step(self, closure=None):
        """grad-norm + clip + Adam on the flat buffers.  Returns nothing; the pre-clip norm is
        ``grad_norm()`` (a device scalar until read)."""
        st = torch.cuda.current_stream(self.flat_g.device).cuda_stream
        n = self.flat_p.numel()
        L.check(self.lib.avc_sqnorm(self.flat_g.data_ptr(), n, self.scratch.data_ptr(), self.sqnorm.data_ptr(), st), "sqnorm")
        L.check(self.lib.avc_adam_step(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
                                       self.flat_v.data_ptr(), self.flat_vmax.data_ptr(), n, self.hp.data_ptr(),
                                       self.sqnorm.data_ptr(), self.step_dev.data_ptr(), st), "adam_step")

    def grad_norm(self) -> torch.Tensor:
        return self.sqnorm.sqrt() / self.world_size

    # ---- checkpoint format of torch.optim.Adam
    def state_dict(self):
        t = float(self.step_dev.item())
        for p in self.param_groups[0]["params"]:
            self.state[p]["step"] = torch.tensor(t)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        step = 0.0
        loaded = {p: dict(self.state[p]) for p in self.param_groups[0]["params"] if p in self.state}
        self._views()
        for p, s in loaded.items():
            self.state[p]["exp_avg"].copy_(s["exp_avg"])
            self.state[p]["exp_avg_sq"].copy_(s["exp_avg_sq"])
            if "max_exp_avg_sq" in s:
                self.state[p]["max_exp_avg_sq"].copy_(s["max_exp_avg_sq"])
            step = max(step, float(s["step"]))
        self.step_dev.fill_(step)
        self.sync_hparams()
