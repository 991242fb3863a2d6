"""CPU oracle for the AdaIN-VC hot path -- TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (functional style, torch CPU fp32, driven by a plain
``state_dict``) of the algorithm in the reference's ``model.py`` / ``solver.py``.
It is the checker for the CUDA path; it is never the product:

  * only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
    ``--impl reference`` legs of ``bench.py`` may import it;
  * nothing under ``adaptive_voice_conversion_b200/`` imports it, and the product
    path raises if the CUDA library is missing (no CPU fallback).

Parity pinning: the reference ships no tests / golden vectors, so
the oracle is pinned against the *reference itself*: ``oracle/make_golden.py`` imports
``/root/reference/model.py`` in the authoring container, runs it on seeded inputs and
commits the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
restatement against those fixtures (and, when ``/root/reference`` is present, against
the live reference).

Each function cites the reference file:line it restates.  Tensors are ``[B, C, T]``
(channels first, time contiguous) exactly like the reference.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

IN_EPS = 1e-5  # nn.InstanceNorm1d default eps (model.py:296,341)


# --------------------------------------------------------------------------- helpers
def reflect_conv1d(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int = 1) -> Tensor:
    """pad_layer (model.py:21-32): reflect pad (k//2, k//2) for odd k, (k//2, k//2-1)
    for even k, then a valid Conv1d."""
    k = w.shape[-1]
    left = k // 2
    right = k // 2 - 1 if k % 2 == 0 else k // 2
    if left or right:
        x = F.pad(x, (left, right), mode="reflect")
    return F.conv1d(x, w, b, stride=stride)


def instance_norm(x: Tensor) -> Tensor:
    """nn.InstanceNorm1d(affine=False) (model.py:296,341): per (b, c) biased variance
    over T, eps 1e-5, no running statistics."""
    mu = x.mean(dim=2, keepdim=True)
    var = x.var(dim=2, unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + IN_EPS)


def pixel_shuffle_1d(x: Tensor, r: int) -> Tensor:
    """pixel_shuffle_1d (model.py:52-59): out[b, c, r*t + s] = in[b, r*c + s, t]."""
    b, c, t = x.shape
    return x.reshape(b, c // r, r, t).transpose(2, 3).reshape(b, c // r, t * r)


def adain(x: Tensor, cond: Tensor) -> Tensor:
    """append_cond (model.py:77-83): first half of cond is the additive term, second
    half the multiplicative one."""
    p = cond.shape[1] // 2
    return x * cond[:, p:, None] + cond[:, :p, None]


def conv_bank_cat(x: Tensor, sd: State, prefix: str, n_bank: int) -> Tensor:
    """conv_bank (model.py:85-91): ReLU(conv_k(x)) for every bank conv, concatenated
    with x itself along channels."""
    outs = [F.relu(reflect_conv1d(x, sd[f"{prefix}.conv_bank.{i}.weight"],
                                  sd[f"{prefix}.conv_bank.{i}.bias"])) for i in range(n_bank)]
    return torch.cat(outs + [x], dim=1)


def _count(sd: State, pattern: str) -> int:
    n = 0
    while pattern.format(n) in sd:
        n += 1
    return n


# --------------------------------------------------------------------------- stacks
def speaker_encoder(sd: State, x: Tensor, subsample: Sequence[int], prefix: str = "speaker_encoder") -> Tensor:
    """SpeakerEncoder.forward (model.py:265-277) incl. conv_blocks (:237-250) and
    dense_blocks (:252-263).  ReLU activations, dropout p=0 (config.yaml:11-12)."""
    p = prefix
    out = conv_bank_cat(x, sd, p, _count(sd, p + ".conv_bank.{}.weight"))
    out = F.relu(reflect_conv1d(out, sd[f"{p}.in_conv_layer.weight"], sd[f"{p}.in_conv_layer.bias"]))
    for l, s in enumerate(subsample):
        y = F.relu(reflect_conv1d(out, sd[f"{p}.first_conv_layers.{l}.weight"], sd[f"{p}.first_conv_layers.{l}.bias"]))
        y = F.relu(reflect_conv1d(y, sd[f"{p}.second_conv_layers.{l}.weight"], sd[f"{p}.second_conv_layers.{l}.bias"], stride=s))
        if s > 1:
            out = F.avg_pool1d(out, kernel_size=s, ceil_mode=True)
        out = y + out
    out = out.mean(dim=2)  # AdaptiveAvgPool1d(1).squeeze(2)
    for l in range(_count(sd, p + ".first_dense_layers.{}.weight")):
        y = F.relu(F.linear(out, sd[f"{p}.first_dense_layers.{l}.weight"], sd[f"{p}.first_dense_layers.{l}.bias"]))
        y = F.relu(F.linear(y, sd[f"{p}.second_dense_layers.{l}.weight"], sd[f"{p}.second_dense_layers.{l}.bias"]))
        out = y + out
    return F.linear(out, sd[f"{p}.output_layer.weight"], sd[f"{p}.output_layer.bias"])


def content_encoder(sd: State, x: Tensor, subsample: Sequence[int], prefix: str = "content_encoder") -> Tuple[Tensor, Tensor]:
    """ContentEncoder.forward (model.py:301-323)."""
    p = prefix
    out = conv_bank_cat(x, sd, p, _count(sd, p + ".conv_bank.{}.weight"))
    out = reflect_conv1d(out, sd[f"{p}.in_conv_layer.weight"], sd[f"{p}.in_conv_layer.bias"])
    out = F.relu(instance_norm(out))
    for l, s in enumerate(subsample):
        y = reflect_conv1d(out, sd[f"{p}.first_conv_layers.{l}.weight"], sd[f"{p}.first_conv_layers.{l}.bias"])
        y = F.relu(instance_norm(y))
        y = reflect_conv1d(y, sd[f"{p}.second_conv_layers.{l}.weight"], sd[f"{p}.second_conv_layers.{l}.bias"], stride=s)
        y = F.relu(instance_norm(y))
        if s > 1:
            out = F.avg_pool1d(out, kernel_size=s, ceil_mode=True)
        out = y + out
    mu = reflect_conv1d(out, sd[f"{p}.mean_layer.weight"], sd[f"{p}.mean_layer.bias"])
    log_sigma = reflect_conv1d(out, sd[f"{p}.std_layer.weight"], sd[f"{p}.std_layer.bias"])
    return mu, log_sigma


def decoder(sd: State, z: Tensor, cond: Tensor, upsample: Sequence[int], prefix: str = "decoder") -> Tensor:
    """Decoder.forward (model.py:347-371)."""
    p = prefix
    out = reflect_conv1d(z, sd[f"{p}.in_conv_layer.weight"], sd[f"{p}.in_conv_layer.bias"])
    out = F.relu(instance_norm(out))
    for l, up in enumerate(upsample):
        y = reflect_conv1d(out, sd[f"{p}.first_conv_layers.{l}.weight"], sd[f"{p}.first_conv_layers.{l}.bias"])
        y = instance_norm(y)
        y = F.relu(adain(y, F.linear(cond, sd[f"{p}.conv_affine_layers.{2 * l}.weight"], sd[f"{p}.conv_affine_layers.{2 * l}.bias"])))
        y = reflect_conv1d(y, sd[f"{p}.second_conv_layers.{l}.weight"], sd[f"{p}.second_conv_layers.{l}.bias"])
        if up > 1:
            y = pixel_shuffle_1d(y, up)
        y = instance_norm(y)
        y = F.relu(adain(y, F.linear(cond, sd[f"{p}.conv_affine_layers.{2 * l + 1}.weight"], sd[f"{p}.conv_affine_layers.{2 * l + 1}.bias"])))
        if up > 1:
            out = y + F.interpolate(out, scale_factor=up, mode="nearest")
        else:
            out = y + out
    return reflect_conv1d(out, sd[f"{p}.out_conv_layer.weight"], sd[f"{p}.out_conv_layer.bias"])


# --------------------------------------------------------------------------- AE
def ae_forward(sd: State, config: dict, x: Tensor, eps: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """AE.forward (model.py:380-385) with the N(0,1) draw ``eps`` injected."""
    emb = speaker_encoder(sd, x, config["SpeakerEncoder"]["subsample"])
    mu, log_sigma = content_encoder(sd, x, config["ContentEncoder"]["subsample"])
    z = mu + torch.exp(log_sigma / 2) * eps
    dec = decoder(sd, z, emb, config["Decoder"]["upsample"])
    return mu, log_sigma, emb, dec


def ae_inference(sd: State, config: dict, x: Tensor, x_cond: Tensor) -> Tensor:
    """AE.inference (model.py:387-391): speaker from x_cond, content mean from x."""
    emb = speaker_encoder(sd, x_cond, config["SpeakerEncoder"]["subsample"])
    mu, _ = content_encoder(sd, x, config["ContentEncoder"]["subsample"])
    return decoder(sd, mu, emb, config["Decoder"]["upsample"])


def ae_losses(x: Tensor, mu: Tensor, log_sigma: Tensor, dec: Tensor) -> Tuple[Tensor, Tensor]:
    """solver.py:84-86: L1 reconstruction (mean) and KL = 0.5*mean(exp(ls)+mu^2-1-ls)."""
    loss_rec = (dec - x).abs().mean()
    loss_kl = 0.5 * torch.mean(torch.exp(log_sigma) + mu ** 2 - 1 - log_sigma)
    return loss_rec, loss_kl


def ae_loss_and_grads(sd: State, config: dict, x: Tensor, eps: Tensor, lambda_kl: float):
    """solver.py:83-90: total loss and d(loss)/d(param) for every tensor of ``sd``."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    mu, log_sigma, emb, dec = ae_forward(leaves, config, x, eps)
    loss_rec, loss_kl = ae_losses(x, mu, log_sigma, dec)
    loss = config["lambda"]["lambda_rec"] * loss_rec + lambda_kl * loss_kl
    names = list(leaves)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    gd = {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, grads)}
    outs = dict(mu=mu.detach(), log_sigma=log_sigma.detach(), emb=emb.detach(), dec=dec.detach(),
                loss_rec=loss_rec.detach(), loss_kl=loss_kl.detach())
    return outs, gd


class AdamState:
    """torch.optim.Adam(amsgrad=True, weight_decay=wd) state (solver.py:75-77),
    restated per tensor: L2 decay folded into the gradient, bias-corrected, max of the
    second moment."""

    def __init__(self, sd: State):
        self.step = 0
        self.m = {k: torch.zeros_like(v) for k, v in sd.items()}
        self.v = {k: torch.zeros_like(v) for k, v in sd.items()}
        self.vmax = {k: torch.zeros_like(v) for k, v in sd.items()}


def clip_and_adam(sd: State, grads: State, st: AdamState, opt_cfg: dict) -> float:
    """solver.py:91-93: clip_grad_norm_(max_norm) then Adam.step().  Updates ``sd`` in
    place and returns the pre-clip global gradient norm."""
    lr, b1, b2 = opt_cfg["lr"], opt_cfg["beta1"], opt_cfg["beta2"]
    wd, max_norm, eps = opt_cfg["weight_decay"], opt_cfg["grad_norm"], 1e-8
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
    coef = min(1.0, max_norm / (total + 1e-6))
    st.step += 1
    bc1 = 1 - b1 ** st.step
    bc2 = 1 - b2 ** st.step
    for k in sd:
        g = grads[k] * coef + wd * sd[k]
        st.m[k].mul_(b1).add_(g, alpha=1 - b1)
        st.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        if opt_cfg.get("amsgrad", True):
            torch.maximum(st.vmax[k], st.v[k], out=st.vmax[k])
            second = st.vmax[k]
        else:
            second = st.v[k]
        denom = second.sqrt() / math.sqrt(bc2) + eps
        sd[k].addcdiv_(st.m[k], denom, value=-lr / bc1)
    return total


def ae_train_step(sd: State, st: AdamState, config: dict, x: Tensor, eps: Tensor, lambda_kl: float) -> dict:
    """Solver.ae_step (solver.py:81-97) on CPU with injected eps; mutates sd / st."""
    outs, grads = ae_loss_and_grads(sd, config, x, eps, lambda_kl)
    gnorm = clip_and_adam(sd, grads, st, config["optimizer"])
    return {"loss_rec": float(outs["loss_rec"]), "loss_kl": float(outs["loss_kl"]), "grad_norm": gnorm,
            "outs": outs, "grads": grads}


# --------------------------------------------------------------------------- init / config
def default_config(c_in: int = 80) -> dict:
    """config.yaml of the reference with c_in/c_out overridable (BASELINE uses
    80 mels, the shipped file 512)."""
    enc = dict(c_in=c_in, c_h=128, c_out=128, kernel_size=5, bank_size=8, bank_scale=1, c_bank=128,
               n_conv_blocks=6, subsample=[1, 2, 1, 2, 1, 2], act="relu", dropout_rate=0)
    return {
        "SpeakerEncoder": dict(enc, n_dense_blocks=6),
        "ContentEncoder": dict(enc),
        "Decoder": dict(c_in=128, c_cond=128, c_h=128, c_out=c_in, kernel_size=5, n_conv_blocks=6,
                        upsample=[2, 1, 2, 1, 2, 1], act="relu", sn=False, dropout_rate=0),
        "data_loader": dict(segment_size=128, frame_size=1, batch_size=128, shuffle=True),
        "optimizer": dict(lr=0.0005, beta1=0.9, beta2=0.999, amsgrad=True, weight_decay=0.0001, grad_norm=5),
        "lambda": dict(lambda_rec=10, lambda_kl=1),
        "annealing_iters": 20000,
    }


def param_shapes(config: dict):
    """Names and shapes of the 166 tensors of AE.state_dict() in registration order
    (model.py:210-235, 280-299, 326-345)."""
    out = []

    def conv(name, co, ci, k):
        out.append((name + ".weight", (co, ci, k)))
        out.append((name + ".bias", (co,)))

    def lin(name, o, i):
        out.append((name + ".weight", (o, i)))
        out.append((name + ".bias", (o,)))

    for enc_name, key in (("speaker_encoder", "SpeakerEncoder"), ("content_encoder", "ContentEncoder")):
        c = config[key]
        ks = list(range(c["bank_scale"], c["bank_size"] + 1, c["bank_scale"]))
        for i, k in enumerate(ks):
            conv(f"{enc_name}.conv_bank.{i}", c["c_bank"], c["c_in"], k)
        conv(f"{enc_name}.in_conv_layer", c["c_h"], c["c_bank"] * len(ks) + c["c_in"], 1)
        for l in range(c["n_conv_blocks"]):
            conv(f"{enc_name}.first_conv_layers.{l}", c["c_h"], c["c_h"], c["kernel_size"])
        for l in range(c["n_conv_blocks"]):
            conv(f"{enc_name}.second_conv_layers.{l}", c["c_h"], c["c_h"], c["kernel_size"])
        if key == "SpeakerEncoder":
            for l in range(c["n_dense_blocks"]):
                lin(f"{enc_name}.first_dense_layers.{l}", c["c_h"], c["c_h"])
            for l in range(c["n_dense_blocks"]):
                lin(f"{enc_name}.second_dense_layers.{l}", c["c_h"], c["c_h"])
            lin(f"{enc_name}.output_layer", c["c_out"], c["c_h"])
        else:
            conv(f"{enc_name}.mean_layer", c["c_out"], c["c_h"], 1)
            conv(f"{enc_name}.std_layer", c["c_out"], c["c_h"], 1)
    d = config["Decoder"]
    conv("decoder.in_conv_layer", d["c_h"], d["c_in"], 1)
    for l in range(d["n_conv_blocks"]):
        conv(f"decoder.first_conv_layers.{l}", d["c_h"], d["c_h"], d["kernel_size"])
    for l, up in zip(range(d["n_conv_blocks"]), d["upsample"]):
        conv(f"decoder.second_conv_layers.{l}", d["c_h"] * up, d["c_h"], d["kernel_size"])
    for l in range(2 * d["n_conv_blocks"]):
        lin(f"decoder.conv_affine_layers.{l}", d["c_h"] * 2, d["c_cond"])
    conv("decoder.out_conv_layer", d["c_out"], d["c_h"], 1)
    return out


def init_state(config: dict, seed: int = 0) -> State:
    """Deterministic random state_dict with the fan-in-uniform scale of nn.Conv1d /
    nn.Linear default init (bound 1/sqrt(fan_in) for weight and bias).  Portable (does not
    need the reference) -- used for seeded GPU-vs-oracle tests and the synthetic bench."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    bound = 1.0
    for name, shape in param_shapes(config):
        if name.endswith(".weight"):
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(fan_in)
        sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


class TorchOptimStep:
    """Solver.ae_step exactly as the reference drives it (solver.py:81-93): the oracle's
    functional forward, ``loss.backward()``, ``clip_grad_norm_`` and the stock
    ``torch.optim.Adam(amsgrad, weight_decay)``.  This is the CPU leg ``bench.py`` times
    (stock torch kernels end to end, like the reference) and a second opinion for
    ``clip_and_adam``."""

    def __init__(self, sd: State, config: dict):
        self.config = config
        self.params = {k: torch.nn.Parameter(v.detach().clone()) for k, v in sd.items()}
        o = config["optimizer"]
        self.opt = torch.optim.Adam(list(self.params.values()), lr=o["lr"], betas=(o["beta1"], o["beta2"]),
                                    amsgrad=o["amsgrad"], weight_decay=o["weight_decay"])

    def step(self, x: Tensor, eps: Tensor, lambda_kl: float) -> dict:
        mu, log_sigma, emb, dec = ae_forward(self.params, self.config, x, eps)
        loss_rec, loss_kl = ae_losses(x, mu, log_sigma, dec)
        loss = self.config["lambda"]["lambda_rec"] * loss_rec + lambda_kl * loss_kl
        self.opt.zero_grad()
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(list(self.params.values()), max_norm=self.config["optimizer"]["grad_norm"])
        self.opt.step()
        return {"loss_rec": float(loss_rec.detach()), "loss_kl": float(loss_kl.detach()), "grad_norm": float(gn)}
