"""GPU diagnostic: per-layer dy (grad w.r.t. each conv block's pre-residual output) of the
engine's hand-written backward vs fp64 autograd of the oracle ops, content encoder + decoder."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import oracle.ae_oracle as orc
from adaptive_voice_conversion_b200.model import AE
from adaptive_voice_conversion_b200.optim import FusedAdam
from adaptive_voice_conversion_b200.trainer import FusedTrainer
from adaptive_voice_conversion_b200 import engine as E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = orc.default_config(80)
sd = orc.init_state(cfg, 0)
x = torch.randn((B, 80, 128), generator=torch.Generator().manual_seed(1))
eps = torch.randn((B, 128, 16), generator=torch.Generator().manual_seed(50))
lam = 0.37
sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
ys = {}

def keep(name, t):
    t.retain_grad(); ys[name] = t; return t

def content(sd, x):
    p = "content_encoder"
    out = orc.conv_bank_cat(x, sd, p, 8)
    out = keep(p + ".in_conv_layer", F.relu(orc.instance_norm(orc.reflect_conv1d(out, sd[p + ".in_conv_layer.weight"], sd[p + ".in_conv_layer.bias"]))))
    for l, s in enumerate(cfg["ContentEncoder"]["subsample"]):
        y = keep(f"{p}.first_conv_layers.{l}", F.relu(orc.instance_norm(orc.reflect_conv1d(out, sd[f"{p}.first_conv_layers.{l}.weight"], sd[f"{p}.first_conv_layers.{l}.bias"]))))
        y = keep(f"{p}.second_conv_layers.{l}", F.relu(orc.instance_norm(orc.reflect_conv1d(y, sd[f"{p}.second_conv_layers.{l}.weight"], sd[f"{p}.second_conv_layers.{l}.bias"], stride=s))))
        if s > 1:
            out = F.avg_pool1d(out, kernel_size=s, ceil_mode=True)
        out = y + out
    mu = keep(p + ".mean_layer", orc.reflect_conv1d(out, sd[p + ".mean_layer.weight"], sd[p + ".mean_layer.bias"]))
    ls = keep(p + ".std_layer", orc.reflect_conv1d(out, sd[p + ".std_layer.weight"], sd[p + ".std_layer.bias"]))
    return mu, ls

def decoder(sd, z, cond):
    p = "decoder"
    out = keep(p + ".in_conv_layer", F.relu(orc.instance_norm(orc.reflect_conv1d(z, sd[p + ".in_conv_layer.weight"], sd[p + ".in_conv_layer.bias"]))))
    for l, up in enumerate(cfg["Decoder"]["upsample"]):
        y = orc.instance_norm(orc.reflect_conv1d(out, sd[f"{p}.first_conv_layers.{l}.weight"], sd[f"{p}.first_conv_layers.{l}.bias"]))
        y = keep(f"{p}.first_conv_layers.{l}", F.relu(orc.adain(y, F.linear(cond, sd[f"{p}.conv_affine_layers.{2*l}.weight"], sd[f"{p}.conv_affine_layers.{2*l}.bias"]))))
        y = orc.reflect_conv1d(y, sd[f"{p}.second_conv_layers.{l}.weight"], sd[f"{p}.second_conv_layers.{l}.bias"])
        if up > 1:
            y = orc.pixel_shuffle_1d(y, up)
        y = orc.instance_norm(y)
        y = keep(f"{p}.second_conv_layers.{l}", F.relu(orc.adain(y, F.linear(cond, sd[f"{p}.conv_affine_layers.{2*l+1}.weight"], sd[f"{p}.conv_affine_layers.{2*l+1}.bias"]))))
        out = y + (F.interpolate(out, scale_factor=up, mode="nearest") if up > 1 else out)
    return keep(p + ".out_conv_layer", orc.reflect_conv1d(out, sd[p + ".out_conv_layer.weight"], sd[p + ".out_conv_layer.bias"]))

x64, eps64 = x.double(), eps.double()
emb = orc.speaker_encoder(sd64, x64, cfg["SpeakerEncoder"]["subsample"])
mu, ls = content(sd64, x64)
dec = decoder(sd64, mu + torch.exp(ls / 2) * eps64, emb)
lr, lk = orc.ae_losses(x64, mu, ls, dec)
(10 * lr + lam * lk).backward()

m2 = AE(cfg); m2.load_state_dict(sd); m2 = m2.cuda(); m2.flatten_parameters()
opt = FusedAdam(m2, lr=5e-4, weight_decay=1e-4, max_norm=5.0)
tr = FusedTrainer(m2, opt, cfg)
tr.set_lambda_kl(lam)
eng = tr.eng
captured = {}
orig = E.Engine.conv_bwd
def spy(self, P, G, rec, dy, **kw):
    captured[rec["name"]] = self.unpack_a4(dy).cpu() if dy.bstride == dy.C * dy.T else None
    dx = orig(self, P, G, rec, dy, **kw)
    if dx is not None:
        captured[rec["name"] + "#dx"] = self.unpack_a4(dx).cpu()
    return dx
E.Engine.conv_bwd = spy
tr._fwd_bwd(x.cuda(), eps.cuda())
torch.cuda.synchronize()

def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))
print("%-44s %10s %10s" % ("layer (backward order)", "dy err", "dW err"))
order = ["decoder.out_conv_layer"] + [f"decoder.{w}_conv_layers.{l}" for l in reversed(range(6)) for w in ("second", "first")] + ["decoder.in_conv_layer",
         "content_encoder.mean_layer", "content_encoder.std_layer"] + [f"content_encoder.{w}_conv_layers.{l}" for l in reversed(range(6)) for w in ("second", "first")] + ["content_encoder.in_conv_layer"]
for n in order:
    dy = captured.get(n)
    e1 = rel(dy, ys[n].grad) if dy is not None else float("nan")
    e2 = rel(tr.G[n + ".weight"].cpu(), sd64[n + ".weight"].grad)
    print("%-44s %10.2e %10.2e  T=%d" % (n, e1, e2, ys[n].shape[-1]))
