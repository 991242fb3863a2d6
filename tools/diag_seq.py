"""GPU diagnostic: in-sequence snapshots (dc, dW right after wgrad, dx) of two layers vs fp64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import oracle.ae_oracle as orc
from adaptive_voice_conversion_b200.model import AE
from adaptive_voice_conversion_b200.optim import FusedAdam
from adaptive_voice_conversion_b200.trainer import FusedTrainer
from adaptive_voice_conversion_b200 import engine as E

B = 8
cfg = orc.default_config(80)
sd = orc.init_state(cfg, 0)
x = torch.randn((B, 80, 128), generator=torch.Generator().manual_seed(1))
eps = torch.randn((B, 128, 16), generator=torch.Generator().manual_seed(50))
m2 = AE(cfg); m2.load_state_dict(sd); m2 = m2.cuda(); m2.flatten_parameters()
opt = FusedAdam(m2, lr=5e-4, weight_decay=1e-4, max_norm=5.0)
tr = FusedTrainer(m2, opt, cfg); tr.set_lambda_kl(0.37)
eng = tr.eng
snap, recs, dys = {}, {}, {}
def dbg(name, stage, obj):
    snap[(name, stage)] = (eng.unpack_a4(obj) if isinstance(obj, E.A4) else obj.detach()).clone()
eng.debug = dbg
orig = E.Engine.conv_bwd
def spy(self, P, G, rec, dy, **kw):
    recs[rec["name"]] = rec
    dys[rec["name"]] = (self.unpack_a4(dy).clone(), {k: (v is not None) for k, v in kw.items()})
    return orig(self, P, G, rec, dy, **kw)
E.Engine.conv_bwd = spy
tr._fwd_bwd(x.cuda(), eps.cuda())
torch.cuda.synchronize()

def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / (b.double().abs().max().cpu() + 1e-30))

for name in ["decoder.second_conv_layers.1", "content_encoder.first_conv_layers.0", "decoder.first_conv_layers.2"]:
    rec = recs[name]
    dy_pl, kw = dys[name]
    xin = eng.unpack_a4(rec["xin"]).detach().double().cpu().requires_grad_(True)
    w = tr.P[name + ".weight"].detach().double().cpu().requires_grad_(True)
    b = tr.P[name + ".bias"].detach().double().cpu().requires_grad_(True)
    c = orc.reflect_conv1d(xin, w, b, rec["stride"]); c.retain_grad()
    y = orc.instance_norm(c)
    if rec["cond"] is not None:
        y = orc.adain(y, rec["cond"].detach().double().cpu())
    y = F.relu(y)
    y.backward(dy_pl.double().cpu())
    print(f"== {name} kw={kw}")
    print(f"   in-seq dc err {rel(snap[(name,'dc')], c.grad):.2e}   dW(after wgrad) err {rel(snap[(name,'dw')], w.grad):.2e}   dW(final) err {rel(tr.G[name+'.weight'], w.grad):.2e}")
    dcs = snap[(name, 'dc')].double().cpu()
    e = (dcs - c.grad).abs()
    print("   dc err per sample:", ["%.1e" % v for v in (e.amax(dim=(1, 2)) / c.grad.abs().max()).tolist()])
    ch = e.amax(dim=(0, 2)) / c.grad.abs().max()
    print("   worst channels:", [(int(i), "%.1e" % float(ch[i])) for i in ch.argsort(descending=True)[:6]])
    bi, ci_ = divmod(int(e.amax(dim=2).argmax()), e.shape[1])
    print(f"   worst row b={bi} c={ci_}: ours {dcs[bi, ci_, :6].tolist()}  ref {c.grad[bi, ci_, :6].tolist()}")
    st = rec["stats"].cpu()
    print(f"   stats there: mean {float(st[bi, ci_, 0]):.5f} rstd {float(st[bi, ci_, 1]):.5f}  ref mean {float(c[bi, ci_].mean()):.5f} rstd {float(1/torch.sqrt(c[bi, ci_].var(unbiased=False)+1e-5)):.5f}")
