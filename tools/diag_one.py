"""GPU diagnostic: isolate norm_bwd / wgrad / dgrad on decoder.second_conv_layers.1 and
content_encoder.first_conv_layers.0 with real network data vs fp64 autograd."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import oracle.ae_oracle as orc
from adaptive_voice_conversion_b200.model import AE
from adaptive_voice_conversion_b200.optim import FusedAdam
from adaptive_voice_conversion_b200.trainer import FusedTrainer
from adaptive_voice_conversion_b200 import engine as E, _lib as L

B = 8
cfg = orc.default_config(80)
sd = orc.init_state(cfg, 0)
x = torch.randn((B, 80, 128), generator=torch.Generator().manual_seed(1))
eps = torch.randn((B, 128, 16), generator=torch.Generator().manual_seed(50))
m2 = AE(cfg); m2.load_state_dict(sd); m2 = m2.cuda(); m2.flatten_parameters()
opt = FusedAdam(m2, lr=5e-4, weight_decay=1e-4, max_norm=5.0)
tr = FusedTrainer(m2, opt, cfg); tr.set_lambda_kl(0.37)
eng = tr.eng
recs, dys = {}, {}
orig = E.Engine.conv_bwd
def spy(self, P, G, rec, dy, **kw):
    recs[rec["name"]] = rec
    dys[rec["name"]] = (self.unpack_a4(dy).clone(), kw)
    return orig(self, P, G, rec, dy, **kw)
E.Engine.conv_bwd = spy
tr._fwd_bwd(x.cuda(), eps.cuda())
torch.cuda.synchronize()
E.Engine.conv_bwd = orig

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
    cond = rec["cond"]
    if cond is not None:
        cond64 = cond.detach().double().cpu().requires_grad_(True)
        y = orc.adain(y, cond64)
    y = F.relu(y)
    y.backward(dy_pl.double().cpu())
    # saved c vs recomputed
    c_saved = eng.unpack_a4(rec["c"])
    print(f"== {name}: T={c.shape[-1]} saved_c err {rel(c_saved, c.detach()):.2e}")
    st = rec["stats"].cpu().double()
    mu64 = c.detach().mean(2); rstd64 = 1 / torch.sqrt(c.detach().var(2, unbiased=False) + 1e-5)
    print(f"   stats mean err {rel(st[..., 0], mu64):.2e} rstd err {rel(st[..., 1], rstd64):.2e}  max rstd {float(rstd64.max()):.1f}")
    # our norm_bwd
    G = {name + ".weight": torch.zeros_like(tr.P[name + ".weight"]), name + ".bias": torch.zeros_like(tr.P[name + ".bias"])}
    dcond = torch.zeros_like(cond) if cond is not None else None
    dya = E.A4.empty(*dy_pl.shape, dy_pl.device); eng.pack_a4(dy_pl.contiguous(), dya)
    dx = orig(eng, tr.P, G, rec, dya, dcond=dcond)
    torch.cuda.synchronize()
    print(f"   isolated: dW err {rel(G[name + '.weight'], w.grad):.2e}  dx err {rel(eng.unpack_a4(dx), xin.grad):.2e}")
    # which samples are wrong in dx?
    d = (eng.unpack_a4(dx).double().cpu() - xin.grad).abs().amax(dim=(1, 2)) / xin.grad.abs().max()
    print("   dx err per sample:", ["%.1e" % v for v in d.tolist()])
    # feed exact dc to wgrad-only: emulate by calling with norm disabled
    rec2 = dict(rec); rec2["norm"] = False; rec2["relu"] = False; rec2["cond"] = None
    dca = E.A4.empty(B, rec["Cout"], rec["Tout"], dy_pl.device); eng.pack_a4(c.grad.float().cuda().contiguous(), dca)
    G2 = {k: torch.zeros_like(v) for k, v in G.items()}
    dx2 = orig(eng, tr.P, G2, rec2, dca)
    print(f"   exact-dc -> wgrad err {rel(G2[name + '.weight'], w.grad):.2e}  dgrad err {rel(eng.unpack_a4(dx2), xin.grad):.2e}")
