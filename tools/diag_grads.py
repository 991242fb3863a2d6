"""GPU diagnostic: per-tensor gradient error of (fused trainer | autograd path) vs the CPU
oracle in fp32 and fp64.  Not a test; prints a table."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle.ae_oracle as orc
from adaptive_voice_conversion_b200.model import AE
from adaptive_voice_conversion_b200.optim import FusedAdam
from adaptive_voice_conversion_b200.trainer import FusedTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = orc.default_config(80)
sd = orc.init_state(cfg, 0)
x = torch.randn((B, 80, 128), generator=torch.Generator().manual_seed(1))
eps = torch.randn((B, 128, 16), generator=torch.Generator().manual_seed(50))
lam = 0.37
o32, g32 = orc.ae_loss_and_grads(sd, cfg, x, eps, lam)
o64, g64 = orc.ae_loss_and_grads({k: v.double() for k, v in sd.items()}, cfg, x.double(), eps.double(), lam)

m = AE(cfg); m.load_state_dict(sd); m = m.cuda()
mu, ls, emb, dec = m(x.cuda(), eps=eps.cuda())
loss = 10 * (dec - x.cuda()).abs().mean() + lam * 0.5 * torch.mean(torch.exp(ls) + mu ** 2 - 1 - ls)
loss.backward()
ga = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}

m2 = AE(cfg); m2.load_state_dict(sd); m2 = m2.cuda(); m2.flatten_parameters()
opt = FusedAdam(m2, lr=5e-4, weight_decay=1e-4, max_norm=5.0)
tr = FusedTrainer(m2, opt, cfg)
tr.set_lambda_kl(lam)
outs = tr._fwd_bwd(x.cuda(), eps.cuda())
torch.cuda.synchronize()
gf = {k: v.detach().cpu().clone() for k, v in tr.G.items()}

def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))

print("dec err fused vs fp64:", rel(outs[3].cpu(), o64["dec"]), " autograd:", rel(dec.cpu(), o64["dec"]), " oracle32:", rel(o32["dec"], o64["dec"]))
rows = []
for k in sd:
    if float(g64[k].abs().max()) < 1e-6:
        continue
    rows.append((rel(gf[k], g64[k]), rel(ga[k], g64[k]), rel(g32[k], g64[k]), rel(gf[k], ga[k]), k))
rows.sort(reverse=True)
print("fused_vs_64  autograd_vs_64  oracle32_vs_64  fused_vs_autograd  name")
for r in rows[:25]:
    print("%.2e  %.2e  %.2e  %.2e  %s" % r)
