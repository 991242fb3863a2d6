"""GPU diagnostic: (1) MN-major tcgen05 descriptor variants, (2) full-model error and speed of the
tf32 path vs fp32 path vs the CPU oracle."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import oracle.ae_oracle as orc
from test_gpu_tc import run_probe, rel

g = torch.Generator().manual_seed(9)
Kt, N = 64, 128
At, Bt = torch.randn((Kt, 128), generator=g), torch.randn((Kt, N), generator=g)
a_img = At.reshape(Kt, 32, 4).permute(1, 0, 2).contiguous()
b_img = Bt.reshape(Kt, N // 4, 4).permute(1, 0, 2).contiguous()
ref = At.t() @ Bt
print("MN-major variants (both operands MN-major), rel err:")
for name, st in {
    "lbo=128,sbo=K*16,kstep=128": [128, Kt * 16, 128, Kt * 16, 128, 128, 0, 0],
    "lbo=K*16,sbo=128,kstep=128": [Kt * 16, 128, Kt * 16, 128, 128, 128, 0, 0],
    "lbo=0,sbo=K*16": [0, Kt * 16, 0, Kt * 16, 128, 128, 0, 0],
    "lbo=K*16,sbo=0": [Kt * 16, 0, Kt * 16, 0, 128, 128, 0, 0],
    "lbo=16,sbo=K*16": [16, Kt * 16, 16, Kt * 16, 128, 128, 0, 0],
    "lbo=K*16,sbo=16": [Kt * 16, 16, Kt * 16, 16, 128, 128, 0, 0],
}.items():
    try:
        D = run_probe(a_img, b_img, st, Kt // 8, N, a_mn=1, b_mn=1)
        print("  %-32s %.3e  (D absmax %.3f)" % (name, rel(D, ref), float(D.abs().max())))
    except AssertionError as e:
        print("  %-32s FAILED %s" % (name, e))
# A MN-major with B K-major and vice versa (B image K-major: [k/4][n][4])
Bk = Bt.t().contiguous()  # [N][Kt]
b_img_k = Bk.reshape(N, Kt // 4, 4).permute(1, 0, 2).contiguous()
Ak = At.t().contiguous()
a_img_k = Ak.reshape(128, Kt // 4, 4).permute(1, 0, 2).contiguous()
for name, (ai, bi, st, amn, bmn) in {
    "A MN / B K": (a_img, b_img_k, [128, Kt * 16, N * 16, 128, 128, 2 * N * 16, 0, 0], 1, 0),
    "A MN(swapped) / B K": (a_img, b_img_k, [Kt * 16, 128, N * 16, 128, 128, 2 * N * 16, 0, 0], 1, 0),
    "A K / B MN": (a_img_k, b_img, [128 * 16, 128, 128, Kt * 16, 2 * 128 * 16, 128, 0, 0], 0, 1),
    "A K / B MN(swapped)": (a_img_k, b_img, [128 * 16, 128, Kt * 16, 128, 2 * 128 * 16, 128, 0, 0], 0, 1),
}.items():
    try:
        D = run_probe(ai, bi, st, Kt // 8, N, a_mn=amn, b_mn=bmn)
        print("  %-32s %.3e  (D absmax %.3f)" % (name, rel(D, ref), float(D.abs().max())))
    except AssertionError as e:
        print("  %-32s FAILED %s" % (name, e))

# ---- full model
from adaptive_voice_conversion_b200.model import AE
cfg = orc.default_config(80)
sd = orc.init_state(cfg, 0)
B = 16
x = torch.randn((B, 80, 128), generator=torch.Generator().manual_seed(1))
eps = torch.randn((B, 128, 16), generator=torch.Generator().manual_seed(2))
o, gr = orc.ae_loss_and_grads(sd, cfg, x, eps, 1.0)
for prec in ("fp32", "tf32"):
    os.environ["AVC_PRECISION"] = prec
    m = AE(cfg); m.load_state_dict(sd); m = m.cuda()
    mu, ls, emb, dec = m(x.cuda(), eps=eps.cuda())
    lr = (dec - x.cuda()).abs().mean(); lk = 0.5 * torch.mean(torch.exp(ls) + mu ** 2 - 1 - ls)
    (10 * lr + lk).backward()
    m.engine("cuda:0").check_tc_status()
    num = den = 0.0
    for k, p in m.named_parameters():
        num += float((p.grad.cpu().double() - gr[k].double()).pow(2).sum()); den += float(gr[k].double().pow(2).sum())
    r = lambda a, b: float((a.detach().cpu() - b).abs().max() / b.abs().max())
    print(f"[{prec}] mu {r(mu, o['mu']):.2e} ls {r(ls, o['log_sigma']):.2e} emb {r(emb, o['emb']):.2e} dec {r(dec, o['dec']):.2e} | "
          f"loss_rec rel {abs(float(lr) - float(o['loss_rec'])) / float(o['loss_rec']):.2e} loss_kl rel {abs(float(lk) - float(o['loss_kl'])) / float(o['loss_kl']):.2e} | grad relL2 {(num / den) ** 0.5:.2e}")
