"""Driver for `ncu --set full`: a few launches of the tcgen05 fused conv block on rotating (> L2) inputs.
usage: ncu_conv_tc.py [T] [B]   (128->128, k=5, IN+ReLU, saves c)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_voice_conversion_b200.engine import A4, Engine
from adaptive_voice_conversion_b200.config import default_config
T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
eng = Engine(default_config(80), dev)
w = torch.randn(128, 128, 5, device=dev) * 0.05
P = {"r.weight": w, "r.bias": torch.zeros(128, device=dev)}
eng.conv_names = lambda: ["r"]; eng.pack_weights(P, need_dgrad=False)
n = max(2, min(10, (160 << 20) // (B * 128 * T * 4)))
xs = [A4.empty(B, 128, T, dev) for _ in range(n)]
for x in xs:
    x.t.normal_()
    x.tf32 = True   # as in the real model: producers round, no rounding pass
keep = []
for i in range(6):
    keep.append(eng.conv(P, "r", xs[i % n], norm=True, relu=True, train=True))
torch.cuda.synchronize()
eng.check_tc_status()
