"""Driver for `ncu --set full`: a few launches of the tcgen05 fused conv block (B=256, 128->128, k=5, T=128, IN+ReLU, saves c)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_voice_conversion_b200.engine import A4, Engine
from adaptive_voice_conversion_b200.config import default_config
dev = torch.device("cuda", 0)
eng = Engine(default_config(80), dev)
w = torch.randn(128, 128, 5, device=dev) * 0.05
P = {"r.weight": w, "r.bias": torch.zeros(128, device=dev)}
eng.conv_names = lambda: ["r"]; eng.pack_weights(P, need_dgrad=False)
xs = [A4.empty(256, 128, 128, dev) for _ in range(10)]
for x in xs: x.t.normal_()
for i in range(8):
    eng.conv(P, "r", xs[i], norm=True, relu=True, train=True)
torch.cuda.synchronize()
eng.check_tc_status()
