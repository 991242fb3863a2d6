"""GPU diagnostic: per-role wait / work cycles of the persistent conv kernel (conv_tc2.cu, avc_tc2_set_debug)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_voice_conversion_b200.engine import A4, Engine
from adaptive_voice_conversion_b200.config import default_config
dev = torch.device("cuda", 0)
eng = Engine(default_config(80), dev)
B = int(os.environ.get("DIAG_B", "256"))
for (Cin, Cout, K, T, kw, tag) in [(128, 128, 5, 128, dict(norm=True, relu=True), "conv5 T128 IN"), (128, 128, 5, 128, dict(), "conv5 T128 plain"),
                                   (1104, 128, 1, 128, dict(norm=True, relu=True), "in_conv"), (128, 128, 5, 64, dict(norm=True, relu=True), "conv5 T64"),
                                   (128, 128, 5, 16, dict(norm=True, relu=True), "conv5 T16"), (80, 128, 8, 128, dict(relu=True), "bank k8")]:
    w = torch.randn(Cout, Cin, K, device=dev) * 0.05
    P = {"r.weight": w, "r.bias": torch.zeros(Cout, device=dev)}
    eng.packed.pop("r", None); eng.conv_names = lambda: ["r"]; eng.pack_weights(P, need_dgrad=False)
    x = A4.empty(B, Cin, T, dev); x.t.normal_()
    x.tf32 = True   # as in the real model (producers round): the kernel skips its rounding pass
    train = bool(kw)
    for _ in range(3):
        eng.conv(P, "r", x, train=train, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.conv(P, "r", x, train=train, **kw)
    e1.record(); torch.cuda.synchronize()
    from adaptive_voice_conversion_b200 import _lib as L
    L.set_option("tc_conv_v2", False)
    try:
        for _ in range(3):
            eng.conv(P, "r", x, train=train, **kw)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(10):
            eng.conv(P, "r", x, train=train, **kw)
        f1.record(); torch.cuda.synchronize()
        v1_us = f0.elapsed_time(f1) * 100
    except Exception as e:
        v1_us = float("nan")
    L.set_option("tc_conv_v2", True)
    dbg = torch.zeros(16 * 256, dtype=torch.int64, device=dev)
    eng.lib.avc_tc2_set_debug(dbg.data_ptr())
    eng.conv(P, "r", x, train=train, **kw)
    torch.cuda.synchronize()
    eng.lib.avc_tc2_set_debug(None)
    t = dbg.view(-1, 16).cpu()
    t = t[t[:, 0] != 0]
    f = lambda i: float(t[:, i].float().mean())
    span = float(t[:, 1].max() - t[:, 0].min())
    print(f"{tag:18s} {e0.elapsed_time(e1) * 100:7.1f} us/launch (hot L2; round-1 kernel {v1_us:7.1f} us)  CTAs {len(t)}  tiles/CTA {f(12):.2f}  CTA life {float((t[:,1]-t[:,0]).float().mean()):8.0f} cyc  span {span:8.0f}")
    print(f"{'':18s} producer wait-empty {f(2):7.0f} | patch wait-full {f(3):7.0f} work {f(4):7.0f} | mma wait-ready {f(5):7.0f} wait-acc {f(6):7.0f} issue {f(7):7.0f}"
          f" | epi wait-acc {f(8):7.0f} tmem-pass {f(9):7.0f} params {f(10):7.0f} store-pass {f(11):7.0f}", flush=True)
eng.check_tc_status()
