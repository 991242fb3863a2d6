"""GPU diagnostic: does a sample's conv-block output depend on the batch it is computed in?
For several layer shapes, run Engine.conv on B=256 inputs and on sub-batches (1, 128 samples: other
samples-per-CTA tilings) and compare the common samples bit for bit; also run B=256 twice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_voice_conversion_b200.engine import A4, Engine
from adaptive_voice_conversion_b200 import _lib as L
import oracle.ae_oracle as orc

dev = torch.device("cuda", 0)
eng = Engine(orc.default_config(80), dev)
torch.manual_seed(0)


def sub(a: A4, n):
    t = a.t[:n].contiguous()
    return A4(t, t.data_ptr(), n, a.C, a.T, a.C * a.T, a.tf32)


def run(name, Cin, Cout, K, T, **kw):
    w = torch.randn(Cout, Cin, K, device=dev) * (1.0 / (Cin * K) ** 0.5)
    P = {"r.weight": w, "r.bias": torch.randn(Cout, device=dev) * 0.1}
    eng.conv_names = lambda: ["r"]
    eng.pack_weights(P, need_dgrad=False)
    x = A4.empty(256, Cin, T, dev)
    x.t.normal_()
    rounded = kw.pop("rounded", True)
    if rounded:   # TF32-exact input like the real model's activations
        x.t.copy_((x.t.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32))
        x.tf32 = True
    res = None
    if kw.pop("res", False):
        res = A4.empty(256, Cout, T, dev)
        res.t.normal_()
    outs = {}
    for n in (256, 256, 128, 1, 3):
        xi = sub(x, n)
        ri = sub(res, n) if res is not None else None
        y, _ = eng.conv(P, "r", xi, res=ri, res_mode=L.RES_SAME if ri is not None else L.RES_NONE, **kw)
        torch.cuda.synchronize()
        outs.setdefault(n, []).append(y.t.clone())
    ref = outs[256][0]
    line = [f"{name:34s}"]
    line.append("rerun %s" % ("bit-equal" if torch.equal(outs[256][1], ref) else "DIFF %.2e" % float((outs[256][1] - ref).abs().max())))
    for n in (128, 1, 3):
        d = (outs[n][0] - ref[:n])
        line.append(f"B={n}: " + ("bit-equal" if float(d.abs().max()) == 0 else "DIFF max %.2e rel-L2 %.2e" % (float(d.abs().max()), float(d.norm() / ref[:n].norm()))))
    print("  ".join(line), flush=True)


run("k5 128->128 T128 IN relu", 128, 128, 5, 128, norm=True, relu=True)
run("k5 128->128 T128 plain", 128, 128, 5, 128)
run("k5 128->128 T128 relu unrounded", 128, 128, 5, 128, relu=True, rounded=False)
run("k5 128->128 T64 IN relu res", 128, 128, 5, 64, norm=True, relu=True, res=True)
run("k5 128->128 T16 IN relu", 128, 128, 5, 16, norm=True, relu=True)
run("k5 128->128 T128 stride2 IN", 128, 128, 5, 128, norm=True, relu=True, stride=2)
run("k1 1104->128 T128 IN relu", 1104, 128, 1, 128, norm=True, relu=True)
run("k1 128->128 T16 plain", 128, 128, 1, 16)
run("k8 80->128 T128 relu", 80, 128, 8, 128, relu=True)
run("k5 128->256 T32 shuffle IN", 128, 256, 5, 32, norm=True, relu=True, shuffle=True)
eng.check_tc_status()
print("done")
