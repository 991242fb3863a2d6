"""One launch of the TMA-staged weight-gradient kernel (for compute-sanitizer)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_voice_conversion_b200.engine import A4, Engine
from adaptive_voice_conversion_b200.config import default_config
from adaptive_voice_conversion_b200 import _lib as L
dev = torch.device("cuda", 0)
eng = Engine(default_config(80), dev)
B, Cin, Cout, K, T = int(os.environ.get("DIAG_B", "4")), 128, 128, 5, int(os.environ.get("DIAG_T", "128"))
x = A4.empty(B, Cin, T, dev); x.t.normal_()
dc = A4.empty(B, Cout, T, dev); dc.t.normal_()
acc = torch.zeros(int(eng.lib.avc_wgrad_acc_floats(Cout, Cin, K)), device=dev)
wd = L.WgradDesc()
wd.B, wd.Cin, wd.Cout, wd.K, wd.stride, wd.pad_left, wd.Tin, wd.Tout = B, Cin, Cout, K, 1, K // 2, T, T
wd.x, wd.x_bstride, wd.dc, wd.dc_bstride = x.ptr, x.bstride, dc.ptr, dc.bstride
st = torch.cuda.current_stream().cuda_stream
L.check(eng.lib.avc_conv_wgrad_tc_acc(C.byref(wd), acc.data_ptr(), eng.tc_status.data_ptr(), st), "wgrad")
torch.cuda.synchronize()
print("tma ok, status", int(eng.tc_status.item()), "acc norm", float(acc.norm()))
L.set_option("wgrad_tma", False)
acc2 = torch.zeros_like(acc)
L.check(eng.lib.avc_conv_wgrad_tc_acc(C.byref(wd), acc2.data_ptr(), eng.tc_status.data_ptr(), st), "wgrad")
torch.cuda.synchronize()
print("cp.async acc norm", float(acc2.norm()), "rel diff", float((acc - acc2).norm() / acc2.norm()))
