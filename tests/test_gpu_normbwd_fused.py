"""GPU: the data-gradient conv that also runs the UPSTREAM block's InstanceNorm / AdaIN / ReLU backward in its
epilogue (AVC_F_NORMBWD, csrc/conv_tc2.cu) against the separate avc_norm_bwd launch: same dc, AdaIN-row gradients,
bias gradient, weight gradients and input gradient, one launch fewer."""
import math

import pytest
import torch

import oracle.ae_oracle as orc
from test_gpu_kernels import relerr, rnd, to_a4, from_a4

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from adaptive_voice_conversion_b200.engine import Engine
    e = Engine(orc.default_config(80), torch.device("cuda", 0))
    e.precision = "tf32"
    return e


def rl2(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


# B, T, upstream norm, upstream AdaIN, residual mode of the downstream block (0 none, 1 same, 2 avg-pool, 3 upsample), need_dx
CASES = [(5, 128, 1, 0, 1, True), (5, 128, 1, 1, 0, False), (300, 16, 1, 1, 1, True), (19, 64, 0, 0, 2, True), (7, 32, 1, 1, 3, True),
         (301, 128, 0, 0, 1, True), (3, 37, 1, 0, 1, False)]


@pytest.mark.parametrize("B,T,norm,use_cond,mode,need_dx", CASES)
def test_normbwd_fused_matches_separate_launch(eng, B, T, norm, use_cond, mode, need_dx):
    from adaptive_voice_conversion_b200 import _lib as L
    C_, K = 128, 5
    x0 = rnd((B, C_, T), 1)
    P = {"up.weight": (rnd((C_, C_, K), 2) / math.sqrt(C_ * K)).cuda(), "up.bias": (rnd((C_,), 3) * 0.1).cuda(),
         "dn.weight": (rnd((C_, C_, K), 4) / math.sqrt(C_ * K)).cuda(), "dn.bias": torch.zeros(C_).cuda()}
    cond = (rnd((B, 2 * C_), 5) * 0.5 + 0.7).cuda() if use_cond else None
    eng.conv_names = lambda: ["up", "dn"]
    eng.packed.pop("up", None)
    eng.packed.pop("dn", None)
    eng.pack_weights(P, need_dgrad=True)
    eng.fold_fused = True
    y_up, rec_up = eng.conv(P, "up", to_a4(eng, x0), norm=bool(norm), cond=cond, relu=True, train=True, round_out=True)
    _, rec_dn = eng.conv(P, "dn", y_up, train=True)
    dy = to_a4(eng, rnd((B, C_, T), 6))
    res_T = {0: 0, 1: T, 2: (T + 1) // 2, 3: 2 * T}[mode]
    dres = to_a4(eng, rnd((B, C_, res_T), 7)) if mode else None
    dmode = {0: L.RES_NONE, 1: L.RES_SAME, 2: L.RES_POOL, 3: L.RES_UP}[mode]
    out = {}
    for fused in (False, True):
        eng.norm_bwd_fused = fused
        G = {k: torch.zeros_like(v) for k, v in P.items()}
        dcond = torch.zeros_like(cond) if cond is not None else None
        n0 = L.launch_count()
        f = dict(rec=rec_up, dcond=dcond, need_dx=need_dx)
        dx = eng.conv_bwd(P, G, rec_dn, dy, dres=dres, dres_mode=dmode, fuse_up=f)
        assert (f["dc"] is not None) == fused
        if fused and not need_dx:
            assert dx is None
        dx0 = eng.conv_bwd(P, G, rec_up, dx, dcond=dcond, dc_pre=f["dc"])
        eng.check_tc_status()
        out[fused] = dict(dx=None if dx is None else from_a4(eng, dx), dx0=from_a4(eng, dx0), G={k: v.clone() for k, v in G.items()},
                          dcond=None if dcond is None else dcond.clone(), n=L.launch_count() - n0)
    eng.norm_bwd_fused = False
    a, b = out[True], out[False]
    assert a["n"] < b["n"], (a["n"], b["n"])                    # no avc_norm_bwd launch for the upstream block
    if need_dx:
        assert rl2(a["dx"], b["dx"]) < 1e-6
    assert rl2(a["dx0"], b["dx0"]) < 2e-5, rl2(a["dx0"], b["dx0"])
    assert rl2(a["G"]["up.weight"], b["G"]["up.weight"]) < 2e-5
    assert rl2(a["G"]["dn.weight"], b["G"]["dn.weight"]) < 1e-6
    if cond is not None:
        assert rl2(a["dcond"], b["dcond"]) < 2e-5, rl2(a["dcond"], b["dcond"])
    if not norm:    # relu-only upstream (speaker encoder): the bias gradient is real
        assert rl2(a["G"]["up.bias"], b["G"]["up.bias"]) < 2e-5
