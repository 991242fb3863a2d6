"""GPU: the fused speaker dense stack and the batched AdaIN affine layers (csrc/dense_fused.cu)
against plain torch fp32 autograd of the reference op sequence (model.py:252-263, :273-276,
:342-343).  fp32 FFMA kernels: tolerance 2e-4 of the tensor max."""
import pytest
import torch

import oracle.ae_oracle as orc
from test_gpu_kernels import relerr, rnd

pytestmark = pytest.mark.gpu
TOL = 2e-4


@pytest.fixture(scope="module")
def eng():
    from adaptive_voice_conversion_b200.engine import Engine
    e = Engine(orc.default_config(80), torch.device("cuda", 0))
    e.precision = "fp32"
    return e


@pytest.mark.parametrize("B", [1, 6, 257])
def test_dense_stack_fwd_bwd(eng, B):
    import ctypes as C
    from adaptive_voice_conversion_b200 import _lib as L
    nb, Cc = 6, 128
    names = [f"f{l}" for l in range(nb)] + [f"s{l}" for l in range(nb)] + ["o"]
    P = {}
    for i, n in enumerate(names):
        P[n + ".weight"] = rnd((Cc, Cc), 10 + i) / 11
        P[n + ".bias"] = rnd((Cc,), 40 + i) * 0.1
    x, dout = rnd((B, Cc), 1), rnd((B, Cc), 2)
    ref = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    xr = x.clone().requires_grad_(True)
    h = xr
    for l in range(nb):
        y = torch.relu(h @ ref[f"f{l}.weight"].T + ref[f"f{l}.bias"])
        h = torch.relu(y @ ref[f"s{l}.weight"].T + ref[f"s{l}.bias"]) + h
    out = h @ ref["o.weight"].T + ref["o.bias"]
    out.backward(dout)

    Pd = {k: v.cuda() for k, v in P.items()}
    Gd = {k: torch.zeros_like(v) for k, v in Pd.items()}
    order = [n + s for n in names for s in (".weight", ".bias")]
    tab = eng._ptr_table(("t_dense", B), [Pd[k] for k in order])
    gtab = eng._ptr_table(("t_dense_g", B), [Gd[k] for k in order])
    xd, doutd = x.cuda(), dout.cuda()
    save, emb = eng.empty(3 * nb + 1, B, Cc), eng.empty(B, Cc)
    gsave, dx = eng.empty(2 * nb + 1, B, Cc), eng.empty(B, Cc)
    d = L.DenseStackDesc()
    d.B, d.C, d.c_out, d.n_blocks = B, Cc, Cc, nb
    d.params, d.x, d.save, d.out = tab.data_ptr(), xd.data_ptr(), save.data_ptr(), emb.data_ptr()
    d.dout, d.gsave, d.dx = doutd.data_ptr(), gsave.data_ptr(), dx.data_ptr()
    L.check(eng.lib.avc_dense_stack_fwd(C.byref(d), eng.stream), "fwd")
    L.check(eng.lib.avc_dense_stack_bwd(C.byref(d), eng.stream), "bwd")
    plane = B * Cc
    slots = [(l, l) for l in range(nb)] + [(nb + l, nb + 1 + l) for l in range(nb)] + [(2 * nb, nb)]
    bd = L.LinearBatchDesc()
    bd.L, bd.B, bd.N, bd.K = len(slots), B, Cc, Cc
    bd.grads, bd.x, bd.x_bstride, bd.y, bd.y_bstride = gtab.data_ptr(), save.data_ptr(), Cc, gsave.data_ptr(), Cc
    for i, (gs, xs) in enumerate(slots):
        bd.y_off[i], bd.x_off[i] = gs * plane, xs * plane
    L.check(eng.lib.avc_linear_batch_dw(C.byref(bd), eng.stream), "dw")
    torch.cuda.synchronize()
    assert relerr(emb, out) < TOL
    assert relerr(dx, xr.grad) < TOL
    for k in order:
        assert relerr(Gd[k], ref[k].grad) < TOL, k
    # inference form: no save buffer
    d.save = None
    emb2 = eng.empty(B, Cc)
    d.out = emb2.data_ptr()
    L.check(eng.lib.avc_dense_stack_fwd(C.byref(d), eng.stream), "fwd(no save)")
    assert torch.equal(emb2, emb)


@pytest.mark.parametrize("B", [3, 256])
def test_linear_batch_affine(eng, B):
    import ctypes as C
    from adaptive_voice_conversion_b200 import _lib as L
    Ln, N, K = 12, 256, 128
    W = [rnd((N, K), 100 + i) / 11 for i in range(Ln)]
    b = [rnd((N,), 200 + i) * 0.1 for i in range(Ln)]
    emb, dconds = rnd((B, K), 1), rnd((B, Ln, N), 2)
    Wr = [w.clone().requires_grad_(True) for w in W]
    br = [v.clone().requires_grad_(True) for v in b]
    er = emb.clone().requires_grad_(True)
    conds_ref = torch.stack([er @ Wr[i].T + br[i] for i in range(Ln)], 1)
    conds_ref.backward(dconds)
    Wd, bd_ = [w.cuda() for w in W], [v.cuda() for v in b]
    gW, gb = [torch.zeros_like(w) for w in Wd], [torch.zeros_like(v) for v in bd_]
    tab = eng._ptr_table(("t_aff", B), [t for i in range(Ln) for t in (Wd[i], bd_[i])])
    gtab = eng._ptr_table(("t_aff_g", B), [t for i in range(Ln) for t in (gW[i], gb[i])])
    embd, dcd = emb.cuda(), dconds.cuda()
    conds, part, demb = eng.empty(B, Ln, N), eng.empty(Ln, B, K), eng.empty(B, K)
    d = L.LinearBatchDesc()
    d.L, d.B, d.N, d.K = Ln, B, N, K
    d.params, d.grads = tab.data_ptr(), gtab.data_ptr()
    d.x, d.x_bstride = embd.data_ptr(), K
    d.out, d.y, d.y_bstride = conds.data_ptr(), dcd.data_ptr(), Ln * N
    for i in range(Ln):
        d.x_off[i], d.y_off[i] = 0, i * N
    d.part, d.dx = part.data_ptr(), demb.data_ptr()
    L.check(eng.lib.avc_linear_batch_fwd(C.byref(d), eng.stream), "fwd")
    L.check(eng.lib.avc_linear_batch_dx(C.byref(d), eng.stream), "dx")
    L.check(eng.lib.avc_linear_batch_dw(C.byref(d), eng.stream), "dw")
    torch.cuda.synchronize()
    assert relerr(conds, conds_ref) < TOL
    assert relerr(demb, er.grad) < TOL
    for i in range(Ln):
        assert relerr(gW[i], Wr[i].grad) < TOL and relerr(gb[i], br[i].grad) < TOL, i


def test_engine_fused_matches_per_layer(eng):
    """Whole speaker encoder + decoder affine path: fused_dense on vs off, same engine, same inputs."""
    cfg = orc.default_config(80)
    sd = {k: v.cuda() for k, v in orc.init_state(cfg, seed=0).items()}
    x = rnd((5, 80, 128), 3).cuda()
    res = {}
    eng.pack_weights(sd, need_dgrad=True)
    for fused in (False, True):
        eng.fused_dense = fused
        G = {k: torch.zeros_like(v) for k, v in sd.items()}
        emb, ctx = eng.speaker_fwd(sd, x, True)
        demb = rnd(tuple(emb.shape), 4).cuda()
        eng.speaker_bwd(sd, G, ctx, demb)
        torch.cuda.synchronize()
        res[fused] = (emb.clone(), {k: v.clone() for k, v in G.items() if k.startswith("speaker_encoder")})
    eng.fused_dense = False
    assert relerr(res[True][0], res[False][0]) < 1e-5
    for k in res[False][1]:
        assert relerr(res[True][1][k], res[False][1][k]) < 1e-4, k
