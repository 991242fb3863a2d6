"""GPU diagnostic: what the TF32 tensor-core path does to the gradient, with data.
  (1) gradient relative L2 error vs the fp32 CPU oracle for: tf32 (default), fp32 forward + tf32 backward
      (AVC_FWD_FP32=1), exact fp32 -- if the middle one is as good as fp32, the tf32 number is ReLU-mask flips of
      the TF32 forward, not an inaccurate backward.
  (2) 200 optimizer steps on one fixed batch: loss trajectories of the tf32 path, the fp32 path and the oracle."""
import os, sys, types, contextlib, io, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle.ae_oracle as orc

B = int(os.environ.get("DIAG_B", "16"))
STEPS = int(os.environ.get("DIAG_STEPS", "200"))
cfg = orc.default_config(80)
cfg["data_loader"]["batch_size"] = B
torch.set_num_threads(16)


def make_solver(precision, fwd_fp32=False):
    os.environ["AVC_PRECISION"] = precision
    os.environ["AVC_FWD_FP32"] = "1" if fwd_fp32 else "0"
    from adaptive_voice_conversion_b200.solver import Solver
    args = types.SimpleNamespace(data_dir="synthetic", train_set="", train_index_file="", logdir="/tmp/avc_log", load_model=False,
                                 load_opt=False, store_model_path=None, load_model_path=None, summary_steps=10 ** 9, save_steps=10 ** 9, tag="d", iters=0)
    with contextlib.redirect_stdout(io.StringIO()):
        s = Solver(cfg, args)
    s.model.load_state_dict(orc.init_state(cfg, seed=0), strict=True)
    s.trainer.eng.pack_weights(s.trainer.P, need_dgrad=True)
    return s


x = torch.randn((B, 80, 128), generator=torch.Generator().manual_seed(1))
eps = torch.randn((B, 128, 16), generator=torch.Generator().manual_seed(2))
sd = orc.init_state(cfg, seed=0)
outs, gref = orc.ae_loss_and_grads(sd, cfg, x, eps, 1.0)
names = list(sd)
ref = torch.cat([gref[k].double().flatten() for k in names])
res = {}
for tag, prec, ff in (("tf32", "tf32", False), ("fp32fwd_tf32bwd", "tf32", True), ("fp32", "fp32", False)):
    s = make_solver(prec, ff)
    tr = s.trainer
    tr.opt.sync_hparams(lambda_rec=10.0, lambda_kl=1.0)
    tr._fwd_bwd(x.cuda(), eps.cuda())
    torch.cuda.synchronize()
    g = torch.cat([tr.G[k].detach().double().cpu().flatten() for k in names])
    res[tag] = float((g - ref).norm() / ref.norm())
    print(f"gradient rel-L2 vs fp32 oracle, B={B}: {tag:18s} {res[tag]:.3e}", flush=True)
    del s, tr
    torch.cuda.empty_cache()

# ---- trajectories on a fixed batch (eps re-drawn per step from a seeded generator, identical for all three)
def eps_of(i):
    return torch.randn((B, 128, 16), generator=torch.Generator().manual_seed(1000 + i))

traj = {}
for tag, prec in (("tf32", "tf32"), ("fp32", "fp32")):
    s = make_solver(prec)
    L = []
    for i in range(STEPS):
        s.trainer.step(x.cuda(), 1.0, eps=eps_of(i).cuda())
        L.append(s.trainer.losses())
    traj[tag] = L
    del s
    torch.cuda.empty_cache()
sd2 = orc.init_state(cfg, seed=0)
st = orc.AdamState(sd2)
L = []
for i in range(STEPS):
    r = orc.ae_train_step(sd2, st, cfg, x, eps_of(i), 1.0)
    L.append((r["loss_rec"], r["loss_kl"], r["grad_norm"]))
traj["oracle"] = L
for i in (0, 1, 9, 49, 99, STEPS - 1):
    if i < STEPS:
        print(f"step {i + 1:4d}  loss_rec tf32 {traj['tf32'][i][0]:.5f}  fp32 {traj['fp32'][i][0]:.5f}  oracle {traj['oracle'][i][0]:.5f}   "
              f"loss_kl tf32 {traj['tf32'][i][1]:.5f}  fp32 {traj['fp32'][i][1]:.5f}  oracle {traj['oracle'][i][1]:.5f}", flush=True)
json.dump({"B": B, "grad_rel_l2": res, "traj": traj}, open("gpurun_out/diag_tf32.json", "w"))
