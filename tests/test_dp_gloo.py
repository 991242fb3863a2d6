"""CPU, world_size=2, gloo: the data-parallel recipe of trainer.FusedTrainer
restated with the oracle -- per-rank gradients on a shard, ONE all-reduce(SUM) of the flat
gradient buffer, 1/world folded into the clip coefficient, then clip + Adam on every rank --
equals the single-process step on the concatenated batch, and replicas stay identical."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle.ae_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flat(d, names):
    return torch.cat([d[k].reshape(-1) for k in names])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = orc.default_config(16)
    for k in ("SpeakerEncoder", "ContentEncoder"):
        cfg[k].update(c_h=32, c_out=32, c_bank=16, bank_size=4, n_conv_blocks=2, subsample=[1, 2])
    cfg["SpeakerEncoder"]["n_dense_blocks"] = 1
    cfg["Decoder"].update(c_in=32, c_cond=32, c_h=32, n_conv_blocks=2, upsample=[2, 1])
    sd = orc.init_state(cfg, seed=0)
    names = list(sd)
    x = torch.randn((4, 16, 32), generator=torch.Generator().manual_seed(1))
    eps = torch.randn((4, 32, 16), generator=torch.Generator().manual_seed(2))
    sl = slice(rank * 2, rank * 2 + 2)
    _, g = orc.ae_loss_and_grads(sd, cfg, x[sl], eps[sl], 0.5)
    flat = _flat(g, names)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)               # the single exchange step
    flat = flat / world                                         # grad_scale = 1/world (hp[2])
    avg, off = {}, 0
    for k in names:
        n = sd[k].numel()
        avg[k] = flat[off:off + n].view(sd[k].shape)
        off += n
    st = orc.AdamState(sd)
    gn = orc.clip_and_adam(sd, avg, st, cfg["optimizer"])
    mine = _flat(sd, names)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        sd1 = orc.init_state(cfg, seed=0)
        _, g1 = orc.ae_loss_and_grads(sd1, cfg, x, eps, 0.5)
        gn1 = orc.clip_and_adam(sd1, g1, orc.AdamState(sd1), cfg["optimizer"])
        gerr = float((_flat(avg, names) - _flat(g1, names)).norm() / _flat(g1, names).norm())
        sel = _flat(g1, names).abs() > 1e-5   # Adam's first step is lr*sign(g): skip noise-gradient elements
        out.put({"replicas_equal": bool(torch.equal(gathered[0], gathered[1])),
                 "grad_err": gerr, "gn": (gn, gn1),
                 "param_err": float(((mine - _flat(sd1, names)).abs() * sel).max())})
    dist.barrier()
    dist.destroy_process_group()


def test_dp_equals_single_process_on_concatenated_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["replicas_equal"]
    assert res["grad_err"] < 1e-5, res
    assert abs(res["gn"][0] - res["gn"][1]) / res["gn"][1] < 1e-5
    assert res["param_err"] < 1e-5, res
