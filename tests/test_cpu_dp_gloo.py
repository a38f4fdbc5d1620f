"""N>1 path on CPU (`gloo`, world_size 2): the data-parallel protocol the CUDA path implements — SyncBN statistics by
exchanging (mean, M2, count) + Chan combine, gradient mean all-reduce on a flat buffer — checked against the single-process
oracle on the full batch; plus bench.py's reference arm under a 2-rank launch (rank 0 alone reports)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from oracle import rssformer_ref as R

BLK = "backbone.hrnet.stage2.0.transformer."


def _chan_combine(stats):
    """stats: (world, C, 3) rows of (n, mean, M2) -> (n, mean, M2); same recurrence as csrc/bn.cu bn_warp_combine"""
    n = torch.zeros_like(stats[0, :, 0]); m = torch.zeros_like(n); q = torch.zeros_like(n)
    for r in range(stats.shape[0]):
        nb, mb, qb = stats[r, :, 0], stats[r, :, 1], stats[r, :, 2]
        nn = n + nb
        d = mb - m
        m = m + d * nb / nn
        q = q + qb + d * d * n * nb / nn
        n = nn
    return n, m, q


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    sd = {k: v.double() for k, v in R.synth_state_dict(2333).items() if k.startswith(BLK)}
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 32, 14, 14, generator=g, dtype=torch.float64)
    y = torch.randn(4, 32, 14, 14, generator=g, dtype=torch.float64)
    dout = torch.randn(4, 32, 14, 14, generator=g, dtype=torch.float64)
    sl = slice(2 * rank, 2 * rank + 2)

    orig_bn = R._bn

    def sync_bn(ctx, t, prefix, eps=R.BN_EPS, momentum=R.BN_MOMENTUM):       # SyncBN forward as the CUDA path does it
        w, b = ctx[prefix + ".weight"], ctx[prefix + ".bias"]
        nloc = t.numel() // t.shape[1]
        mean_l = t.mean((0, 2, 3))
        m2_l = ((t - mean_l[None, :, None, None]) ** 2).sum((0, 2, 3))
        local = torch.stack([torch.full_like(mean_l, float(nloc)), mean_l.detach(), m2_l.detach()], 1)
        allst = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(allst, local)
        n, mean, m2 = _chan_combine(torch.stack(allst))
        # autograd through the global statistics (what all_reduce of (sum dz, sum dz*xhat) implements in backward)
        import torch.distributed.nn.functional as dfn
        s1 = dfn.all_reduce(t.sum((0, 2, 3)))
        s2 = dfn.all_reduce((t * t).sum((0, 2, 3)))
        gmean = s1 / n
        gvar = s2 / n - gmean * gmean
        assert torch.allclose(gmean.detach(), mean, atol=1e-10) and torch.allclose(gvar.detach(), m2 / n, atol=1e-9)
        xh = (t - gmean[None, :, None, None]) * torch.rsqrt(gvar + eps)[None, :, None, None]
        return xh * w[None, :, None, None] + b[None, :, None, None]

    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    R._bn = sync_bn
    try:
        o = R.transformer_block(R.Ctx(sdg, True), BLK, x[sl], y[sl])
        (o * dout[sl]).sum().backward()
    finally:
        R._bn = orig_bn
    keys = [k for k in sdg if sdg[k].requires_grad and sdg[k].grad is not None]
    flat = torch.cat([sdg[k].grad.reshape(-1) for k in keys])          # flat-buffer all-reduce, mean folded in afterwards
    dist.all_reduce(flat)
    flat /= world
    if rank == 0:
        ref = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
        of = R.transformer_block(R.Ctx(ref, True), BLK, x, y)
        ((of * dout).sum() / world).backward()
        rflat = torch.cat([ref[k].grad.reshape(-1) for k in keys])
        out.put(((flat - rflat).abs().max() / rflat.abs().max()).item())
    dist.destroy_process_group()


def test_dp_protocol_world2_gloo_matches_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert err < 1e-9, err


def test_bench_reference_arm_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""          # non-zero ranks exit 0 without work
