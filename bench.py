#!/usr/bin/env python
"""bench.py — RSSFormer 512x512 bf16 training images/sec (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W          # this repo's sm_100a path
  python bench.py --impl reference ...                   # the reference's CPU path (oracle port) on the host cores

A "step" = forward + loss + backward + gradient all-reduce + clip + SGD on one synthetic batch of
16 tiles per GPU (BASELINE config #2; weak scaling: cfg #3 is the same per-GPU batch on 8 GPUs).
`value` has the batch resident in HBM; `e2e` stages every step's batch from pinned host memory and reads
the loss back.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_IMG_TRAIN = 528.11e9        # algorithmic fwd+bwd FLOPs per 512x512 image
METRIC = "RSSFormer 512x512 bf16 training images/sec"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks/throttle reasons every 200 ms while the timed region runs"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = max([int(r[1]) for r in self.rows if r[1].isdigit()] or [0])
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons), samples=len(sm))


def cpu_reference_run(steps, warmup, batch, size):
    """The reference's CPU implementation of the step (oracle port of HRNetFusion.forward + loss + backward + clip + SGD,
    pinned against the reference in oracle/gen_golden.py), fp32 eager on all host cores."""
    import torch
    from oracle import rssformer_ref as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = R.synth_state_dict(2333)
    keys = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    params = {k: sd[k].clone().requires_grad_(True) for k in keys}
    mom = [None] * len(keys)
    img, lbl = R.synth_batch(batch, size)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        cur = dict(sd); cur.update(params)
        out, stats = R.model_forward(cur, img, lbl, training=True)
        loss = out["fc_loss"]
        grads = torch.autograd.grad(loss, [params[k] for k in keys], allow_unused=True)
        R.sgd_step([params[k] for k in keys], list(grads), mom, R.poly_lr(it))
        sd.update(stats)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    tot = sum(times)
    return dict(value=batch * len(times) / tot, ms_per_step=1e3 * tot / len(times), cores=cores,
                sample="%d step(s) of %d tile(s) %dx%d fp32 after %d warm-up" % (len(times), batch, size, size, warmup))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="tiles per GPU (BASELINE cfg2/cfg3: 16)")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--cpu-baseline-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--profile", action="store_true", help="profiling run (under ncu): skip the e2e and cpu legs; numbers are not bench values")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "cfg2: RSSFormer(hrnetv2_w32) train step, %d tiles/GPU of %dx%d, synthetic LoveDA-shape" % (args.batch, args.size, args.size)

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(max(1, min(args.steps, 3)), 1 if args.warmup > 0 else 0, 1, args.size)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "sample": r["sample"]},
            "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    if os.environ.get("RSS_FAULTHANDLER"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["RSS_FAULTHANDLER"]), exit=True)
    # keep stdout clean for the ONE JSON line (NCCL/torch may print banners): everything else goes to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import representationlearning_b200 as P
    from representationlearning_b200 import ops
    from oracle import rssformer_ref as R      # cpu_baseline leg + deterministic synthetic weights/batches only

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P._lib.require_device()

    model = P.build_rssformer(compute_dtype=torch.bfloat16, device=dev)
    model.load_state_dict(R.synth_state_dict(2333))
    model.train()
    if os.environ.get("RSS_NO_SYNCBN"):
        for mod in model.modules():
            if isinstance(mod, P.FusedBNAct):
                mod.sync = False
    opt = P.FlatSGD(model)
    B, S = args.batch, args.size
    img_h, lbl_h = R.synth_batch(B, S, seed_img=7 + rank, seed_lbl=1 + rank)
    img_h, lbl_h = img_h.pin_memory(), lbl_h.pin_memory()
    img_d, lbl_d = img_h.to(dev), lbl_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    last = {}

    def step_eager():
        last["loss"] = P.train_step(model, opt, img_d, lbl_d)

    # ---- eager warm-up with live CUDA-event timing of the hand-written regions (same shapes, same process) ----
    for _ in range(2):
        step_eager()
    dom = "rss_conv_igemm"
    ops.TIMED_OPS.update(["rss_attn_bwd", "rss_attn_fwd", "rss_conv_igemm"]); ops.TIMED.clear()
    c0 = ops.COUNTERS["launches"]
    ms_eager = timed(step_eager, 2) / 2
    launches_per_step = (ops.COUNTERS["launches"] - c0) // 2
    ops.TIMED_OPS.clear()
    torch.cuda.synchronize()
    kt = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in ops.TIMED.items() if v}
    kn = {k: len(v) for k, v in ops.TIMED.items() if v}

    # ---- the timed region: the whole step captured once as a CUDA graph, replayed K times -------------------
    if world > 1 and os.environ.get("RSS_GRAPH_DDP", "1") == "0":
        args.no_graph = True
    if args.no_graph:
        run_step = step_eager
    else:
        graphed = P.GraphedTrainStep(model, opt, img_d, lbl_d, warmup=max(1, args.warmup - 2))

        def run_step():
            last["loss"] = graphed()
    for _ in range(max(args.warmup, 3)):
        run_step()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(run_step, args.steps)
    clocks = sampler.finish() if sampler else None
    launches = launches_per_step * args.steps
    value = world * B * args.steps / (ms / 1e3)
    if args.profile:
        os.write(json_fd, (json.dumps({"profile_run": True, "ms_per_step": ms / args.steps, "region_ms": kt, "gpu_launches": launches}) + "\n").encode())
        return 0

    # ---- end to end: pinned host -> device every step (staged on a copy stream, overlapped), loss read back ---
    copy_stream = torch.cuda.Stream(dev)
    bufs = [(torch.empty_like(img_d), torch.empty_like(lbl_d)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "sink": 0.0}

    def stage(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            bufs[slot][0].copy_(img_h, non_blocking=True)
            bufs[slot][1].copy_(lbl_h, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e():
        slot = state["i"] & 1
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[slot])
        stage(slot ^ 1)                                   # prefetch the next step's batch while this one computes
        if args.no_graph:
            loss = P.train_step(model, opt, bufs[slot][0], bufs[slot][1])
        else:
            graphed.load(bufs[slot][0], bufs[slot][1])    # device-to-device into the graph's static inputs
            loss = graphed()
        consumed[slot].record(cur)
        state["sink"] += float(loss.item())              # device -> host read of the step's result
        state["i"] += 1

    for e in consumed:
        e.record(torch.cuda.current_stream())
    stage(0)
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    h2d = img_h.numel() * img_h.element_size() + lbl_h.numel() * lbl_h.element_size()

    pk = peaks()
    per_gpu = value / world
    # dominant hand-written kernel: the tcgen05/TMA implicit-GEMM convolution running the FFN's dw+dw6+dw12 convs as one
    # GEMM (forward and data gradient: 16 launches per step).  ALGORITHMIC FLOPs per launch = what the reference's three
    # convolutions execute: 2 * (1 + 9 + 9 taps) * 128 * 128 * (B*128*128 pixels) (FFN dil-6 + dil-12 + 1x1);
    # the kernel itself runs 17 taps (the three centre taps are merged).  The live timing below brackets the whole C-ABI call
    # (tensor-map encode + launch + kernel) with CUDA events in an eager pass, so it is an upper bound of the kernel time.
    hw4 = (S // 4) * (S // 4)
    alg_flops = 2.0 * 19 * 128 * 128 * B * hw4
    tok_bytes = B * hw4 * 32 * 2
    roof = None
    if dom in kt:
        ach = alg_flops / (kt[dom] / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "conv_igemm_kernel (FFN 19-tap conv, fwd/dgrad)", "achieved": ach, "peak": pk["tf_burst"],
                "unit": "TFLOP/s", "frac": ach / pk["tf_burst"], "traffic": 88.7e6,
                "peak_source": pk["src"] + " (burst cuBLAS bf16; kernel timed alone per launch)", "ms_per_launch": kt[dom],
                "launches_timed": kn[dom], "algorithmic_flops_per_launch": alg_flops,
                "traffic_source": "profiles/ncu_full_igemm_mm2_r1.csv: dram read 67.7 MB + write 21.0 MB per launch (algorithmic: 67.1 MB in + 67.1 MB out; "
                                  "the output stays in the 126 MB L2 for the consumer); the same capture times the kernel alone at 121.7 us = 0.82 of peak",
                "timing": "CUDA events around the C-ABI call on the launching stream, eager pass of the same step in this process"}
    hbm_regions = {}
    for name, nbytes in (("rss_attn_fwd", 3 * tok_bytes), ("rss_attn_bwd", 5 * tok_bytes)):
        if name in kt:
            a = nbytes / (kt[name] / 1e3) / 1e9
            hbm_regions[name] = {"bound": "hbm", "achieved_gbs": a, "frac": a / pk["hbm"], "algorithmic_bytes": nbytes, "ms": kt[name]}
    step_tf = per_gpu * FLOP_PER_IMG_TRAIN / 1e12
    out = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": workload, "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2": "per-step activations (GBs) far exceed the 126 MB L2; no explicit flush",
                   "launch": "eager" if args.no_graph else "whole step replayed as one CUDA graph", "ms_per_step_eager": ms_eager,
                   "schedule": "data-flow streams per HRNet resolution + 2 weight-gradient side streams (parallel sub-graphs)",
                   "weights": "synthetic, seed 2333 (oracle.synth_state_dict)"},
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": roof,
        "step_roofline": {"bound": "tensor", "achieved": step_tf, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": step_tf / pk["tf_sust"],
                          "note": "whole step vs sustained bf16 GEMM peak, 528.11 GFLOP/img algorithmic"},
        "region_ms": kt, "hbm_regions": hbm_regions, "clocks": clocks, "loss": float(last["loss"].item()),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.cpu_baseline_steps, 1, 1, S)
        out["cpu_baseline"] = {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    if rank == 0:
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        # destroy_process_group() can block while captured graphs still reference the communicator: synchronise and leave
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
