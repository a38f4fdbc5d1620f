#!/usr/bin/env python
"""bench.py — RSSFormer 512x512 bf16 training images/sec (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W          # this repo's sm_100a path
  python bench.py --impl reference ...                   # the UNMODIFIED reference (oracle/_ref archive) on the host cores

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


def reference_step_fn(model, img, lbl, device_type, autocast_dtype=None):
    """one training step of the UNMODIFIED reference model, as BASELINE.md section 6 defines it (configs/base/loveda.py:68-93):
    loss = sum(model(x, y).values()); backward; clip_grad_norm_(35, 2); SGD(lr poly(0.01), momentum .9, wd 1e-4); zero_grad"""
    import torch
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=0.01, momentum=0.9, weight_decay=1e-4)
    state = {"it": 0}

    def step():
        for g in opt.param_groups:
            g["lr"] = 0.01 * (1.0 - min(state["it"], 30000) / 30000) ** 0.9
        with torch.autocast(device_type, dtype=autocast_dtype or torch.bfloat16, enabled=autocast_dtype is not None):
            loss = sum(model(img, {"cls": lbl}).values())
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 35.0, 2)
        opt.step()
        opt.zero_grad(set_to_none=True)
        state["it"] += 1
        return loss.detach()
    return step


def cpu_reference_run(steps, warmup, batch, size, budget_s=None):
    """The reference's own CPU implementation of the step: module.baseline.hrnet_aux.HRNetFusion imported UNMODIFIED from
    /root/reference or the travelling archive oracle/_ref/rssformer_reference.zip (oracle/ref_shim.py: stand-ins for the absent `ever` /
    `timm` packages only, no arithmetic), fp32 eager on all host cores.  Falls back to the oracle port (kind "port") when neither
    is present.  budget_s: stop early (and say so) when the timed steps would exceed it."""
    import torch
    from oracle import rssformer_ref as R
    from oracle import ref_shim
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    img, lbl = R.synth_batch(batch, size)
    kind = "reference" if ref_shim.reference_available() else "port"
    if kind == "reference":
        model = ref_shim.build_reference_model()
        model.load_state_dict(R.synth_state_dict(2333))
        model.train()
        step = reference_step_fn(model, img, lbl, "cpu")
    else:
        sd = R.synth_state_dict(2333)
        keys = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k]
        params = {k: sd[k].clone().requires_grad_(True) for k in keys}
        mom = [None] * len(keys)
        it = {"i": 0}

        def step():
            cur = dict(sd); cur.update(params)
            out, stats = R.model_forward(cur, img, lbl, training=True)
            grads = torch.autograd.grad(out["fc_loss"], [params[k] for k in keys], allow_unused=True)
            R.sgd_step([params[k] for k in keys], list(grads), mom, R.poly_lr(it["i"]))
            sd.update(stats)
            it["i"] += 1
            return out["fc_loss"].detach()
    times, t_start = [], time.perf_counter()
    first_loss = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = step()
        dt = time.perf_counter() - t0
        if first_loss is None:
            first_loss = float(loss)
        if i >= warmup:
            times.append(dt)
        if budget_s is not None and i >= warmup and time.perf_counter() - t_start + dt > budget_s:
            break
    tot = sum(times)
    return dict(value=batch * len(times) / tot, ms_per_step=1e3 * tot / len(times), cores=cores, kind=kind, steps=len(times), warmup=warmup,
                first_loss=first_loss,
                sample="%d step(s) of %d tile(s) %dx%d fp32 after %d warm-up, %s, torch %d threads"
                       % (len(times), batch, size, size, warmup,
                          "unmodified reference (HRNetFusion from the reference sources)" if kind == "reference" else "oracle port", cores))


def gpu_eager_baseline(dev, batch, size, steps=5, warmup=2):
    """The UNMODIFIED reference on the same B200 in eager PyTorch, bf16 autocast: once as shipped
    (train.py:73 sets torch.backends.cudnn.enabled = False) and once with cuDNN enabled -- the like-for-like GPU baseline."""
    import torch
    from oracle import rssformer_ref as R
    from oracle import ref_shim
    if not ref_shim.reference_available():
        return {"unavailable": "reference sources not present (oracle/_ref/rssformer_reference.zip)"}
    img, lbl = R.synth_batch(batch, size)
    img, lbl = img.to(dev), lbl.to(dev)
    out = {"unit": "images/s", "dtype": "bf16 autocast", "batch": batch, "steps": steps, "warmup": warmup,
           "what": "RSSFormer-TIP2023 HRNetFusion, unmodified, eager torch %s on this GPU; clip 35 + torch.optim.SGD" % torch.__version__}
    saved = torch.backends.cudnn.enabled
    for name, cudnn_on in (("cudnn_enabled", True), ("as_shipped_cudnn_disabled", False)):
        try:
            torch.backends.cudnn.enabled = cudnn_on
            model = ref_shim.build_reference_model()
            model.load_state_dict(R.synth_state_dict(2333))
            model = model.to(dev).train()
            step = reference_step_fn(model, img, lbl, "cuda", torch.bfloat16)
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": batch / (ms / 1e3), "ms_per_step": ms, "loss": float(loss.item())}
            del model, step
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001  (an OOM or an unsupported op in the reference must not kill the bench line)
            out[name] = {"error": repr(e)[:300]}
    torch.backends.cudnn.enabled = saved
    return out


# kernel-name substring -> family, for the CUPTI summary of the replayed step
FAMILIES = (
    ("stem_conv", "stem conv (stem_conv_{fwd,wgrad}_kernel: 3->64 stride-2 conv from the planar image, cast + layout + BN sums fused)"),
    ("bn_", "batchnorm (rss bn_* kernels: statistics / apply / backward reduce / backward apply)"),
    ("conv_cf_kernel", "fused tcgen05 conv (conv_cf_kernel: HRNet branch-0 BasicBlocks)"),
    ("conv_igemm_kernel", "tcgen05 implicit GEMM (conv_igemm_kernel: FFN 19-tap conv)"),
    ("win_attn_fwd", "window attention forward"), ("win_attn_bwd", "window attention backward"), ("gate_", "saliency gate"),
    ("ln_", "layernorm"), ("conv_wgrad", "own weight-gradient kernels"), ("fuse_sum", "multi-resolution fuse"),
    ("neck_gather", "neck gather"), ("head_", "head"), ("seg_loss", "loss"), ("sgd_step", "optimiser"), ("sumsq", "optimiser"), ("accum_list", "optimiser"),
    ("shadow_", "optimiser"), ("cutlass", "library conv (cuDNN cutlass3x / xmma)"), ("xmma", "library conv (cuDNN cutlass3x / xmma)"),
    ("cudnn", "library conv (cuDNN cutlass3x / xmma)"), ("nvjet", "library GEMM (cuBLAS nvjet: 1x1 convs)"),
    ("splitKreduce", "library GEMM (cuBLAS nvjet: 1x1 convs)"), ("elementwise", "torch elementwise (gradient accumulation adds)"),
    ("nccl", "nccl"), ("Memset", "memset"), ("Memcpy", "memcpy"),
)


def cupti_step_summary(run_step, n=2, profile_here=True):
    """per-kernel durations of the REPLAYED step (CUPTI through torch.profiler): {name: (total_us, count)} averaged over n replays.
    Durations under the profiler are used for shares and per-launch kernel times only, never for the bench value.
    EVERY rank must call this (the step contains collectives); only the rank with profile_here=True records."""
    import contextlib
    import torch
    from torch.profiler import ProfilerActivity, profile
    run_step()
    torch.cuda.synchronize()
    ctx = profile(activities=[ProfilerActivity.CUDA]) if profile_here else contextlib.nullcontext()
    with ctx as prof:
        for _ in range(n):
            run_step()
        torch.cuda.synchronize()
    if not profile_here:
        return None
    per = {}
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0.0)
        if t and e.count:
            per[e.key] = (t / n, e.count / n)
    return per


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="tiles per GPU (BASELINE cfg2/cfg3: 16)")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--cpu-baseline-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the eager-PyTorch run of the unmodified reference on this GPU")
    ap.add_argument("--no-bilateral", action="store_true", help="skip the bilateral-filter side measurement")
    ap.add_argument("--no-cupti", action="store_true", help="skip the CUPTI kernel summary of the replayed step (family rooflines)")
    ap.add_argument("--no-graph", action="store_true", help="issue the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--profile", action="store_true", help="profiling run (under ncu): skip the e2e and cpu legs; numbers are not bench values")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    global METRIC, FLOP_PER_IMG_TRAIN
    cfg = "cfg2" if args.size == 512 else ("cfg5" if args.size == 1024 else "custom")
    workload = "%s: RSSFormer(hrnetv2_w32) train step, %d tiles/GPU of %dx%d, synthetic LoveDA-shape" % (cfg, args.batch, args.size, args.size)
    if args.size != 512:        # BASELINE cfg5 (1024x1024 large tiles, 4 per GPU): same step, its own metric name and FLOP count
        METRIC = "RSSFormer %dx%d bf16 training images/sec" % (args.size, args.size)
        FLOP_PER_IMG_TRAIN = 2111.09e9 if args.size == 1024 else FLOP_PER_IMG_TRAIN * (args.size / 512.0) ** 2

    if args.impl == "reference":
        # the reference's own CPU implementation on the host cores: each step = a bounded SAMPLE of the workload (4 of the 16 tiles,
        # BASELINE.md section 6), exactly --warmup + --steps of them unless the time budget cuts the run short (then the line says
        # how many were timed).  Rank 0 only.
        if rank != 0:
            return 0
        sample_b = int(os.environ.get("RSS_REF_SAMPLE_TILES", "4"))
        r = cpu_reference_run(args.steps, args.warmup, sample_b, args.size, budget_s=float(os.environ.get("RSS_REF_BUDGET_S", "420")))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": r["warmup"], "steps_requested": args.steps, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "global_batch": args.batch * world, "parallelism": "cpu x%d threads" % r["cores"],
                       "sample": r["sample"], "weights": "synthetic, seed 2333 (oracle.synth_state_dict)"},
            "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "loss_first_step": r["first_loss"]}))
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
    step_eager()
    loss0 = float(last["loss"].item())
    # the step being benched must compute the reference's loss: step-0 loss vs the UNMODIFIED reference's fp32 loss on exactly this
    # rank's batch (oracle/gen_bench_loss.py -> tests/golden/bench_cfg2_loss.json); 2x the reference's own bf16-autocast envelope
    loss_check = None
    gl = os.path.join(ROOT, "tests", "golden", "bench_cfg2_loss.json")
    if os.path.exists(gl) and B == 16 and S == 512 and not os.environ.get("RSS_NO_SYNCBN"):
        ref_l = json.load(open(gl))["loss_fp32_step0"].get(str(rank))
        if ref_l is not None and world == 1:
            rel_err = abs(loss0 - ref_l) / abs(ref_l)
            loss_check = {"loss_step0": loss0, "reference_fp32": ref_l, "rel_err": rel_err, "tolerance": 1e-3}
            if not rel_err < 1e-3:
                raise SystemExit("bench.py: step-0 loss %.6f differs from the reference's %.6f (rel %.2e): not benching a wrong step"
                                 % (loss0, ref_l, rel_err))
    step_eager()
    dom = "rss_conv_igemm"
    ops.TIMED_OPS.update(["rss_attn_bwd", "rss_attn_fwd", "rss_conv_igemm"]); ops.TIMED.clear()
    c0 = ops.COUNTERS["launches"]
    ops.ACCOUNT.clear(); ops.ACCOUNT_ON[0] = True          # algorithmic bytes per kernel family of ONE step (for roofline_families)
    ms_eager = timed(step_eager, 2) / 2
    ops.ACCOUNT_ON[0] = False
    for k in list(ops.ACCOUNT):
        ops.ACCOUNT[k] /= 2
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

    fam = None
    if not args.no_cupti:
        try:
            fam = cupti_step_summary(run_step, profile_here=(rank == 0))      # all ranks replay: the step contains collectives
        except Exception as e:  # noqa: BLE001
            fam = None
            sys.stderr.write("CUPTI summary failed: %r\n" % (e,))

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
    igemm_ms, igemm_src = kt.get(dom), "CUDA events around the C-ABI call on the launching stream, eager pass of the same step in this process"
    if fam:
        k = [v for n, v in fam.items() if "conv_igemm_kernel" in n]
        if k and k[0][1] > 0:
            igemm_ms = k[0][0] / k[0][1] / 1e3
            kn[dom] = int(round(k[0][1]))
            igemm_src = "CUPTI kernel durations inside the replayed CUDA graph (torch.profiler), mean over the step's launches"
    if igemm_ms:
        kt[dom] = igemm_ms
        ach = alg_flops / (kt[dom] / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "conv_igemm_kernel (FFN 19-tap conv, fwd/dgrad)", "achieved": ach, "peak": pk["tf_burst"],
                "unit": "TFLOP/s", "frac": ach / pk["tf_burst"], "traffic": 88.7e6,
                "peak_source": pk["src"] + " (burst cuBLAS bf16; kernel timed alone per launch)", "ms_per_launch": kt[dom],
                "launches_timed": kn[dom], "algorithmic_flops_per_launch": alg_flops,
                "traffic_source": "profiles/ncu_full_igemm_mm2_r1.csv: dram read 67.7 MB + write 21.0 MB per launch (algorithmic: 67.1 MB in + 67.1 MB out; "
                                  "the output stays in the 126 MB L2 for the consumer); the same capture times the kernel alone at 121.7 us = 0.82 of peak",
                "timing": igemm_src}
    hbm_regions = {}
    for name, nbytes in (("rss_attn_fwd", 3 * tok_bytes), ("rss_attn_bwd", 5 * tok_bytes)):
        if name in kt:
            a = nbytes / (kt[name] / 1e3) / 1e9
            hbm_regions[name] = {"bound": "hbm", "achieved_gbs": a, "frac": a / pk["hbm"], "algorithmic_bytes": nbytes, "ms": kt[name]}
    families = None
    if fam:
        tot = sum(v[0] for v in fam.values())
        agg = {}
        for n, (t, c) in fam.items():
            f = next((lab for key, lab in FAMILIES if key in n), "other")
            a = agg.setdefault(f, [0.0, 0.0])
            a[0] += t; a[1] += c
        acct = dict(ops.ACCOUNT)
        families = {"kernel_time_sum_ms": tot / 1e3, "step_ms": ms / args.steps, "average_concurrency": tot / 1e3 / (ms / args.steps),
                    "source": "CUPTI (torch.profiler) over 2 replays of the captured step; shares of the summed kernel time", "by_family": {}}
        for f, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            ent = {"ms": t / 1e3, "share": t / tot, "launches": int(round(c))}
            key = "bn" if f.startswith("batchnorm") else ("cf" if f.startswith("fused tcgen05") else ("attn_fwd" if f == "window attention forward"
                  else ("attn_bwd" if f == "window attention backward" else ("stem" if f.startswith("stem conv") else None))))
            if key and acct.get(key):
                gbs = acct[key] / (t / 1e6) / 1e9
                ent.update({"bound": "hbm", "algorithmic_bytes": acct[key], "achieved_gbs": gbs, "peak_gbs": pk["hbm"], "frac": gbs / pk["hbm"]})
            families["by_family"][f] = ent
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
        "roofline_families": families,
        "region_ms": kt, "hbm_regions": hbm_regions, "clocks": clocks, "loss": float(last["loss"].item()), "loss_check": loss_check,
    }
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        if not args.no_graph:
            del graphed
        torch.cuda.empty_cache()
        out["gpu_eager_baseline"] = gpu_eager_baseline(dev, B, S)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.cpu_baseline_steps, 1, 4, S, budget_s=60.0)
        out["cpu_baseline"] = {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    if rank == 0 and world == 1 and not args.no_bilateral:
        # second C-ABI path of the repo: the SCD DenseEnergyLoss bilateral filter, device vs the compiled reference on
        # the host threads it can use; reported beside the headline, not part of it
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bilateral_bench as BB
            out["bilateral_filter"] = [BB.run_workload(*w, iters=20, hbm=pk["hbm"]) for w in BB.WORKLOADS[:2]]
        except Exception as e:  # noqa: BLE001
            out["bilateral_filter"] = {"error": repr(e)}
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
