"""Times the device bilateral filter (csrc/bilateral.cu) next to the compiled reference / the C restatement on the host cores.
Prints one JSON object per workload; numbers under a profiler are not bench values.
    python tools/bilateral_bench.py [--iters 50] [--profile]     (--profile: 3 plain calls per workload, for ncu)"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bilateral as B                      # noqa: E402  (seeded inputs + the CPU baseline leg)
from representationlearning_b200 import scd, _lib      # noqa: E402

WORKLOADS = [("scd_voc: N=2 K=21 160x160 (crop 320, scale 0.5; configs/voc_attn_reg.yaml:10,23)", 2, 21, 160, 160),
             ("cfg4: N=8 K=21 224x224 (448 crop, scale 0.5)", 8, 21, 224, 224),
             ("large: N=16 K=21 512x512", 16, 21, 512, 512)]


def run_workload(name, N, K, H, W, iters=30, hbm=6535.1, cpu_reps=3):
    """one workload: device-resident (eager + graph replay), end to end through the reference's argument list, CPU reference"""
    img, seg = B.synth(N, K, H, W, seed=41)
    d_img, d_seg = torch.from_numpy(img).cuda(), torch.from_numpy(seg).cuda()
    out = torch.empty_like(d_seg)
    m = torch.zeros(1, dtype=torch.int32, device="cuda")
    for _ in range(3):
        scd.bilateral_filter(d_img, d_seg, 15.0, 50.0, out=out, lattice_points=m)
    torch.cuda.synchronize()
    # eager launches (21 kernels + 2 memsets per call)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        scd.bilateral_filter(d_img, d_seg, 15.0, 50.0, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms_eager = e0.elapsed_time(e1) / iters
    # the same call captured once and replayed
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        scd.bilateral_filter(d_img, d_seg, 15.0, 50.0, out=out)
        with torch.cuda.graph(g, stream=s):
            scd.bilateral_filter(d_img, d_seg, 15.0, 50.0, out=out)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms_graph = e0.elapsed_time(e1) / iters
    # end to end through the reference's own argument list: host arrays in, host array out
    images, ins = img.flatten(), seg.flatten()
    AS = np.zeros_like(ins)
    for _ in range(2):
        scd.bilateralfilter_batch(images, ins, AS, N, K, H, W, 15.0, 50.0)
    t = time.perf_counter()
    reps = max(3, iters // 5)
    for _ in range(reps):
        scd.bilateralfilter_batch(images, ins, AS, N, K, H, W, 15.0, 50.0)
    ms_e2e = (time.perf_counter() - t) / reps * 1e3
    # CPU: the reference itself (OpenMP over images, as shipped) when oracle/_ref holds it, else the C restatement
    fn, kind = (B.reference_filter, "reference") if B.have_reference() else (B.oracle_filter, "port")
    ref = fn(img, seg, 15.0, 50.0)
    t = time.perf_counter()
    for _ in range(cpu_reps):
        fn(img, seg, 15.0, 50.0)
    ms_cpu = (time.perf_counter() - t) / cpu_reps * 1e3
    same = bool(np.array_equal(AS.view(np.uint32), ref.ravel().view(np.uint32)))
    alg = (3 + 2 * K) * 4 * N * H * W
    return {"workload": name, "lattice_points": int(m.item()), "ms_graph": ms_graph, "ms_eager": ms_eager, "ms_e2e_host_arrays": ms_e2e,
            "ms_cpu": ms_cpu, "cpu_kind": kind, "cpu_threads": min(os.cpu_count(), N), "bit_identical_to_cpu": same,
            "speedup_device": ms_cpu / ms_graph, "speedup_e2e": ms_cpu / ms_e2e, "algorithmic_bytes": alg,
            "roofline": {"bound": "hbm", "achieved": alg / (ms_graph * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": alg / (ms_graph * 1e-3) / 1e9 / hbm}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--only", type=int, default=-1)
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6535.1)
    B.build()
    for wi, (name, N, K, H, W) in enumerate(WORKLOADS):
        if args.only >= 0 and wi != args.only:
            continue
        if args.profile:
            img, seg = B.synth(N, K, H, W, seed=41)
            d_img, d_seg = torch.from_numpy(img).cuda(), torch.from_numpy(seg).cuda()
            for _ in range(3):
                scd.bilateral_filter(d_img, d_seg, 15.0, 50.0)
            torch.cuda.synchronize()
            continue
        print(json.dumps(run_workload(name, N, K, H, W, iters=args.iters, hbm=hbm)))


if __name__ == "__main__":
    _lib.require_device()
    main()
