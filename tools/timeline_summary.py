"""Summaries of a tools/timeline.py capture (gpurun_out/timeline_<tag>.csv): per-kernel totals, concurrency histogram and the
kernel sequence of the busiest stream between two attention kernels (one HRNet module of the critical chain).
    python tools/timeline_summary.py gpurun_out/timeline_x.csv [out.txt]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = [(float(r["start_us"]), float(r["dur_us"]), r["stream"], r["name"]) for r in csv.DictReader(open(path))]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
span = max(s + d for s, d, _, _ in rows)
tot = sum(d for _, d, _, _ in rows)
print("kernels %d  span %.1f us  sum of kernel durations %.1f us  (average concurrency %.2f)" % (len(rows), span, tot, tot / span), file=out)
loss = next((s for s, d, st, n in rows if "seg_loss_fwd" in n), None)
if loss is not None:
    print("forward %.1f us, loss+backward+optimiser %.1f us" % (loss, span - loss), file=out)
ev = sorted([(s, 1) for s, d, _, _ in rows] + [(s + d, -1) for s, d, _, _ in rows])
hist, cur, last = defaultdict(float), 0, 0.0
for t, k in ev:
    hist[cur] += t - last
    cur += k; last = t
print("time with k kernels running: " + ", ".join("%d: %.0f us" % (k, v) for k, v in sorted(hist.items())), file=out)
agg = defaultdict(lambda: [0.0, 0])
for s, d, st, n in rows:
    a = agg[n[:64]]
    a[0] += d; a[1] += 1
print("\n   total us  share  count  avg us  kernel", file=out)
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    print("%10.0f  %4.1f%%  %5d  %6.1f  %s" % (t, 100 * t / tot, c, t / c, n), file=out)
# critical chain: the stream that runs the attention kernels
att = [(s, d, st, n) for s, d, st, n in rows if "win_attn_fwd" in n]
if len(att) >= 6:
    st0 = att[0][2]
    lo, hi = att[4][0], att[5][0]
    print("\nforward chain of stream %s between the 5th and 6th attention kernels (start, dur, gap before, name):" % st0, file=out)
    prev = None
    for s, d, st, n in rows:
        if st == st0 and lo <= s < hi:
            print("%10.1f %7.1f %7.1f  %s" % (s - lo, d, (s - prev) if prev is not None else 0.0, n[:70]), file=out)
            prev = s + d
att = [(s, d, st, n) for s, d, st, n in rows if "win_attn_bwd" in n]
if len(att) >= 4:
    st0 = att[0][2]
    lo, hi = att[2][0], att[3][0]
    print("\nbackward chain of stream %s between the 3rd and 4th attention-backward kernels:" % st0, file=out)
    prev = None
    for s, d, st, n in rows:
        if st == st0 and lo <= s < hi:
            print("%10.1f %7.1f %7.1f  %s" % (s - lo, d, (s - prev) if prev is not None else 0.0, n[:70]), file=out)
            prev = s + d
