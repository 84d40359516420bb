"""Offline analysis of tools/step_timeline.py output: how full the GPU is over one training step and which kernels run
while it is mostly empty."""
import json, re, sys, collections

ev = [json.loads(l) for l in open(sys.argv[1])]
ev.sort(key=lambda e: e["ts"])
marks = [e["ts"] for e in ev if "diversity_finalize" in e["name"]]
t0, t1 = marks[0], marks[1]  # one full step between two occurrences
step = [e for e in ev if t0 <= e["ts"] < t1]
print("step %.3f ms, %d kernels, streams %s" % ((t1 - t0) / 1e3, len(step), sorted({e["stream"] for e in step})))

def short(n):
    n = re.sub(r"^void\s+", "", n)
    n = re.sub(r"\(anonymous namespace\)::|<unnamed>::|at::native::", "", n)
    return re.sub(r"[<(].*$", "", n)[:34]

def weight(e):
    g, b = e["grid"], e["block"]
    ctas = g[0] * g[1] * g[2]
    thr = b[0] * b[1] * b[2]
    big = (e["smem"] or 0) > 100000
    per_sm = 1 if big else max(1, min(16, 2048 // max(thr, 1)))
    return min(1.0, ctas / (148.0 * per_sm))

pts = []
for i, e in enumerate(step):
    pts.append((e["ts"], 1, i))
    pts.append((e["ts"] + e["dur"], 0, i))
pts.sort()
running = set()
last = t0
hist = collections.Counter()
low_by_kernel = collections.defaultdict(float)
busy_sm_time = 0.0
for t, kind, i in pts:
    t = min(t, t1)
    dt = t - last
    if dt > 0:
        occ = min(1.0, sum(weight(step[j]) for j in running))
        busy_sm_time += occ * dt
        bucket = "idle" if not running else ("<25%" if occ < 0.25 else ("<50%" if occ < 0.5 else ("<100%" if occ < 0.999 else "full")))
        hist[bucket] += dt
        if occ < 0.5:
            if running:
                for j in running:
                    low_by_kernel[short(step[j]["name"])] += dt / len(running)
            else:
                low_by_kernel["(nothing running)"] += dt
        last = t
    if kind == 1:
        running.add(i)
    else:
        running.discard(i)
tot = t1 - t0
print("GPU fill over the step: " + ", ".join("%s %.2f ms (%.0f%%)" % (k, v / 1e3, 100 * v / tot) for k, v in hist.most_common()))
print("SM-time (fill-weighted) %.2f ms of %.2f ms" % (busy_sm_time / 1e3, tot / 1e3))
print("time below 50%% fill, attributed to what runs then:")
for k, v in sorted(low_by_kernel.items(), key=lambda kv: -kv[1])[:25]:
    print("  %-36s %7.1f us" % (k, v))
# per-stream busy time
bs = collections.defaultdict(float)
for e in step:
    bs[e["stream"]] += e["dur"]
nccl = [e for e in step if "nccl" in e["name"].lower()]
for e in nccl:
    print("NCCL kernel at +%.3f ms: %.3f ms  grid %s  %s" % ((e["ts"] - t0) / 1e3, e["dur"] / 1e3, e["grid"], e["name"][:50]))
print("per-stream kernel time (ms):", {k: round(v / 1e3, 2) for k, v in bs.items()})
# coarse timeline: 0.25 ms slots with fill
slot = 250.0
nslots = int(tot / slot) + 1
fill = [0.0] * nslots
for e in step:
    a, b, w = e["ts"] - t0, min(e["ts"] + e["dur"], t1) - t0, weight(e)
    s = int(a / slot)
    while s < nslots and s * slot < b:
        lo, hi = max(a, s * slot), min(b, (s + 1) * slot)
        if hi > lo:
            fill[s] += w * (hi - lo) / slot
        s += 1
print("fill per 0.25 ms slot:", " ".join("%d" % min(9, int(f * 10 / 1.0)) if f < 1 else "F" for f in fill))
