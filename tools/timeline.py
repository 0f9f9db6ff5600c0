"""Kernel timeline of ONE graph-replayed training step (CUPTI activity records through torch.profiler; nsys is not installed):
per-stream busy time, idle gaps between consecutive kernels of the main stream, overlap of the weight-gradient side stream, and the
per-kernel-family totals as they are inside the step (warm L2, concurrent streams) — the numbers the serialised ncu launch list
cannot give.  Usage: timeline.py [out.json]   (B, S, DP env as tools/op_table.py)"""
import json
import os
import re
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import T96, synth_batch  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss  # noqa: E402
from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys  # noqa: E402

B, S = int(os.environ.get("B", 16)), int(os.environ.get("S", 512))
dev = torch.device("cuda:0")
m = MSUNetSys(img_size=S, drop_path_rate=float(os.environ.get("DP", 0.1)), **T96).to(dev).train()
crit = DynamicLoss(alpha=0.2, beta=0.8, tversky_bce_mix=0.45)
x, y = synth_batch(B, S, 1)
x, y = x.to(dev), y.to(dev)


def step():
    for p in m.parameters():
        p.grad = None
    crit(m(x), y).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ks = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "stream", None)) for e in ev), key=lambda t: t[0])
if not ks:
    print("no CUDA kernel records (CUPTI unavailable?)")
    sys.exit(0)
t0, t1 = ks[0][0], max(k[1] for k in ks)


def fam(name):
    n = re.sub(r"^void\s+", "", name)
    n = re.sub(r"^msu::", "", n)
    return re.split(r"[<(]", n)[0][:40]


by_stream = defaultdict(list)
for s_, e_, n_, st in ks:
    by_stream[st].append((s_, e_, n_))
out = {"step_us": t1 - t0, "kernels": len(ks), "streams": {}}
main = max(by_stream, key=lambda s: sum(e - b for b, e, _ in by_stream[s]))
for st, lst in by_stream.items():
    busy = sum(e - b for b, e, _ in lst)
    gaps = [lst[i + 1][0] - lst[i][1] for i in range(len(lst) - 1)]
    pos = [g_ for g_ in gaps if g_ > 0]
    out["streams"][str(st)] = {"kernels": len(lst), "busy_us": busy, "gap_us": sum(pos), "gaps_over_5us": sum(1 for g_ in pos if g_ > 5),
                               "median_gap_us": sorted(pos)[len(pos) // 2] if pos else 0, "is_main": st == main}
# union busy time over all streams (any kernel running)
iv = sorted((b, e) for b, e, _, _ in ks)
cov, cur_b, cur_e = 0.0, iv[0][0], iv[0][1]
for b, e in iv[1:]:
    if b > cur_e:
        cov += cur_e - cur_b
        cur_b, cur_e = b, e
    else:
        cur_e = max(cur_e, e)
cov += cur_e - cur_b
out["any_kernel_running_us"] = cov
out["idle_us"] = (t1 - t0) - cov
fams = defaultdict(lambda: [0.0, 0])
for b, e, n_, st in ks:
    f = fams[fam(n_) + ("" if st == main else " [side]")]
    f[0] += e - b
    f[1] += 1
out["families_us"] = {k: {"us": round(v[0], 1), "n": v[1]} for k, v in sorted(fams.items(), key=lambda kv: -kv[1][0])}
# the 12 largest main-stream gaps with the kernels around them
lst = by_stream[main]
big = sorted(((lst[i + 1][0] - lst[i][1], fam(lst[i][2]), fam(lst[i + 1][2])) for i in range(len(lst) - 1)), reverse=True)[:12]
out["largest_main_gaps"] = [{"gap_us": round(g_, 1), "after": a, "before": b_} for g_, a, b_ in big]
# the first kernels of the replay and the surroundings of the largest gap (start offset, duration, name)
out["first_kernels"] = [{"t_us": round(b - t0, 1), "dur_us": round(e - b, 1), "name": fam(n_)} for b, e, n_, _ in ks[:24]]
gi = max(range(len(ks) - 1), key=lambda i: ks[i + 1][0] - max(k[1] for k in ks[:i + 1]))
out["around_largest_idle"] = [{"t_us": round(b - t0, 1), "dur_us": round(e - b, 1), "name": fam(n_)} for b, e, n_, _ in ks[max(0, gi - 6):gi + 6]]
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
