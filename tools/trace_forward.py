"""Developer tool: CUPTI timeline of one backbone forward (torch.profiler) -> gpurun_out/trace_<mode>.json
and a digest: GPU busy time, host time inside CUDA runtime calls, gaps."""
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpn_util  # noqa: E402
import detection_3d_b200.sparseconvnet as scn  # noqa: E402
from detection_3d_b200 import synthetic  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
scn.set_math_mode(mode)
net = scn.FPN_Net(**scn.sw4c_fpn432_config())
net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
net = net.cuda().eval()
coords = torch.from_numpy(synthetic.building_coords()).cuda()
feats = torch.from_numpy(fpn_util.features_for(coords.cpu().numpy())).cuda()
stream_mode = os.environ.get("TRACE_STREAM") == "1"  # streaming: forward + prefetch of the next building, two steps traced
with torch.no_grad():
    for _ in range(12 if stream_mode else 3):
        net([coords, feats])
        if stream_mode:
            net.prefetch(coords)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2 if stream_mode else 1):
            net([coords, feats])
            if stream_mode:
                net.prefetch(coords)
        torch.cuda.synchronize()
out = os.path.join(ROOT, "gpurun_out", f"trace_{mode}.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
prof.export_chrome_trace(out)
ev = json.load(open(out))["traceEvents"]
k = sorted([e for e in ev if e.get("cat") == "kernel"], key=lambda e: e["ts"])
rt = [e for e in ev if e.get("cat") == "cuda_runtime"]
t0, t1 = k[0]["ts"], max(e["ts"] + e["dur"] for e in k)
busy, cur_s, cur_e = 0.0, None, None
for e in k:  # union of kernel intervals
    s, en = e["ts"], e["ts"] + e["dur"]
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, en
    else:
        cur_e = max(cur_e, en)
busy += cur_e - cur_s
print(f"{len(k)} kernels, span {(t1 - t0) / 1e3:.2f} ms, GPU busy (union) {busy / 1e3:.2f} ms, sum of kernel durations {sum(e['dur'] for e in k) / 1e3:.2f} ms")
agg = {}
for e in rt:
    a = agg.setdefault(e["name"], [0, 0.0])
    a[0] += 1
    a[1] += e["dur"]
for n, (c, d) in sorted(agg.items(), key=lambda x: -x[1][1])[:8]:
    print(f"  host {n:32s} {c:5d} calls {d / 1e3:8.2f} ms")
# the ten longest GPU-idle gaps and what ran right after them
gaps = []
end = k[0]["ts"] + k[0]["dur"]
for e in k[1:]:
    if e["ts"] > end:
        gaps.append((e["ts"] - end, e["name"][:50]))
    end = max(end, e["ts"] + e["dur"])
print("idle total %.2f ms in %d gaps; largest:" % (sum(g[0] for g in gaps) / 1e3, len(gaps)))
for g in sorted(gaps, reverse=True)[:10]:
    print("   %.1f us before %s" % g)
rt_sorted = sorted(rt, key=lambda e: -e["dur"])[:14]
tmin = min(e["ts"] for e in rt)
print("longest CUDA runtime calls:")
for e in rt_sorted:
    print("   tid %12d at %9.2f ms  %8.3f ms  %s" % (e["tid"], (e["ts"] - tmin) / 1e3, e["dur"] / 1e3, e["name"]))
