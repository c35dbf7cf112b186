"""Developer tool: CUPTI trace of e2e-style steps (fresh H2D inputs, results kept alive across steps);
prints the longest CUDA runtime calls per thread."""
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

scn.set_math_mode("bf16")
dev = torch.device("cuda", 0)
net = scn.FPN_Net(**scn.sw4c_fpn432_config())
net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
net = net.cuda().eval()
c = synthetic.building_coords()
coords_pin = torch.from_numpy(c).pin_memory()
feats_pin = torch.from_numpy(fpn_util.features_for(c)).pin_memory()


def step():
    cc = coords_pin.to(dev, non_blocking=True)
    ff = feats_pin.to(dev, non_blocking=True)
    rpn, roi = net([cc, ff])
    host = [m.features.to("cpu", non_blocking=True) for m in rpn + roi]
    torch.cuda.synchronize()
    return rpn, roi, host


with torch.no_grad():
    keep = None
    for _ in range(4):
        keep = step()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            keep = step()
out = os.path.join(ROOT, "gpurun_out", "trace_e2e.json")
prof.export_chrome_trace(out)
ev = json.load(open(out))["traceEvents"]
rt = [e for e in ev if e.get("cat") == "cuda_runtime"]
t0 = min(e["ts"] for e in rt)
for e in sorted(rt, key=lambda e: -e["dur"])[:25]:
    print("tid %12d at %8.2f ms dur %8.3f ms %s" % (e["tid"], (e["ts"] - t0) / 1e3, e["dur"] / 1e3, e["name"]))
