"""Developer tool: per-step end-to-end times (host pinned inputs -> outputs on host), as bench.py's e2e leg."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpn_util  # noqa: E402
import detection_3d_b200.sparseconvnet as scn  # noqa: E402
from detection_3d_b200 import synthetic  # noqa: E402

scn.set_math_mode(sys.argv[1] if len(sys.argv) > 1 else "bf16")
dev = torch.device("cuda", 0)
net = scn.FPN_Net(**scn.sw4c_fpn432_config())
net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
net = net.cuda().eval()
c = synthetic.building_coords()
coords_pin = torch.from_numpy(c).pin_memory()
feats_pin = torch.from_numpy(fpn_util.features_for(c)).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with torch.no_grad():
    for i in range(12):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ff = feats_pin.to(dev, non_blocking=True)
        t1 = time.perf_counter()
        rpn, roi = net([coords_pin if len(sys.argv) < 3 else coords_pin.to(dev, non_blocking=True), ff])
        t2 = time.perf_counter()
        host = [m.features.to("cpu", non_blocking=True) for m in rpn + roi]
        b.record()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print(f"step {i}: events {a.elapsed_time(b):.2f} ms | host: copies issued {1e3*(t1-t0):.2f}, forward returned {1e3*(t2-t0):.2f}, done {1e3*(t3-t0):.2f}")
