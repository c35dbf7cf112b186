"""Developer tool: per-step times of streaming inference (forward + prefetch of the next building) and pool counters."""
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
from detection_3d_b200._lib import lib  # noqa: E402

scn.set_math_mode("bf16")
net = scn.FPN_Net(**scn.sw4c_fpn432_config())
net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
net = net.cuda().eval()
coords = torch.from_numpy(synthetic.building_coords()).cuda()
feats = torch.from_numpy(fpn_util.features_for(coords.cpu().numpy())).cuda()
pre = os.environ.get("PREFETCH", "1") == "1"
n = int(os.environ.get("STEPS", "30"))
evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
host = []
depth = int(os.environ.get("DEPTH", "2"))
with torch.no_grad():
    net([coords, feats]); net([coords, feats])
    if pre:
        for _ in range(depth - 1):
            net.prefetch(coords)
    evs[0].record()
    for i in range(n):
        t0 = time.perf_counter()
        net([coords, feats])
        t1 = time.perf_counter()
        if pre:
            net.prefetch(coords)
        t2 = time.perf_counter()
        evs[i + 1].record()
        host.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, lib().scn_debug_counter(0), lib().scn_debug_counter(1), lib().scn_debug_counter(2)))
torch.cuda.synchronize()
for i in range(n):
    print("step %2d gpu %.2f ms | host run %.2f prefetch %.2f | mallocs %d waits %d pool %d MiB" % ((i, evs[i].elapsed_time(evs[i + 1])) + host[i]))
