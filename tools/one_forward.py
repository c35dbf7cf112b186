"""N backbone forwards of B470 (resident inputs) -- the command the ncu launch list is taken from."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpn_util  # noqa: E402
import detection_3d_b200.sparseconvnet as scn  # noqa: E402
from detection_3d_b200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
scn.set_math_mode(sys.argv[2] if len(sys.argv) > 2 else "tf32")
net = scn.FPN_Net(**scn.sw4c_fpn432_config())
net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
net = net.cuda().eval()
coords = torch.from_numpy(synthetic.building_coords()).cuda()
feats = torch.from_numpy(fpn_util.features_for(coords.cpu().numpy())).cuda()
with torch.no_grad():
    for i in range(n):
        l0 = scn.kernel_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        net([coords, feats])
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"forward {i}: {e0.elapsed_time(e1):.2f} ms on the stream, host returned after {1e3 * (t1 - t0):.2f} ms, {scn.kernel_launch_count() - l0} launches of ours")
