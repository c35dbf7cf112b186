"""Developer tool: per-parameter gradient error of one whole-network training step (mini4 backbone) against the reference's
autograd (tests/golden/fpn_mini4_train.npz), in every math mode."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpn_util  # noqa: E402
import detection_3d_b200.sparseconvnet as scn  # noqa: E402
from detection_3d_b200 import synthetic  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "fpn_mini4_train.npz"))
cfg = dict(fpn_util.mini4_config(), track_running_stats=True)
coords = synthetic.building_coords(nx=60, ny=56, nz=24, n_walls=3, seed=3)
feats = fpn_util.features_for(coords)
for math in sys.argv[1:] or ["fp32", "tf32", "bf16"]:
    scn.set_math_mode(math)
    net = scn.FPN_Net(**cfg)
    net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
    net = net.cuda().train()
    rpn, roi = net([torch.from_numpy(coords), torch.from_numpy(feats).cuda()])
    loss = sum((m.features ** 2).mean() for m in rpn + roi)
    loss.backward()
    torch.cuda.synchronize()
    print(f"== {math}: loss {float(loss):.6f} (reference {float(g['loss']):.6f})")
    rows = []
    for k, p in net.named_parameters():
        if p.grad is None:
            continue
        ref = g["grad:" + k].astype(np.float64)
        d = p.grad.cpu().numpy().astype(np.float64) - ref
        rows.append((float(np.abs(d).max() / max(1e-6, np.abs(ref).max())), float(np.sqrt((d * d).mean()) / max(1e-12, np.sqrt((ref * ref).mean()))), k, tuple(p.shape)))
    for err, rms, k, shp in rows:
        print("   %-36s %-18s max %.3e  rms %.3e" % (k, shp, err, rms))
scn.set_math_mode("fp32")
