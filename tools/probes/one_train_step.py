"""Developer probe: two 6c_fpn4321 training steps (for ncu captures of the backward kernels)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpn_util
import detection_3d_b200.sparseconvnet as scn
from detection_3d_b200 import synthetic
scn.set_math_mode(os.environ.get("SCN_MATH", "bf16"))
net = scn.FPN_Net(**scn.c6_fpn4321_config())
net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
net = net.cuda().train()
c = synthetic.building_coords()
coords, feats = torch.from_numpy(c), torch.from_numpy(fpn_util.features_for(c)).cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    for p in net.parameters():
        p.grad = None
    rpn, roi = net([coords, feats])
    sum((m.features ** 2).sum() for m in rpn + roi).backward()
torch.cuda.synchronize()
print("ok")
