"""Developer probe: 6c_fpn4321 forward in eval (replayed inference program) vs the training forward of a replayed step."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpn_util
import detection_3d_b200.sparseconvnet as scn
from detection_3d_b200 import synthetic
scn.set_math_mode("bf16")
net = scn.FPN_Net(**scn.c6_fpn4321_config())
net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
net = net.cuda()
c = synthetic.building_coords()
coords, feats = torch.from_numpy(c).pin_memory(), torch.from_numpy(fpn_util.features_for(c)).cuda()
def timed(fn, n=10):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
net.eval()
def ev():
    with torch.no_grad(): net([coords, feats])
print("eval forward (host coords) %.2f ms" % timed(ev))
net.train()
def tr():
    rpn, roi = net([coords, feats])
print("train forward only (no backward) %.2f ms" % timed(tr))
