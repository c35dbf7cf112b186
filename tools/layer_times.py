"""Per-module CUDA-event times of one backbone forward (developer tool; synchronises per module)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpn_util  # noqa: E402
import detection_3d_b200.sparseconvnet as scn  # noqa: E402
from detection_3d_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--math", default="fp32")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
scn.set_math_mode(a.math)
cfg = scn.sw4c_fpn432_config()
net = scn.FPN_Net(**cfg)
net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
net = net.cuda().eval()
coords = torch.from_numpy(synthetic.building_coords()).cuda()
feats = torch.from_numpy(fpn_util.features_for(coords.cpu().numpy())).cuda()
times = {}
leaf = (scn.InputLayer, scn.SubmanifoldConvolution, scn.Convolution, scn.Deconvolution, scn.BatchNormalization, scn.AddTable)
for name, m in net.named_modules():
    if isinstance(m, leaf):
        def pre(mod, inp, name=name):
            torch.cuda.synchronize()
            mod._ev = torch.cuda.Event(enable_timing=True)
            mod._ev.record()

        def post(mod, inp, out, name=name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            torch.cuda.synchronize()
            times.setdefault(name, []).append(mod._ev.elapsed_time(e))
        m.register_forward_pre_hook(pre)
        m.register_forward_hook(post)
with torch.no_grad():
    for _ in range(a.reps):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); net([coords, feats]); t1.record(); torch.cuda.synchronize()
        print("forward (with per-module syncs): %.2f ms" % t0.elapsed_time(t1))
tot = 0
for name, v in times.items():
    m = dict(net.named_modules())[name]
    print("%-28s %-44s %9.3f ms" % (name, repr(m)[:44], v[-1]))
    tot += v[-1]
print("sum of modules: %.2f ms" % tot)
