import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpn_util
import detection_3d_b200.sparseconvnet as scn
from detection_3d_b200 import synthetic
for math in ("fp32", "bf16"):
    scn.set_math_mode(math)
    cfg = fpn_util.mini4_config()
    net = scn.FPN_Net(**cfg)
    net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
    net = net.cuda().eval()
    big = synthetic.building_coords(nx=44, ny=40, nz=20, n_walls=3, seed=9)
    with torch.no_grad():
        for name, c in (("building", big), ("one point", big[:1]), ("two points far apart", np.array([[0, 0, 0, 0], [40, 39, 19, 0]])), ("empty", big[:0]), ("building again", big)):
            c = np.ascontiguousarray(c)
            f = torch.from_numpy(fpn_util.features_for(c) if len(c) else np.zeros((0, 9), np.float32)).cuda()
            try:
                for rep in range(2):
                    rpn, roi = net([torch.from_numpy(c), f])
                torch.cuda.synchronize()
                print(math, name, [tuple(m.features.shape) for m in rpn + roi], all(bool(torch.isfinite(m.features).all()) for m in rpn + roi))
            except Exception as e:
                print(math, name, "raised", type(e).__name__, str(e)[:120])
