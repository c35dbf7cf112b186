"""Developer probe: where config 3's tail (RPN head, anchors, post-processing) spends its time -- host (cProfile) and GPU (events)."""
import cProfile, os, pstats, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpn_util
import detection_3d_b200.sparseconvnet as scn
from detection_3d_b200 import detector, synthetic
scn.set_math_mode("bf16")
dev = torch.device("cuda", 0)
net = scn.FPN_Net(**scn.sw4c_fpn432_config())
net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
net = net.to(dev).eval()
torch.manual_seed(0)
det = detector.SparseRPNDetector(net, detector.RPNModule()).to(dev).eval()
with torch.no_grad():
    det.rpn.head.conv.weight.mul_(10.0); det.rpn.head.cls_logits.weight.mul_(20.0); det.rpn.head.bbox_pred.weight.mul_(1.5)
c = synthetic.building_coords()
coords, feats = torch.from_numpy(c).pin_memory(), torch.from_numpy(fpn_util.features_for(c)).pin_memory()
pts = [coords, feats]
def tail(maps):
    groups = det.rpn(pts, maps)[0]
    return [(b.bbox3d.to("cpu", non_blocking=True), b.get_field("objectness").to("cpu", non_blocking=True)) for gr in groups for b in gr]
with torch.no_grad():
    for _ in range(4):
        maps, _ = net(pts); tail(maps)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); maps, _ = net(pts); e[1].record(); tail(maps); e[2].record(); torch.cuda.synchronize()
        ts.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    print("backbone %.2f ms, tail %.2f ms" % (sum(t[0] for t in ts) / 10, sum(t[1] for t in ts) / 10))
    maps, _ = net(pts); torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(10):
        tail(maps)
    torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
from torch.profiler import ProfilerActivity, profile
import collections, re
with torch.no_grad():
    maps, _ = net(pts); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        tail(maps); torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print(f"{e.time_range.start - t0:8.0f} {e.device_time:7.1f} us  {re.sub(r'\(.*', '', e.name)[:70]}")
