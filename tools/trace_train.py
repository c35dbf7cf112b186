"""Developer tool: kernel-time digest of one 6c_fpn4321 training step (forward + backward) under torch.profiler."""
import collections, os, re, sys
import torch
from torch.profiler import ProfilerActivity, profile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
def step():
    for p in net.parameters():
        p.grad = None
    rpn, roi = net([coords, feats])
    sum((m.features ** 2).sum() for m in rpn + roi).backward()
for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[re.sub(r"\(.*", "", e.name)[:70]]
        a[0] += 1
        a[1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"kernel time {tot / 1e3:.2f} ms")
for n, (cnt, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
    print(f"{t / 1e3:8.2f} ms {cnt:4d}x {n}")
if os.environ.get("LIST"):
    pat = os.environ["LIST"]
    for e in sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and pat in e.name], key=lambda e: e.time_range.start):
        print(f"{e.time_range.start - prof.events()[0].time_range.start:10.0f} {e.device_time:8.1f} us  {e.name[:60]}")
