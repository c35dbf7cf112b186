python tools/trace_train.py 2>&1 | tail -30
python - <<'PY'
import os, sys, time, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import fpn_util
import detection_3d_b200.sparseconvnet as scn
from detection_3d_b200 import synthetic
scn.set_math_mode("bf16")
net = scn.FPN_Net(**scn.c6_fpn4321_config()); net.load_state_dict(fpn_util.deterministic_state(net, seed=1)); net = net.cuda().train()
c = synthetic.building_coords(); coords, feats = torch.from_numpy(c), torch.from_numpy(fpn_util.features_for(c)).cuda()
def step(sync_mid):
    for p in net.parameters(): p.grad = None
    t0 = time.perf_counter(); rpn, roi = net([coords, feats]); loss = sum((m.features ** 2).sum() for m in rpn + roi); t1 = time.perf_counter()
    if sync_mid: torch.cuda.synchronize()
    t2 = time.perf_counter(); loss.backward(); t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3
for _ in range(3): step(True)
for s in (True, True, False, False):
    print("host fwd submit %.2f ms | gpu tail after fwd %.2f | host bwd submit %.2f | gpu tail after bwd %.2f  (sync_mid=%s)" % (*step(s), s))
PY
