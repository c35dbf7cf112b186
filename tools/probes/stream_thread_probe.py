"""Developer probe: forward time in the main thread / a worker thread, on the default stream / a side stream."""
import os, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpn_util
import detection_3d_b200.sparseconvnet as scn
from detection_3d_b200 import synthetic
scn.set_math_mode("bf16")
dev = torch.device("cuda", 0)
net = scn.FPN_Net(**scn.sw4c_fpn432_config())
net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
net = net.to(dev).eval()
c = synthetic.building_coords()
coords, feats = torch.from_numpy(c).to(dev), torch.from_numpy(fpn_util.features_for(c)).to(dev)
side = torch.cuda.Stream()

def loop(stream, steps):
    with torch.cuda.stream(stream), torch.no_grad():
        for _ in range(steps):
            net([coords, feats])
        stream.synchronize()

def timed(fn):
    fn(5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(100)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 10

def in_thread(stream):
    def f(steps):
        th = threading.Thread(target=loop, args=(stream, steps))
        th.start(); th.join()
    return f

cur = torch.cuda.current_stream()
for name, fn in [("main thread, default stream", lambda k: loop(cur, k)), ("main thread, side stream", lambda k: loop(side, k)),
                 ("worker thread, default stream", in_thread(cur)), ("worker thread, side stream", in_thread(side)),
                 ("main thread, default stream", lambda k: loop(cur, k))]:
    print(f"{name}: {timed(fn):.3f} ms per forward")
