"""Developer tool: throughput of T backbone forwards in flight (one host thread, one CUDA stream and one recorded program each)
against one at a time.  python tools/concurrent_forwards.py [threads] [steps]"""
import os, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpn_util
import detection_3d_b200.sparseconvnet as scn
from detection_3d_b200 import synthetic

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
e2e = os.environ.get("E2E", "0") == "1"
scn.set_math_mode(os.environ.get("SCN_MATH", "bf16"))
dev = torch.device("cuda", 0)
cfg = scn.sw4c_fpn432_config()
nets = []
for t in range(T):
    net = scn.FPN_Net(**cfg)
    net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
    nets.append(net.to(dev).eval())
c = synthetic.building_coords()
f = fpn_util.features_for(c)
coords = [torch.from_numpy(c).pin_memory() if e2e else torch.from_numpy(c).to(dev) for _ in range(T)]
feats = [torch.from_numpy(f).pin_memory() if e2e else torch.from_numpy(f).to(dev) for _ in range(T)]
streams = [torch.cuda.Stream() for _ in range(T)]
sums = [None] * T

def work(t, steps):
    with torch.cuda.stream(streams[t]), torch.no_grad():
        for _ in range(steps):
            rpn, roi = nets[t]([coords[t], feats[t]])
            if e2e:
                host = [m.features.to("cpu", non_blocking=True) for m in rpn + roi]
        streams[t].synchronize()
        sums[t] = [float(m.features.double().abs().sum()) for m in rpn + roi]

def run(n_threads, steps):
    th = [threading.Thread(target=work, args=(t, steps)) for t in range(n_threads)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3

if os.environ.get("PRE", "1") == "1":
    run(T, 5)
for n in ([1, T, 1] if T > 1 else [1]):
    run(n, 5)
    ms = run(n, K)
    print(f"{n} in flight: {ms / (n * K):.3f} ms per forward ({n * K / ms * 1e3:.1f} buildings/s)  e2e={e2e}")
print("checksums thread 0:", ["%.6g" % v for v in sums[0]])
for t in range(1, T):
    print(f"checksums thread {t} rel diff:", ["%.2e" % (abs(a - b) / max(1e-9, abs(b))) for a, b in zip(sums[t], sums[0])])
