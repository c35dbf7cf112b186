"""Developer tool: one SubmanifoldConvolution Cin->Cout 3^3 (or 1^3) on level 0 of B470, timed alone with an L2 flush."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import detection_3d_b200.sparseconvnet as scn  # noqa: E402
from detection_3d_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--math", default="bf16")
ap.add_argument("--cin", type=int, default=32)
ap.add_argument("--cout", type=int, default=32)
ap.add_argument("--f", type=int, default=3)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--nx", type=int, default=542)  # footprint of the synthetic building (smaller = a stand-in for the deep levels)
ap.add_argument("--ny", type=int, default=542)
a = ap.parse_args()
scn.set_math_mode(a.math)
L = torch.LongTensor
coords = torch.from_numpy(synthetic.building_coords(nx=a.nx, ny=a.ny)).cuda()
md = scn.Metadata(3)
x0 = torch.empty(0, device="cuda")
scn.SCN.InputLayer_updateOutput(md, L([2048, 2048, 512]), coords, torch.zeros(coords.size(0), 1, device="cuda"), x0, 0, 4)
n = md.getNActive(L([2048, 2048, 512]))
x = torch.randn(n, a.cin, device="cuda")
if a.math == "bf16" and a.cin % 32 == 0:
    x._scn_bf16 = (x.to(torch.bfloat16), x._version)
K = a.f ** 3
w = torch.randn(K, 1, a.cin, a.cout, device="cuda") * 0.05
out = torch.empty(0, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(a.reps):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    macs = scn.SCN.SubmanifoldConvolution_updateOutput(L([2048, 2048, 512]), L([a.f] * 3), md, x, out, w, torch.Tensor())
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts[1:])[len(ts[1:]) // 2]
print(f"Cin={a.cin} Cout={a.cout} f={a.f} math={a.math} n={n} macs={macs:.4g} ms={ms:.3f} TFLOP/s={2 * macs / ms / 1e9:.1f} all={['%.3f' % t for t in ts]}", flush=True)
