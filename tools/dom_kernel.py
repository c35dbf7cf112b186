"""Runs only the dominant layer (SubmanifoldConvolution C->C 3^3 on level 0 of B470) a few times:
the command ncu profiles (`ncu -k regex:conv_plan_tc ...`)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import detection_3d_b200.sparseconvnet as scn  # noqa: E402
from detection_3d_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--math", default="tf32")
ap.add_argument("--c", type=int, default=128)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--internal", action="store_true", help="internally numbered Metadata (row id = spatial rank), as a replayed forward uses")
ap.add_argument("--sorted", action="store_true", help="input points in spatial (8^3 block) order: rows numbered like the internally numbered Metadata of a replayed forward")
a = ap.parse_args()
scn.set_math_mode(a.math)
L = torch.LongTensor
c_np = synthetic.building_coords()
if a.sorted:
    import numpy as np
    key = (c_np[:, 0] // 8 * 4096 + c_np[:, 1] // 8) * 4096 + c_np[:, 2] // 8
    c_np = c_np[np.lexsort((c_np[:, 2], c_np[:, 1], c_np[:, 0], key))]
coords = torch.from_numpy(c_np).cuda()
md = scn.Metadata(3)
if a.internal:
    from detection_3d_b200._lib import check, lib
    check(lib().scn_metadata_set_internal_numbering(md._h, 1))
x0 = torch.empty(0, device="cuda")
scn.SCN.InputLayer_updateOutput(md, L([2048, 2048, 512]), coords, torch.zeros(coords.size(0), 1, device="cuda"), x0, 0, 4)
n = md.getNActive(L([2048, 2048, 512]))
x = torch.randn(n, a.c, device="cuda")
if a.math == "bf16":  # as inside the network: the producer (BatchNorm / add kernel) has already written the bf16 copy
    x._scn_bf16 = (x.to(torch.bfloat16), x._version)
w = torch.randn(27, 1, a.c, a.c, device="cuda") * 0.02
out = torch.empty(0, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(a.reps):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    macs = scn.SCN.SubmanifoldConvolution_updateOutput(L([2048, 2048, 512]), L([3, 3, 3]), md, x, out, w, torch.Tensor())
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts[1:])[len(ts[1:]) // 2]
print(f"C={a.c} math={a.math} n={n} macs={macs:.4g} ms={ms:.3f} TFLOP/s={2 * macs / ms / 1e9:.1f} all={['%.3f' % t for t in ts]}")
