"""Developer tool: time of the Metadata build alone (input layer, then every rulebook of the recorded backbone program
through the prefetch workers) on an otherwise idle GPU."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import detection_3d_b200.sparseconvnet as scn  # noqa: E402
from detection_3d_b200 import synthetic  # noqa: E402

L = torch.LongTensor
coords = torch.from_numpy(synthetic.building_coords()).cuda()
sizes = [[2048 >> k, 2048 >> k, max(1, 512 >> k)] for k in range(9)]
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    md = scn.Metadata(3)
    x0 = torch.empty(0, device="cuda")
    scn.SCN.InputLayer_updateOutput(md, L(sizes[0]), coords, torch.zeros(coords.size(0), 1, device="cuda"), x0, 0, 4)
    t1 = time.perf_counter()
    ts = []
    for k in range(8):
        x = torch.zeros(md.getNActive(L(sizes[k])), 32, device="cuda")
        w = torch.zeros(8, 1, 32, 32, device="cuda")
        out = torch.empty(0, device="cuda")
        scn.SCN.Convolution_updateOutput(L(sizes[k]), L(sizes[k + 1]), L([2, 2, 2]), L([2, 2, 2]), md, x, out, w, torch.Tensor())
        torch.cuda.synchronize()
        ts.append(time.perf_counter())
    print("input %.0f us | levels " % ((t1 - t0) * 1e6) + " ".join("%.0f" % ((b - a) * 1e6) for a, b in zip([t1] + ts[:-1], ts)) + " | total %.0f us" % ((ts[-1] - t0) * 1e6), flush=True)
