"""Developer tool: the strided convolution, the deconvolution (+ lateral) and the weight-gradient kernel on B470 level 0 / 1, each
timed alone with an L2 flush -- the launches `ncu --set full -k regex:conv_plan_tc|conv_dw_tc` profiles for profiles/r2_*_ncu.txt.

    python tools/strided_kernels.py conv|deconv|dw [--math bf16]
"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import detection_3d_b200.sparseconvnet as scn  # noqa: E402
from detection_3d_b200 import synthetic  # noqa: E402
from detection_3d_b200._lib import check, l3, lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["conv", "deconv", "dw"])
ap.add_argument("--math", default="bf16")
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--sorted", action="store_true", help="input points in spatial order (rows then numbered like the internally numbered Metadata of a replayed step)")
a = ap.parse_args()
scn.set_math_mode(a.math)
L = torch.LongTensor
full, half = [2048, 2048, 512], [1024, 1024, 256]
c_np = synthetic.building_coords()
if a.sorted:
    import numpy as np
    key = (c_np[:, 0] // 8 * 4096 + c_np[:, 1] // 8) * 4096 + c_np[:, 2] // 8  # 8^3 blocks, then row-major inside
    c_np = c_np[np.lexsort((c_np[:, 2], c_np[:, 1], c_np[:, 0], key))]
coords = torch.from_numpy(c_np).cuda()
md = scn.Metadata(3)
scn.SCN.InputLayer_updateOutput(md, L(full), coords, torch.zeros(coords.size(0), 1, device="cuda"), torch.empty(0, device="cuda"), 0, 4)
n0 = md.getNActive(L(full))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())


def shadow(x):
    if a.math == "bf16":
        x._scn_bf16 = (x.to(torch.bfloat16), x._version)
    return x


if a.what == "conv":     # m_downs.1: BatchNorm output (32 channels) -> Convolution 2^3 / 2, 32 -> 64
    x = shadow(torch.randn(n0, 32, device="cuda"))
    w = torch.randn(8, 1, 32, 64, device="cuda") * 0.05
    out = torch.empty(0, device="cuda")
    run = lambda: scn.SCN.Convolution_updateOutput(L(full), L(half), L([2] * 3), L([2] * 3), md, x, out, w, torch.Tensor())
    bytes_alg = lambda: n0 * 32 * (2 if a.math == "bf16" else 4) + out.numel() * 4 + n0 * 4 * 2
elif a.what == "deconv":  # m_ups.7 + m_shortcuts.0 folded in: Deconvolution 2^3 / 2, 128 -> 128, + lateral 32 -> 128, level 1 -> level 0
    scn.SCN.Convolution_updateOutput(L(full), L(half), L([2] * 3), L([2] * 3), md, torch.zeros(n0, 4, device="cuda"), torch.empty(0, device="cuda"),
                                     torch.zeros(8, 1, 4, 32, device="cuda"), torch.Tensor())
    n1 = md.getNActive(L(half))
    x = shadow(torch.randn(n1, 128, device="cuda"))
    w = torch.randn(8, 1, 128, 128, device="cuda") * 0.05
    y = torch.randn(n0, 32, device="cuda")
    y16 = y.to(torch.bfloat16)
    wl = torch.randn(32, 128, device="cuda") * 0.05
    out = torch.empty(n0, 128, device="cuda")
    macs_c = C.c_double()

    def run():
        check(lib().scn_fuse_next_lateral(p(y), p(y16), p(wl), 0, 32, n0))
        check(lib().scn_deconvolution_forward(md._h, l3(half), l3(full), l3([2] * 3), l3([2] * 3), p(x), p(out), p(w), None, 128, 128, C.byref(macs_c),
                                              p(x._scn_bf16[0]) if a.math == "bf16" else None, 0, None, None))
        lib().scn_fuse_result(None, None)
        return macs_c.value + n0 * 32 * 128
    bytes_alg = lambda: n1 * 128 * 2 + n0 * 32 * 2 + n0 * 128 * 4 + n0 * 4 * 2
else:                     # weight gradient of SubmanifoldConvolution 128 -> 128, 3^3, level 0 (conv_dw_tc in bf16 mode)
    x = torch.randn(n0, 128, device="cuda")
    dy = torch.randn(n0, 128, device="cuda")
    w = torch.randn(27, 1, 128, 128, device="cuda") * 0.02
    out = torch.empty(0, device="cuda")
    scn.SCN.SubmanifoldConvolution_updateOutput(L(full), L([3] * 3), md, x, out, w, torch.Tensor())
    dw = torch.zeros_like(w)

    def run():
        scn.SCN.SubmanifoldConvolution_backward(L(full), L([3] * 3), md, x, None, dy, w, dw, torch.Tensor())
        return 10715792 * 128 * 128
    bytes_alg = lambda: 2 * n0 * 128 * 4 + 10715792 * 8
ts = []
for i in range(a.reps):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    macs = run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts[1:])[len(ts[1:]) // 2]
print(f"{a.what} math={a.math} macs={macs:.4g} ms={ms:.3f} TFLOP/s={2 * macs / ms / 1e9:.1f} algorithmic GB={bytes_alg() / 1e9:.3f} -> {bytes_alg() / ms / 1e6:.0f} GB/s all={['%.3f' % t for t in ts]}", flush=True)
