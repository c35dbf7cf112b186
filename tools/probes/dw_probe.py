"""Developer probe: weight gradient of one SubmanifoldConvolution on the tensor-core kernel vs a torch reference."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import detection_3d_b200.sparseconvnet as scn
from detection_3d_b200 import synthetic
L = torch.LongTensor
math = sys.argv[1] if len(sys.argv) > 1 else "bf16"
cin, cout, f = int(sys.argv[2]) if len(sys.argv) > 2 else 128, int(sys.argv[3]) if len(sys.argv) > 3 else 128, 3
c = synthetic.small_building(40, 36, 12, 3, seed=4)
md = scn.Metadata(3)
x0 = torch.empty(0, device="cuda")
scn.SCN.InputLayer_updateOutput(md, L([64, 64, 32]), torch.from_numpy(c), torch.zeros(len(c), 1, device="cuda"), x0, 0, 4)
n = md.getNActive(L([64, 64, 32]))
torch.manual_seed(0)
x = torch.randn(n, cin, device="cuda"); dy = torch.randn(n, cout, device="cuda")
w = torch.randn(f ** 3, 1, cin, cout, device="cuda") * 0.05
res = {}
for mode in ("fp32", math):
    scn.set_math_mode(mode)
    out = torch.empty(0, device="cuda")
    scn.SCN.SubmanifoldConvolution_updateOutput(L([64, 64, 32]), L([f] * 3), md, x, out, w, torch.Tensor())
    din, dw = torch.empty(0, device="cuda"), torch.zeros_like(w)
    scn.SCN.SubmanifoldConvolution_backward(L([64, 64, 32]), L([f] * 3), md, x, din, dy, w, dw, torch.Tensor())
    torch.cuda.synchronize()
    res[mode] = dw.clone()
a, b = res["fp32"], res[math]
print("n", n, "ref absmax %.3f got absmax %.3f maxdiff %.4f" % (a.abs().max(), b.abs().max(), (a - b).abs().max()))
for k in (0, 13, 26):
    print(k, "ref", a[k, 0, :2, :4].flatten().tolist(), "\n   got", b[k, 0, :2, :4].flatten().tolist())
print("nonzero fraction got", float((b != 0).float().mean()), "corr", float((a * b).sum() / (a.norm() * b.norm() + 1e-9)))
