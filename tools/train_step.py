"""BASELINE.json config 5: 6c_fpn4321 backbone training step (forward + backward, bs 1 per GPU, one B470 building per rank)
with the data-parallel gradient all-reduce over NCCL.  Launch with torchrun (1..8 ranks) or plain python (1 rank).
Prints per-phase times (CUDA events, max over ranks)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpn_util  # noqa: E402
import detection_3d_b200.sparseconvnet as scn  # noqa: E402
from detection_3d_b200 import distributed, synthetic  # noqa: E402

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
scn.set_math_mode(os.environ.get("SCN_MATH", "bf16"))
net = scn.FPN_Net(**scn.c6_fpn4321_config())
net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
net = net.to(dev).train()
coords_np = synthetic.building_coords(seed=rank)  # every rank its own building
coords = torch.from_numpy(coords_np)
feats = torch.from_numpy(fpn_util.features_for(coords_np)).to(dev)
params = [p for p in net.parameters() if p.requires_grad]
steps = int(os.environ.get("STEPS", "5"))
times = []
for it in range(steps + 2):
    for p in params:
        p.grad = None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    rpn, roi = net([coords, feats])
    loss = sum((m.features ** 2).sum() for m in rpn + roi)
    ev[1].record()
    loss.backward()
    ev[2].record()
    n_coll = distributed.allreduce_gradients(params)
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 2:
        times.append([ev[i].elapsed_time(ev[i + 1]) for i in range(3)])
t = torch.tensor(times, device=dev, dtype=torch.float64).mean(0)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
nbytes = sum(p.numel() * 4 for p in params)
with_grad = sum(1 for p in params if p.grad is not None)
if rank == 0:
    print(f"train step 6c_fpn4321 bs1/GPU x{world}: forward {t[0]:.2f} ms, backward {t[1]:.2f} ms, gradient all-reduce {t[2]:.2f} ms "
          f"({nbytes / 1e6:.1f} MB fp32 in {n_coll} collectives), loss {float(loss):.4e}, {with_grad}/{len(params)} parameters with gradients", flush=True)
if world > 1:
    dist.destroy_process_group()
