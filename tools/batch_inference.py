"""BASELINE.json config 4: batched inference over 64 synthetic buildings (B470 footprint jittered +-15 %, seeds 0..63),
sharded by building across the ranks (longest first, no collective on the data path), every rank streaming its shard
through the backbone with the Metadata built two buildings ahead.  Launch with torchrun (1..8 ranks) or plain python.
Prints buildings/s over the whole batch (wall clock from a barrier to the last rank's completion)."""
import os
import sys
import time

import numpy as np
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
n_build = int(os.environ.get("BUILDINGS", "64"))
scn.set_math_mode("bf16")
net = scn.FPN_Net(**scn.sw4c_fpn432_config())
net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
net = net.to(dev).eval()

# every rank generates the same list (cheap, deterministic) and keeps its shard, inputs in pinned host memory
dims = []
for seed in range(n_build):
    rs = np.random.RandomState(1000 + seed)
    dims.append((int(542 * (1 + 0.15 * (2 * rs.rand() - 1))), int(542 * (1 + 0.15 * (2 * rs.rand() - 1)))))
sizes = [2 * nx * ny + 16 * 68 * (nx + ny) // 2 for nx, ny in dims]
mine = distributed.shard_buildings(sizes, world, rank)
inputs = []
for i in mine:
    c = synthetic.building_coords(nx=dims[i][0], ny=dims[i][1], nz=68, seed=i)
    inputs.append((i, torch.from_numpy(c).pin_memory(), torch.from_numpy(fpn_util.features_for(c)).pin_memory()))
with torch.no_grad():
    warm = synthetic.building_coords(nx=620, ny=620, nz=68, seed=99)  # at least as large as any building of the batch
    wc, wf = torch.from_numpy(warm).pin_memory(), torch.from_numpy(fpn_util.features_for(warm)).pin_memory()
    net([wc, wf])  # records the program
    for _ in range(2):
        net.prefetch(wc)
    for j in range(6):  # streams a few buildings so that the Metadata / register pools reach their steady size
        net([wc, wf])
        if j < 4:
            net.prefetch(wc)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    results = {}
    depth = int(os.environ.get("DEPTH", "2"))
    for j in range(min(depth, len(inputs))):
        net.prefetch(inputs[j][1])
    for j, (i, c, f) in enumerate(inputs):
        rpn, roi = net([c, f])
        if j + depth < len(inputs):
            net.prefetch(inputs[j + depth][1])
        results[i] = [(tuple(m.features.shape), float(m.features.abs().sum())) for m in rpn + roi]  # D2H of a checksum per map
        if os.environ.get("VERBOSE") == "1" and rank == 0:
            from detection_3d_b200._lib import lib
            print("  building %2d rows %8d  t=%.1f ms  mallocs %d waits %d pool %d MiB" % (i, c.size(0), (time.perf_counter() - t0) * 1e3,
                  lib().scn_debug_counter(0), lib().scn_debug_counter(1), lib().scn_debug_counter(2)), flush=True)
    torch.cuda.synchronize()
    t_local = time.perf_counter() - t0
t = torch.tensor([t_local], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
merged = distributed.gather_results(results)
if rank == 0:
    assert len(merged) == n_build
    vox = sum(sizes)
    print(f"batched inference: {n_build} buildings ({vox / 1e6:.1f} M input rows) on {world} GPU(s): {t.item() * 1e3:.1f} ms "
          f"= {n_build / t.item():.1f} buildings/s ({t.item() * 1e3 / n_build:.2f} ms per building), shard sizes "
          f"{[len(distributed.shard_buildings(sizes, world, r)) for r in range(world)]}", flush=True)
if world > 1:
    dist.destroy_process_group()
