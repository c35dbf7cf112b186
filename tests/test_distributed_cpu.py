"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: building sharding without a collective,
host-side result gathering, the data-parallel gradient all-reduce."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from detection_3d_b200 import distributed as D  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shards_are_disjoint_cover_everything_and_balance():
    sizes = [1155656, 900000, 1300000, 40000, 700000, 1100000, 950000, 20000, 1250000, 600000, 1000000]
    for world in (1, 2, 4, 8):
        for policy in ("lpt", "round_robin"):
            shards = [D.shard_buildings(sizes, world, r, policy) for r in range(world)]
            flat = sorted(i for s in shards for i in s)
            assert flat == list(range(len(sizes))), (world, policy)
        loads = [sum(sizes[i] for i in D.shard_buildings(sizes, world, r)) for r in range(world)]
        if world > 1:  # LPT bound: max load <= mean + largest item
            assert max(loads) <= sum(sizes) / world + max(sizes)
            rr = [sum(sizes[i] for i in D.shard_buildings(sizes, world, r, "round_robin")) for r in range(world)]
            assert max(loads) <= max(rr)
    assert D.shard_buildings([], 2, 1) == []
    with pytest.raises(ValueError):
        D.shard_buildings(sizes, 2, 2)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- inference: each rank "processes" its shard; only the host-side gather communicates
        sizes = [50, 10, 40, 30, 20, 60, 5]
        mine = D.shard_buildings(sizes, world, rank)
        local = {i: {"rows": sizes[i] * 2, "rank": rank} for i in mine}
        merged = D.gather_results(local, dst=0)
        if rank == 0:
            assert sorted(merged) == list(range(len(sizes)))
            assert all(merged[i]["rows"] == sizes[i] * 2 for i in merged)
        else:
            assert merged is None
        # ---- training: gradient all-reduce; parameter 1 has no gradient on rank 1, parameter 2 on no rank
        torch.manual_seed(0)
        ps = [torch.nn.Parameter(torch.zeros(7, 3)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2)),
              torch.nn.Parameter(torch.zeros(1000))]
        ps[0].grad = torch.full((7, 3), float(rank + 1))
        if rank == 0:
            ps[1].grad = torch.arange(5, dtype=torch.float32)
        ps[3].grad = torch.full((1000,), 2.0 * rank)
        n = D.allreduce_gradients(ps, bucket_bytes=2048)  # forces several buckets
        assert n >= 2
        assert torch.allclose(ps[0].grad, torch.full((7, 3), (1 + 2) / 2.0))
        assert torch.allclose(ps[1].grad, torch.arange(5, dtype=torch.float32) / 2.0)  # zeros from rank 1
        assert ps[2].grad is not None and torch.count_nonzero(ps[2].grad) == 0
        assert torch.allclose(ps[3].grad, torch.full((1000,), 1.0))
        # ---- overlapped reducer: gradients are views of one flat buffer, buckets fire from hooks, dead parameters stay zero
        torch.manual_seed(1)
        lin = [torch.nn.Linear(6, 6) for _ in range(4)]
        dead = torch.nn.Parameter(torch.ones(50))  # never used: no gradient on any rank
        params = [q for l in lin for q in l.parameters()] + [dead]
        for q in params:
            dist.broadcast(q.data, 0)
        red = D.GradientReducer(params, bucket_bytes=200)
        assert len(red.buckets) >= 3
        for step in range(3):
            red.zero()
            x = torch.full((2, 6), float(rank + 1 + step))
            y = x
            for l in lin:
                y = torch.tanh(l(y))
            y.sum().backward()
            n_coll = red.finish()
            assert n_coll == len(red.buckets)
            # reference: the same computation for both ranks' inputs, averaged
            want = []
            for r in range(world):
                ps2 = [q.detach().clone().requires_grad_(True) for q in params[:-1]]
                yy = torch.full((2, 6), float(r + 1 + step))
                for k in range(4):
                    yy = torch.tanh(torch.nn.functional.linear(yy, ps2[2 * k], ps2[2 * k + 1]))
                want.append(torch.autograd.grad(yy.sum(), ps2))
            for k, q in enumerate(params[:-1]):
                assert q.grad.data_ptr() == red.flat.data_ptr() + red.slices[q][0] * 4  # still a view of the flat buffer
                assert torch.allclose(q.grad, (want[0][k] + want[1][k]) / 2, atol=1e-6), (step, k)
            assert torch.count_nonzero(dead.grad) == 0
            if step > 0:
                assert dead not in red.live and len(red.live) == 8
        # ---- timing helper: max over ranks
        assert D.max_over_ranks_ms(10.0 + rank, torch.device("cpu")) == 10.0 + world - 1
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_gather_and_gradient_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(ret) == {0: "ok", 1: "ok"}


def test_single_process_paths_need_no_process_group():
    assert D.gather_results({3: "x"}) == {3: "x"}
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    assert D.allreduce_gradients([p]) == 0 and torch.equal(p.grad, torch.ones(3))


def test_gradient_reducer_attach_is_a_no_op_without_a_recorded_program():
    """GradientReducer.attach(net) before the first training step (no recorded program) or on CPU tensors returns False and leaves the
    hook-driven path in charge."""
    import torch
    from detection_3d_b200 import distributed

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(4, 3))

    net = Net()
    red = distributed.GradientReducer(list(net.parameters()))
    assert red.attach(net) is False
    red.zero()
    (net.w * 2).sum().backward()
    assert red.finish() == 0
    assert torch.equal(net.w.grad, torch.full((4, 3), 2.0)) and net.w.grad.data_ptr() == red.flat.data_ptr()


def test_truncate_lazy_cuts_padded_proposal_lists_with_one_read():
    """postproc.truncate_lazy: the padded keep lists of several class groups are cut to their device-side counts (host logic, CPU tensors)."""
    import torch
    from detection_3d_b200 import postproc
    a = postproc.Boxes3D(torch.arange(35, dtype=torch.float32).view(5, 7))
    a.add_field("objectness", torch.arange(5, dtype=torch.float32))
    b = postproc.Boxes3D(torch.arange(21, dtype=torch.float32).view(3, 7))
    b.add_field("objectness", torch.arange(3, dtype=torch.float32))
    out = postproc.truncate_lazy([(a, torch.tensor([2])), (b, torch.tensor([3]))])
    assert len(out[0]) == 2 and torch.equal(out[0].bbox3d, a.bbox3d[:2]) and torch.equal(out[0].get_field("objectness"), torch.tensor([0.0, 1.0]))
    assert len(out[1]) == 3 and out[1] is b
    assert postproc.truncate_lazy([]) == []
