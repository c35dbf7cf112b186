"""Multi-GPU host logic of the backbone path: one process per GPU.

* Inference shards by BUILDING: every forward builds its own Metadata (sparseconvnet/ioLayers.py:52-55),
  buildings share nothing, so ranks never exchange data on the data path; the per-building results
  are collected on the host at the end (`gather_results`).
* Training is data parallel: the only collective is one all-reduce of the gradients
  (tools/train_net_sparse3d.py:52-58 wraps the model in DistributedDataParallel with
  broadcast_buffers=False; BatchNorm statistics stay per GPU).  Parameters of the dead top-down
  levels never receive a gradient; they are reduced as zeros so that every rank
  reduces the same flat layout.

Backend-agnostic: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_buildings(sizes, world_size, rank, policy="lpt"):
    """Indices of the buildings rank `rank` processes.

    sizes: per-building cost proxy (active voxels / input rows).  policy 'lpt' = longest processing
    time first onto the least loaded rank (ties -> lowest rank), 'round_robin' = i % world_size.
    Deterministic and identical on every rank; the shards are disjoint and cover every building."""
    n = len(sizes)
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    if policy == "round_robin":
        return [i for i in range(n) if i % world_size == rank]
    if policy != "lpt":
        raise ValueError("policy must be 'lpt' or 'round_robin'")
    order = sorted(range(n), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world_size
    mine = []
    for i in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        load[r] += int(sizes[i])
        if r == rank:
            mine.append(i)
    return sorted(mine)


def gather_results(local, dst=0, group=None):
    """local: {building index: picklable result} of this rank.  Returns the merged dict on `dst`
    (None elsewhere).  Host-side only: nothing on the GPU data path."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(local)
    world = dist.get_world_size(group)
    bucket = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object(dict(local), bucket, dst=dst, group=group)
    if bucket is None:
        return None
    merged = {}
    for part in bucket:
        dup = set(merged) & set(part)
        if dup:
            raise RuntimeError(f"buildings processed twice: {sorted(dup)}")
        merged.update(part)
    return merged


def allreduce_gradients(params, group=None, bucket_bytes=64 << 20, average=True):
    """Sum (or average) the gradients of `params` over all ranks with as few collectives as possible:
    gradients are packed into flat fp32 buckets of <= bucket_bytes, all-reduced, and copied back.
    A parameter whose .grad is None on this rank contributes zeros and receives the reduced value,
    so ranks whose buildings leave different levels untouched still agree on the layout.
    Returns the number of collectives issued."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return 0
    dev = params[0].device
    buckets, cur, cur_bytes = [], [], 0
    for p in params:
        nb = p.numel() * 4
        if cur and cur_bytes + nb > bucket_bytes:
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(p)
        cur_bytes += nb
    if cur:
        buckets.append(cur)
    for bucket in buckets:
        flat = torch.zeros(sum(p.numel() for p in bucket), dtype=torch.float32, device=dev)
        off = 0
        for p in bucket:
            if p.grad is not None:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
            off += p.numel()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
        off = 0
        for p in bucket:
            g = flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += p.numel()
    return len(buckets)


class GradientReducer(object):
    """Data-parallel gradient all-reduce overlapped with the backward pass ("bucketed to overlap with
    backward"; the reference wraps the model in DistributedDataParallel, tools/train_net_sparse3d.py:52-58).

    * ONE persistent flat fp32 buffer holds every gradient; each `p.grad` is a view into it (autograd accumulates in place), so
      there is no packing or unpacking and no allocation per step -- `zero()` is one memset.
    * The buffer is cut into buckets in REVERSE parameter order (the order in which the backward pass finishes them).  A bucket
      is all-reduced asynchronously as soon as every LIVE parameter in it has received its gradient (post-accumulate hooks);
      the collective runs on NCCL's own stream while the backward kernels of the earlier layers keep the SMs busy.
    * Parameters that never receive a gradient (the dead top-down levels, App. D.2) are learnt in the first step (`calibrate`):
      they stay zero in the buffer and never hold a bucket back; every rank reduces the same flat layout regardless.
    `finish()` launches what is left, waits for all collectives and divides by the world size."""

    def __init__(self, params, bucket_bytes=24 << 20, group=None, average=True):
        self.params = [p for p in params if p.requires_grad]
        self.group, self.average = group, average
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        order = list(reversed(self.params))
        self.slices, self.bucket_of, self.buckets = {}, {}, []   # param -> (lo, hi); param -> bucket index; bucket -> (lo, hi)
        off, cur_lo, cur_bytes = 0, 0, 0
        for p in order:
            n = p.numel()
            if cur_bytes and cur_bytes + n * 4 > bucket_bytes:
                self.buckets.append((cur_lo, off))
                cur_lo, cur_bytes = off, 0
            self.slices[p] = (off, off + n)
            self.bucket_of[p] = len(self.buckets)
            off += n
            cur_bytes += n * 4
        self.buckets.append((cur_lo, off))
        self.live = None            # set of parameters that receive gradients (after the first step)
        self.pending, self.seen, self.handles, self.launched = [], set(), [], set()
        for p in self.params:
            lo, hi = self.slices[p]
            p.grad = self.flat[lo:hi].view_as(p)
            p.register_post_accumulate_grad_hook(self._hook)

    def attach(self, net):
        """Replayed training steps (FPN_Net's recorded training program, one autograd node per step): the native backward pass writes
        every parameter gradient straight into this reducer's flat buffer and records an event behind each; when the pass has been
        QUEUED, every bucket's all-reduce is launched on a side stream that waits for the events of that bucket's parameters only, so
        the collectives overlap the rest of the backward pass as they do with per-layer hooks.  Call once the program exists (after
        the first training step); returns False while it does not."""
        prog = net.__dict__.get("_program_train")
        if prog is None or not self.flat.is_cuda:
            return False
        sink, events, self._prog_bucket = {}, {}, [[] for _ in self.buckets]
        for i, p in enumerate(prog.params):
            if p in self.slices:
                lo, hi = self.slices[p]
                sink[i] = self.flat[lo:hi].view_as(p)
                ev = torch.cuda.Event()
                ev.record()  # (creates the native event; re-recorded by scn_program_backward every step)
                events[i] = ev
                self._prog_bucket[self.bucket_of[p]].append(i)
        prog.grad_sink, prog.grad_events, prog.after_backward = sink, events, self._after_replayed_backward
        self._side = torch.cuda.Stream()
        self._zero_ev = torch.cuda.Event()
        return True

    def _after_replayed_backward(self, prog):
        if self.world == 1:
            return
        for b in range(len(self.buckets)):  # (reverse parameter order = the order in which the backward pass completes them)
            with torch.cuda.stream(self._side):
                self._side.wait_event(self._zero_ev)  # buckets without a live parameter are all zeros: only the memset must be done
                for i in self._prog_bucket[b]:
                    if prog.last_live[i]:
                        self._side.wait_event(prog.grad_events[i])
                self._launch(b)

    def zero(self):
        """Start of a step: gradients to zero (one memset), bucket bookkeeping reset."""
        self.flat.zero_()
        if getattr(self, "_zero_ev", None) is not None:
            self._zero_ev.record()
        for p in self.params:  # (an optimizer / zero_grad(set_to_none=True) may have dropped the views)
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + self.slices[p][0] * 4:
                lo, hi = self.slices[p]
                p.grad = self.flat[lo:hi].view_as(p)
        self.seen, self.handles, self.launched = set(), [], set()
        if self.live is not None:
            self.pending = [sum(1 for p in self.params if self.bucket_of[p] == b and p in self.live) for b in range(len(self.buckets))]

    def _launch(self, b):
        if b in self.launched or self.world == 1:
            self.launched.add(b)
            return
        lo, hi = self.buckets[b]
        self.handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.launched.add(b)

    def _hook(self, p):
        self.seen.add(p)
        if self.live is None or p not in self.live:
            return
        b = self.bucket_of[p]
        self.pending[b] -= 1
        if self.pending[b] == 0:
            self._launch(b)

    def finish(self):
        """After backward: launch the buckets that did not complete on their own (first step, buckets without live
        parameters), wait, average.  Returns the number of collectives of this step."""
        if self.live is None:
            self.live = set(self.seen)
        for b in range(len(self.buckets)):
            self._launch(b)
        for h in self.handles:
            h.wait()
        n = len(self.handles)
        if self.average and self.world > 1:
            self.flat.div_(self.world)
        return n

    def nbytes(self):
        return self.flat.numel() * 4


def max_over_ranks_ms(ms, device, group=None):
    """Device-timed milliseconds -> max over ranks (how every multi-GPU number of bench.py is taken)."""
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
