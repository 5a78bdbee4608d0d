"""``horovod.torch`` as the reference uses it, over ``torch.distributed`` (NCCL on GPUs, gloo on CPU): one process per GPU
launched by ``torch.distributed.run`` instead of ``horovodrun`` (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the env).

Calls covered (every ``hvd.*`` under /src of the reference): init, size, rank, local_rank, allreduce_, broadcast_, allgather,
broadcast_parameters, broadcast_optimizer_state, Compression, DistributedOptimizer (+ synchronize / skip_synchronize).
``DistributedOptimizer.synchronize()`` all-reduces whole gradient *storages*: with clipbert_b200 models every ``.grad`` is
a view into one of two flat fp32 buffers, so the exchange is two collectives, not one per parameter
(src/tasks/run_video_retrieval.py:299-301,432).
"""
import contextlib
import os

import torch
import torch.distributed as dist


def init(backend=None):
    if dist.is_initialized():
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and "MASTER_ADDR" not in os.environ:
        return                                     # single process: every call below degenerates to a no-op
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank())
    dist.init_process_group(backend)


def _on():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def size():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else int(os.environ.get("WORLD_SIZE", "1"))


def rank():
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else int(os.environ.get("RANK", "0"))


def local_rank():
    return int(os.environ.get("LOCAL_RANK", "0"))


def allreduce_(tensor, average=True, name=None, op=None):
    """In-place all-reduce; like Horovod the default is the AVERAGE over ranks."""
    if not _on():
        return tensor
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    if average and op is None:
        tensor.div_(dist.get_world_size())
    return tensor


def broadcast_(tensor, root_rank, name=None):
    if _on():
        dist.broadcast(tensor, src=root_rank)
    return tensor


def allgather(tensor, name=None):
    """Concatenation along dim 0 of every rank's tensor; first dimensions may differ (src/utils/distributed.py:167)."""
    if not _on():
        return tensor
    world = dist.get_world_size()
    n = torch.tensor([tensor.shape[0]], dtype=torch.int64, device=tensor.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    pad = torch.zeros((max(sizes),) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    pad[: tensor.shape[0]] = tensor
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def broadcast_parameters(params, root_rank):
    """``params``: a state_dict or an iterable of (name, tensor) (run_video_retrieval.py:304)."""
    if not _on():
        return
    items = params.items() if isinstance(params, dict) else params
    for _, t in sorted(items, key=lambda kv: kv[0]):
        if torch.is_tensor(t):
            dist.broadcast(t.data if isinstance(t, torch.nn.Parameter) else t, src=root_rank)


def broadcast_optimizer_state(optimizer, root_rank):
    """Tensors of ``optimizer.state`` in place; scalar entries and param_group hyper-parameters through one object broadcast."""
    if not _on():
        return
    opt = getattr(optimizer, "_opt", optimizer)
    scalars = None
    if dist.get_rank() == root_rank:
        scalars = dict(groups=[{k: v for k, v in g.items() if k != "params"} for g in opt.param_groups],
                       state=[{k: v for k, v in opt.state.get(p, {}).items() if not torch.is_tensor(v)}
                              for g in opt.param_groups for p in g["params"]])
    box = [scalars]
    dist.broadcast_object_list(box, src=root_rank)
    scalars = box[0]
    for g, hyper in zip(opt.param_groups, scalars["groups"]):
        g.update(hyper)
    i = 0
    for g in opt.param_groups:
        for p in g["params"]:
            st = opt.state.get(p)
            if st is not None:
                st.update(scalars["state"][i])
                for v in st.values():
                    if torch.is_tensor(v):
                        dist.broadcast(v, src=root_rank)
            i += 1


class Compression:
    """Horovod's gradient compression selectors; gradients are exchanged as they are stored (fp32 flat buffers)."""
    none = "none"
    fp16 = "fp16"


def _grad_buffers(params):
    """Unique gradient storages as flat tensors: one entry per flat buffer when .grad tensors are views of a shared one."""
    seen, out = set(), []
    for p in params:
        g = p.grad
        if g is None:
            continue
        st = g.untyped_storage()
        key = (st.data_ptr(), g.dtype)
        if key in seen:
            continue
        seen.add(key)
        out.append(torch.empty(0, dtype=g.dtype, device=g.device).set_(st))
    return out


class _DistributedOptimizer:
    """Wraps any torch optimizer: ``step()`` exchanges the gradients first unless ``synchronize()`` already did
    (the reference calls ``synchronize()`` explicitly, then ``step()`` under ``skip_synchronize()``, :432,492)."""

    def __init__(self, optimizer, named_parameters=None, compression=Compression.none, **_kw):
        self._opt = optimizer
        self._skip = False
        self._synchronized = False

    def __getattr__(self, name):
        return getattr(self._opt, name)

    def _params(self):
        for g in self._opt.param_groups:
            for p in g["params"]:
                yield p

    def synchronize(self):
        if _on():
            world = dist.get_world_size()
            works = []
            bufs = _grad_buffers(self._params())
            for b in bufs:
                if dist.get_backend() == "nccl":
                    works.append(dist.all_reduce(b, op=dist.ReduceOp.AVG, async_op=True))
                else:
                    b.div_(world)
                    works.append(dist.all_reduce(b, op=dist.ReduceOp.SUM, async_op=True))
            for w in works:
                w.wait()
        self._synchronized = True

    @contextlib.contextmanager
    def skip_synchronize(self):
        self._skip = True
        try:
            yield
        finally:
            self._skip = False

    def step(self, *a, **k):
        if not self._skip and not self._synchronized:
            self.synchronize()
        self._synchronized = False
        return self._opt.step(*a, **k)

    def zero_grad(self, *a, **k):
        return self._opt.zero_grad(*a, **k)


def DistributedOptimizer(optimizer, named_parameters=None, compression=Compression.none, **kw):
    return _DistributedOptimizer(optimizer, named_parameters, compression, **kw)
