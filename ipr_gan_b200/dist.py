"""Data parallelism for the protected step: one process per GPU, ``torch.distributed`` over NCCL/NVLink.

The reference scales with ``nn.DataParallel`` (experiments/base.py:24-39, models/dcgan.py:16-17): the global
batch is scattered in equal chunks, every replica keeps its own BatchNorm statistics, gradients are summed on
device 0 and parameters re-broadcast before every forward.  Here each rank owns ``torch.chunk(batch, world)[rank]``
(the same partition), BatchNorm stays per rank, and the only exchange is ONE all-reduce per network per optimizer
step on the flat gradient arena (mean over ranks == gradient of the global-batch mean loss); replicas stay
bit-identical because they apply identical reduced gradients, so no parameter broadcast is ever needed.
"""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard(batch, rank_=None, world_=None):
    """This rank's slice of a global batch: the partition nn.DataParallel's scatter makes."""
    r = rank() if rank_ is None else rank_
    w = world() if world_ is None else world_
    return torch.chunk(batch, w, dim=0)[r]


def allreduce_mean_(flat_tensor):
    """In-place mean over ranks of a flat buffer (gradient arena, or a vector of per-rank loss sums)."""
    w = world()
    if w > 1:
        dist.all_reduce(flat_tensor)
        flat_tensor.mul_(1.0 / w)
    return flat_tensor


def allreduce_sum_(flat_tensor):
    """In-place SUM over ranks (the 1/world factor is folded into the consumer: ``ipr_adam_flat_f32(grad_scale)``
    for gradients, the host-side division in ``get_metrics`` for the loss slots).  No-op for a single process."""
    if world() > 1:
        dist.all_reduce(flat_tensor)
    return flat_tensor


def reduce_metrics(metrics):
    """Average a metrics dict over ranks with one small all-reduce (per-rank means of equal-size shards)."""
    w = world()
    if w == 1:
        return metrics
    keys = sorted(metrics)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([metrics[k] for k in keys], dtype=torch.float64, device=dev)
    dist.all_reduce(t)
    return {k: float(v) / w for k, v in zip(keys, t.tolist())}
