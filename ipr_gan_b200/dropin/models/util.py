"""Helpers used by the step logic.

``DisableBatchNormStats`` (models/util.py:55-69): the trigger forward uses batch statistics but must
not touch running statistics or ``num_batches_tracked``.
``Replica``: the one-process-per-GPU stand-in for ``torch.nn.DataParallel`` -- same ``.module``
attribute and ``module.``-prefixed state_dict keys (checkpoint format), inputs staged to the module's
device; data parallelism itself is NCCL all-reduce across processes (ipr_gan_b200.dist).  Asking ONE process for
several devices (``resource.ngpu > 1``, experiments/base.py:24-39) raises: the reference would scatter the enlarged
batch over its replicas with per-replica BatchNorm, which a single-device run does not reproduce."""
import torch
import torch.nn as nn


class DisableBatchNormStats(object):
    def __init__(self, model):
        self.layers = [m for m in model.modules() if isinstance(m, nn.BatchNorm2d)]
        self.saved = None

    def __enter__(self):
        self.saved = [m.track_running_stats for m in self.layers]
        for m in self.layers:
            m.track_running_stats = False

    def __exit__(self, *exc):
        for m, flag in zip(self.layers, self.saved):
            m.track_running_stats = flag


class Replica(nn.Module):
    def __init__(self, module, device_ids=None):
        super().__init__()
        self.module = module
        self.device_ids = [d for d in (device_ids or []) if d is not None]
        if len(self.device_ids) > 1:
            raise RuntimeError(
                "ipr_gan_b200 runs one process per GPU: got device_ids=%s in a single process.  Keep resource.ngpu: 1 "
                "and launch `python -m torch.distributed.run --nproc-per-node %d train.py ...` -- every rank then "
                "trains its torch.chunk shard of the batch (nn.DataParallel's partition, per-rank BatchNorm) and "
                "gradients are all-reduced over NCCL inside optimizer.step()." % (self.device_ids, len(self.device_ids)))

    def _device(self):
        for t in list(self.module.parameters()) + list(self.module.buffers()):
            return t.device
        return None

    def forward(self, *inputs):
        dev = self._device()
        if dev is not None:
            inputs = tuple(t.to(dev, non_blocking=True) if isinstance(t, torch.Tensor) else t for t in inputs)
        return self.module(*inputs)
