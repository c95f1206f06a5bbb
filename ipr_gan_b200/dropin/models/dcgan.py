"""Hinge-loss DCGAN step with the reference's interface (models/dcgan.py:7-78): ``update_d(data)`` with
``real_sample`` / ``latent``, ``update_g(data, update=True)`` re-using the generator graph built in
``forward_d``, ``get_metrics()`` -> dict of floats, ``_modules`` = G, D, optG, optD.

On CUDA with the native networks and Adam the step holds no framework arithmetic: the hinge / adversarial losses
and their logit gradients are one launch each (csrc/step_misc.cu), the gradients go straight into the networks'
backward passes, the loss scalars land in an 8-float *board* that lives between the two networks' gradient arenas
(so the data-parallel all-reduce of the gradients carries them along) and ``get_metrics()`` is one 32-byte copy."""
import torch
from torch import optim
from torch.nn import functional as F

import networks
from models.core import Model
from models.util import Replica


class MetricsBoard(object):
    """Eight fp32 loss slots on the device + their pinned host mirror.
    Layout  [0] D/Sum  [1] D/Real  [2] D/Fake  [3] -  |  [4] G/Adv  [5] watermark loss  [6] sign loss  [7] -
    Slots 0-3 ride behind D's gradients, 4-7 in front of G's (``flat.share_gradient_buffer``)."""
    D_SUM, D_REAL, D_FAKE, G_ADV, P_WM, P_SIGN = 0, 1, 2, 4, 5, 6

    def __init__(self, arena_d, arena_g):
        from ipr_gan_b200 import dist, flat
        self.dev = flat.share_gradient_buffer(arena_d, arena_g)
        self.host = torch.zeros(8).pin_memory()
        self.event = torch.cuda.Event()
        self.values = [0.0] * 8
        self.dirty = True
        self.loss_scale = 1.0 / dist.world()      # per-rank share; the summing all-reduce makes the global mean

    def slot(self, i, n=1):
        return self.dev[i:i + n]

    def scalar(self, i):
        return self.dev[i]                        # 0-dim view: what LossD / LossG ... are on this path

    def touch(self):
        self.dirty = True

    def fetch(self):
        """One asynchronous 32-byte copy + one event wait per step, however many metrics are read."""
        if self.dirty:
            self.host.copy_(self.dev, non_blocking=True)
            self.event.record()
            self.event.synchronize()
            self.values = self.host.tolist()
            self.dirty = False
        return self.values


class DCGAN(Model):
    def __init__(self, config, device=[torch.device("cpu"), ]):
        super().__init__()
        self.device = device
        ids = [d.index for d in device]
        self.G = Replica(getattr(networks, config.G)().to(device[0]), device_ids=ids)
        self.D = Replica(getattr(networks, config.D)().to(device[0]), device_ids=ids)
        self.G.train()
        self.D.train()

        make_opt = getattr(optim, config.opt)
        kwargs = config.opt_param.to_dict()
        flat_adam = config.opt == "Adam" and device[0].type == "cuda"
        if flat_adam:
            # same class hierarchy and state_dict format as torch.optim.Adam; one flat-arena launch per step,
            # gradient all-reduce over NCCL when torch.distributed is initialised
            from ipr_gan_b200.optim import FlatAdam as make_opt
        self.optG = make_opt(self.G.parameters(), **kwargs)
        self.optD = make_opt(self.D.parameters(), **kwargs)
        self._modules.update(G=self.G, D=self.D, optG=self.optG, optD=self.optD)
        # fused step: native networks + flat Adam (anything else runs the same sequence on autograd tensors)
        self.board = None
        if flat_adam and isinstance(self.G.module, networks.ConvGenerator) and \
                isinstance(self.D.module, networks.SNDiscriminator):
            self.board = MetricsBoard(self.optD.arena, self.optG.arena)
            self.optD.arena.track_clean = self.optG.arena.track_clean = True
        self.seeded = self.board is not None
        self.g_seeds = []                # (tensor, gradient) pairs the generator step back-propagates from

    # ---- losses (models/dcgan.py:31-40)
    def compute_d_loss(self):
        if self.board is not None:
            from ipr_gan_b200 import ops
            b = self.board
            self._d_seeds = ops.hinge_d_loss(self.real_logits, self.fake_logits, b.slot(b.D_SUM, 3), b.loss_scale)
            self.LossD, self.LossR, self.LossF = b.scalar(b.D_SUM), b.scalar(b.D_REAL), b.scalar(b.D_FAKE)
            return
        self.LossR = F.relu(1.0 - self.real_logits).mean()
        self.LossF = F.relu(1.0 + self.fake_logits).mean()
        self.LossD = self.LossR + self.LossF

    def compute_g_loss(self):
        if self.board is not None:
            from ipr_gan_b200 import ops
            b = self.board
            d = ops.gen_adv_loss(self.gen_logits, b.slot(b.G_ADV), b.loss_scale)
            self.g_seeds = [(self.gen_logits, d)]
            self.LossA = self.LossG = b.scalar(b.G_ADV)
            return
        self.LossA = -self.gen_logits.mean()
        self.LossG = self.LossA

    # ---- forwards (models/dcgan.py:42-52)
    def forward_d(self, data):
        self.latent = data["latent"]
        self.real_sample = data["real_sample"]
        dev = self.device[0]
        if self.board is not None:
            # staged once: G(latent) here and the trigger input fn_inp(latent) in the generator step share the copy
            self.latent = self.latent.to(dev, torch.float32, non_blocking=True)
        if dev.type == "cuda" and self._concurrent():
            # D(real) does not depend on the generator: it runs (forward and, through autograd, backward) on a second
            # stream next to G(z) -> D(fake).  The engine keeps the reference's order for everything the two passes
            # share: spectral-norm power iterations (real first, then fake), weight packing, gradient accumulation.
            from ipr_gan_b200 import engine
            main, aux = torch.cuda.current_stream(dev), engine.aux_stream(dev)
            aux.wait_stream(main)
            with torch.cuda.stream(aux):
                self.real_logits = self.D(self.real_sample)
            self.fake_sample = self.G(self.latent)
            self._d_on_fake()
            main.wait_stream(aux)
            self.real_logits.record_stream(main)
            return
        self.fake_sample = self.G(self.latent)
        self.real_logits = self.D(self.real_sample)
        self._d_on_fake()

    def _d_on_fake(self):
        fake = self.fake_sample.detach()
        self.D.module._ipr_keep_col = True          # patch matrix of the image: update_g's D(generated) reuses it
        try:
            self.fake_logits = self.D(fake)
        finally:
            self.D.module._ipr_keep_col = False
        col = getattr(fake, "_ipr_col", None)
        if col is not None:
            self.fake_sample._ipr_col = col

    def _concurrent(self):
        from ipr_gan_b200 import engine
        return engine.concurrent_passes()

    def forward_g(self, data):
        self.generated = data["fake_sample"]
        # the generator step needs dD/dx only; the reference also fills D's .grad here but zeroes it before
        # its next use (models/dcgan.py:67), so the weight-gradient GEMMs are skipped
        self.D.module._ipr_skip_param_grads = True
        try:
            self.gen_logits = self.D(self.generated)
        finally:
            self.D.module._ipr_skip_param_grads = False

    def get_metrics(self):
        if self.board is not None:
            v = self.board.fetch()
            b = self.board
            return {"D/Sum": v[b.D_SUM], "D/Real": v[b.D_REAL], "D/Fake": v[b.D_FAKE], "G/Sum": v[b.G_ADV],
                    "G/Adv": v[b.G_ADV]}
        vals = torch.stack([self.LossD, self.LossR, self.LossF, self.LossG, self.LossA]).tolist()  # one D2H copy
        return dict(zip(("D/Sum", "D/Real", "D/Fake", "G/Sum", "G/Adv"), vals))

    # ---- updates (models/dcgan.py:63-78)
    def update_d(self, data):
        self.forward_d(data)
        self.compute_d_loss()
        self.optD.zero_grad()
        if self.board is not None:
            self.board.touch()
            torch.autograd.backward([self.real_logits, self.fake_logits], list(self._d_seeds))
            # everything the trigger pass of the coming generator step needs (G(z), the latents) exists by now: it may
            # start here, next to D's gradient all-reduce / Adam update / weight re-packing (models/protect.py)
            self.pre_step_event = torch.cuda.Event()
            self.pre_step_event.record(torch.cuda.current_stream(self.device[0]))
        else:
            self.LossD.backward()
        self.optD.step()

    def backward_g(self, extra=None):
        """Back-propagate the generator step: from the collected (tensor, gradient) seeds on the fused path, from the
        summed loss tensor otherwise.  The wrappers add their seeds / terms and call this once."""
        self.optG.zero_grad()
        if self.board is not None:
            self.board.touch()
            seeds = self.g_seeds
            torch.autograd.backward([t for t, _ in seeds], [g for _, g in seeds])
        else:
            (self.LossG if extra is None else extra).backward()

    def update_g(self, data, update=True):
        self.forward_g(data)
        self.compute_g_loss()
        if update:
            self.backward_g()
            self.optG.step()
