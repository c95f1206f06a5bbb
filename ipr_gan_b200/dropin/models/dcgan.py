"""Hinge-loss DCGAN step with the reference's interface (models/dcgan.py:7-78): ``update_d(data)`` with
``real_sample`` / ``latent``, ``update_g(data, update=True)`` re-using the generator graph built in
``forward_d``, ``get_metrics()`` -> dict of floats, ``_modules`` = G, D, optG, optD."""
import torch
from torch import optim
from torch.nn import functional as F

import networks
from models.core import Model
from models.util import Replica


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
        if config.opt == "Adam" and device[0].type == "cuda":
            # same class hierarchy and state_dict format as torch.optim.Adam; one flat-arena launch per step,
            # gradient all-reduce over NCCL when torch.distributed is initialised
            from ipr_gan_b200.optim import FlatAdam as make_opt
        self.optG = make_opt(self.G.parameters(), **kwargs)
        self.optD = make_opt(self.D.parameters(), **kwargs)
        self._modules.update(G=self.G, D=self.D, optG=self.optG, optD=self.optD)

    # ---- losses (models/dcgan.py:31-40)
    def compute_d_loss(self):
        self.LossR = F.relu(1.0 - self.real_logits).mean()
        self.LossF = F.relu(1.0 + self.fake_logits).mean()
        self.LossD = self.LossR + self.LossF

    def compute_g_loss(self):
        self.LossA = -self.gen_logits.mean()
        self.LossG = self.LossA

    # ---- forwards (models/dcgan.py:42-52)
    def forward_d(self, data):
        self.latent = data["latent"]
        self.real_sample = data["real_sample"]
        dev = self.device[0]
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
        self.fake_logits = self.D(fake)
        col = getattr(fake, "_ipr_col", None)       # patch matrix of the image: update_g's D(generated) reuses it
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
        vals = torch.stack([self.LossD, self.LossR, self.LossF, self.LossG, self.LossA]).tolist()  # one D2H copy
        return dict(zip(("D/Sum", "D/Real", "D/Fake", "G/Sum", "G/Adv"), vals))

    # ---- updates (models/dcgan.py:63-78)
    def update_d(self, data):
        self.forward_d(data)
        self.compute_d_loss()
        self.optD.zero_grad()
        self.LossD.backward()
        self.optD.step()

    def update_g(self, data, update=True):
        self.forward_g(data)
        self.compute_g_loss()
        if update:
            self.optG.zero_grad()
            self.LossG.backward()
            self.optG.step()
