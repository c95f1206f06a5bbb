"""SRGAN step with the reference's interface (models/srgan.py:7-106): generator step first (MSE pre-training or
VGG content + 1e-3 adversarial BCE), then the discriminator step on the same tensors; ``_modules`` = G, D, optG, optD."""
import torch
from torch import optim
from torch.nn import functional as F

import networks
from models.core import Model
from models.util import Replica


def _make_opt(config, params, device):
    kwargs = config.opt_param.to_dict()
    if config.opt == "Adam" and device.type == "cuda":
        from ipr_gan_b200.optim import FlatAdam
        return FlatAdam(params, **kwargs)
    return getattr(optim, config.opt)(params, **kwargs)


class SRGAN(Model):
    def __init__(self, config, device=[torch.device("cpu"), ]):
        super().__init__()
        self.device = device
        ids = [d.index for d in device]
        self.G = Replica(getattr(networks, config.G)().to(device[0]), device_ids=ids)
        self.D = Replica(getattr(networks, config.D)().to(device[0]), device_ids=ids)
        self.V = Replica(getattr(networks, config.V)().to(device[0]), device_ids=ids)
        self.G.train()
        self.D.train()
        self.V.eval()
        self.optG = _make_opt(config, list(self.G.parameters()), device[0])
        self.optD = _make_opt(config, list(self.D.parameters()), device[0])
        self._modules.update(G=self.G, D=self.D, optG=self.optG, optD=self.optD)

    def compute_d_loss(self):
        self.LossR = F.binary_cross_entropy_with_logits(self.real_logits, torch.ones_like(self.real_logits))
        self.LossF = F.binary_cross_entropy_with_logits(self.fake_logits, torch.zeros_like(self.fake_logits))
        self.LossD = self.LossR + self.LossF

    def compute_g_loss(self):
        dev = self.super_res.device
        if self.pretrain:
            self.LossG = F.mse_loss(self.super_res, self.high_res.to(dev))
            return
        self.LossA = F.binary_cross_entropy_with_logits(self.gen_logits, torch.ones_like(self.gen_logits))
        self.LossX = F.mse_loss(self.V(self.super_res), self.V(self.high_res).detach())
        self.LossG = self.LossX + 1e-3 * self.LossA

    def forward_d(self, data):
        self.high_res = data["high_res"]
        self.super_res = data["super_res"]
        self.real_logits = self.D(self.high_res)
        self.fake_logits = self.D(self.super_res.detach())

    def forward_g(self, data):
        self.low_res = data["low_res"]
        self.high_res = data["high_res"]
        self.pretrain = data["pretrain"]
        self.super_res = self.G(self.low_res)
        if not self.pretrain:
            self.gen_logits = self.D(self.super_res)

    def get_metrics(self):
        if self.pretrain:
            g = self.LossG.item()
            return {"G/MSE": g, "G/Sum": g}
        vals = torch.stack([self.LossD, self.LossR, self.LossF, self.LossG, self.LossA, self.LossX]).tolist()
        return dict(zip(("D/Sum", "D/Real", "D/Fake", "G/Sum", "G/Adv", "G/Con"), vals))

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
