"""SRGAN step behind the reference's interface (models/srgan.py:7-106): ``update_g`` first (MSE pre-training, or VGG
content loss + 1e-3 x adversarial BCE), then ``update_d`` on the same tensors; ``_modules`` = G, D, optG, optD.

Native step: SRResNet and Discriminator96 run on the library's engine (``ipr_gan_b200.seqnet``), every loss is ONE
launch producing the value and its gradient (``ops.pointwise_loss``), and the backward pass starts from those seed
gradients (``models.core.Seeds``) -- there is no autograd graph of loss arithmetic.  The frozen VGG feature extractor
is outside the accelerated path and stays a PyTorch module between ``super_res`` and the content-loss seed."""
import torch
from torch import optim

import networks
from models.core import Model, Seeds
from models.util import Replica

ADV_WEIGHT = 1e-3                        # models/srgan.py:60


def make_optimizer(config, params, device):
    kwargs = config.opt_param.to_dict()
    if config.opt == "Adam" and device.type == "cuda":
        from ipr_gan_b200.optim import FlatAdam
        return FlatAdam(params, **kwargs)
    return getattr(optim, config.opt)(params, **kwargs)


class SRGAN(Model):
    seeded = True                        # wrappers add (tensor, gradient) seeds instead of loss-graph terms

    def __init__(self, config, device=[torch.device("cpu"), ]):
        super().__init__()
        self.device = device
        ids = [d.index for d in device]
        for name, key in (("G", config.G), ("D", config.D), ("V", config.V)):
            setattr(self, name, Replica(getattr(networks, key)().to(device[0]), device_ids=ids))
        self.G.train(), self.D.train(), self.V.eval()
        self.optG = make_optimizer(config, self.G.parameters(), device[0])
        self.optD = make_optimizer(config, self.D.parameters(), device[0])
        self._modules.update(G=self.G, D=self.D, optG=self.optG, optD=self.optD)
        self.g_seeds = Seeds()
        self._scalars = {}               # metric name -> (0-dim device tensor, factor applied on the host)

    def _loss(self, name, kind, x, target, weight=1.0, report=1.0):
        from ipr_gan_b200 import ops
        value, grad = ops.pointwise_loss(kind, x, target, weight)
        self._scalars[name] = (value, report)
        return value, grad

    # ---- generator step (models/srgan.py:47-61, 69-77)
    def forward_g(self, data):
        self.low_res, self.high_res, self.pretrain = data["low_res"], data["high_res"], data["pretrain"]
        self._hr_feat = None
        dev = self.device[0]
        if not self.pretrain and dev.type == "cuda":
            # the content target V(high_res) depends on nothing the generator does: it runs on the aux stream next to
            # G(low_res) and D(super_res) (the frozen VGG is outside the accelerated path, its ~1 ms still counts)
            from ipr_gan_b200 import engine
            if engine.concurrent_passes():
                main, aux = torch.cuda.current_stream(dev), engine.aux_stream(dev)
                hr = self.high_res.to(dev, non_blocking=True)
                aux.wait_stream(main)
                with torch.cuda.stream(aux), torch.no_grad():
                    self._hr_feat = (self.V(hr), aux)
                hr.record_stream(aux)
        self.super_res = self.G(self.low_res)
        if not self.pretrain:
            self.D.module._ipr_skip_param_grads = True     # only dD/d(super_res) is used; D's .grad is zeroed before its step
            try:
                self.gen_logits = self.D(self.super_res)
            finally:
                self.D.module._ipr_skip_param_grads = False

    def compute_g_loss(self):
        dev = self.super_res.device
        hr = self.high_res.to(dev, non_blocking=True)
        self.g_seeds = Seeds()
        self._scalars = {k: v for k, v in self._scalars.items() if k.startswith("D/")}
        if self.pretrain:
            self.LossG, d = self._loss("G/MSE", "mse", self.super_res, hr)
            self.g_seeds.add(self.super_res, d)
            return
        adv, d_adv = self._loss("G/Adv", "bce_logits", self.gen_logits, 1.0, weight=ADV_WEIGHT, report=1.0 / ADV_WEIGHT)
        sr_feat = self.V(self.super_res)
        if self._hr_feat is not None:
            hr_feat, aux = self._hr_feat
            torch.cuda.current_stream(dev).wait_stream(aux)
            hr_feat.record_stream(torch.cuda.current_stream(dev))
            self._hr_feat = None
        else:
            with torch.no_grad():
                hr_feat = self.V(hr)
        self.LossX, d_feat = self._loss("G/Con", "mse", sr_feat, hr_feat)
        self.LossA = adv / ADV_WEIGHT
        self.LossG = self.LossX + adv
        self.g_seeds.add(self.gen_logits, d_adv)
        self.g_seeds.add(sr_feat, d_feat)

    def backward_g(self, extra=None):
        self.optG.zero_grad()
        self.g_seeds.backward()

    def update_g(self, data, update=True):
        self.forward_g(data)
        self.compute_g_loss()
        if update:
            self.backward_g()
            self.optG.step()

    # ---- discriminator step (models/srgan.py:34-45, 63-67, 93-100)
    def forward_d(self, data):
        self.high_res, self.super_res = data["high_res"], data["super_res"]
        self.real_logits = self.D(self.high_res)
        self.fake_logits = self.D(self.super_res.detach())

    def compute_d_loss(self):
        self.LossR, d_real = self._loss("D/Real", "bce_logits", self.real_logits, 1.0)
        self.LossF, d_fake = self._loss("D/Fake", "bce_logits", self.fake_logits, 0.0)
        self.LossD = self.LossR + self.LossF
        self._d_seeds = Seeds([(self.real_logits, d_real), (self.fake_logits, d_fake)])

    def update_d(self, data):
        self.forward_d(data)
        self.compute_d_loss()
        self.optD.zero_grad()
        self._d_seeds.backward()
        self.optD.step()

    def get_metrics(self):
        names = ["G/MSE"] if self.pretrain else ["D/Real", "D/Fake", "G/Adv", "G/Con"]
        vals = torch.stack([self._scalars[n][0] for n in names]).tolist()          # one device-to-host copy
        m = {n: v * self._scalars[n][1] for n, v in zip(names, vals)}
        if self.pretrain:
            return {"G/MSE": m["G/MSE"], "G/Sum": m["G/MSE"]}
        return {"D/Sum": m["D/Real"] + m["D/Fake"], "D/Real": m["D/Real"], "D/Fake": m["D/Fake"],
                "G/Sum": m["G/Con"] + ADV_WEIGHT * m["G/Adv"], "G/Adv": m["G/Adv"], "G/Con": m["G/Con"]}
