"""CycleGAN step with the reference's interface (models/cyclegan.py:10-164): two generators, two PatchGAN
discriminators, LSGAN + cycle + identity losses, image-history pools, linearly decaying learning rate;
``_modules`` = GA, GB, DA, DB, optG, optD, schG, schD, poolA, poolB."""
from itertools import chain

import torch
import torch.nn as nn
from torch import optim

import networks
from models.core import Model
from models.srgan import _make_opt
from models.util import Replica


class ImagePool(nn.Module):
    """History buffer of generated images kept as registered buffers so that it is checkpointed (models/util.py:5-35):
    until full every image is stored and returned; afterwards each image is swapped with a random stored one with
    probability 1/2."""

    def __init__(self, pool_size):
        super().__init__()
        self.pool_size = pool_size
        if pool_size > 0:
            self.register_buffer("images", torch.tensor([]))
            self.register_buffer("counts", torch.zeros([]))

    def load_state_dict(self, *args, **kwargs):
        self.images = torch.empty_like(args[0]["images"])
        super().load_state_dict(*args, **kwargs)

    def __call__(self, images):
        if self.pool_size <= 0:
            return images.detach()
        if self.counts < self.pool_size:
            self.images = torch.cat([self.images.to(images.device), images.detach()], dim=0)[:self.pool_size]
            self.counts += images.size(0)
            return images.detach()
        images = images.detach()
        swap = torch.rand(images.size(0)) > 0.5
        slot = torch.randperm(self.pool_size)[:images.size(0)]
        old = self.images[slot[swap]].clone()
        self.images[slot[swap]] = images[swap].detach()
        images[swap] = old
        return images.detach()


class CycleGAN(Model):
    def __init__(self, config, device=[torch.device("cpu"), ]):
        super().__init__()
        self.device = device
        ids = [d.index for d in device]

        def net(name):
            return Replica(getattr(networks, name)().to(device[0]), device_ids=ids)

        self.GA, self.GB = net(config.G), net(config.G)
        self.DA, self.DB = net(config.D), net(config.D)
        self.poolA, self.poolB = ImagePool(config.pool_size), ImagePool(config.pool_size)
        for m in (self.GA, self.GB, self.DA, self.DB):
            m.train()
        self.lambda_A, self.lambda_B, self.lambda_idt = config.lambda_A, config.lambda_B, config.lambda_idt
        self.optG = _make_opt(config, list(chain(self.GA.parameters(), self.GB.parameters())), device[0])
        self.optD = _make_opt(config, list(chain(self.DA.parameters(), self.DB.parameters())), device[0])
        half = config.epoch // 2
        decay = lambda e: 1.0 - max(0, e - half) / half                                   # noqa: E731
        self.schedulerG = optim.lr_scheduler.LambdaLR(self.optG, lr_lambda=decay)
        self.schedulerD = optim.lr_scheduler.LambdaLR(self.optD, lr_lambda=decay)
        self.MSE, self.L1 = nn.MSELoss(), nn.L1Loss()
        self._modules.update(GA=self.GA, GB=self.GB, DA=self.DA, DB=self.DB, optG=self.optG, optD=self.optD,
                             schG=self.schedulerG, schD=self.schedulerD, poolA=self.poolA, poolB=self.poolB)

    def get_metrics(self):
        names = ("G/A", "G/B", "G/CycA", "G/CycB", "G/IdtA", "G/IdtB", "G/Sum", "D/RealA", "D/FakeA", "D/SumA",
                 "D/RealB", "D/FakeB", "D/SumB")
        ts = (self.LossGA, self.LossGB, self.LossCycA, self.LossCycB, self.LossIdtA, self.LossIdtB, self.LossG,
              self.LossDRA, self.LossDFA, self.LossDA, self.LossDRB, self.LossDFB, self.LossDB)
        dev = self.LossG.device
        out = dict(zip(names, torch.stack([t.detach().to(dev).float() for t in ts]).tolist()))
        out["LR"] = self.optG.param_groups[0]["lr"]
        return out

    def forward_g(self, data):
        self.real_A, self.real_B = data["real_A"], data["real_B"]
        self.fake_B, self.fake_A = self.GA(self.real_A), self.GB(self.real_B)
        self.rec_A, self.rec_B = self.GB(self.fake_B), self.GA(self.fake_A)
        self.idt_A, self.idt_B = self.GA(self.real_B), self.GB(self.real_A)
        self.GA_logits, self.GB_logits = self.DA(self.fake_B), self.DB(self.fake_A)

    def forward_d(self, data):
        self.real_A, self.real_B = data["real_A"], data["real_B"]
        self.fake_A, self.fake_B = self.poolA(data["fake_A"]), self.poolB(data["fake_B"])
        self.RA_logits, self.FA_logits = self.DB(self.real_A), self.DB(self.fake_A.detach())
        self.RB_logits, self.FB_logits = self.DA(self.real_B), self.DA(self.fake_B.detach())

    def compute_g_loss(self):
        self.real_A = self.real_A.to(self.rec_A.device)
        self.real_B = self.real_B.to(self.rec_B.device)
        self.LossGA = self.MSE(self.GA_logits, torch.ones_like(self.GA_logits))
        self.LossGB = self.MSE(self.GB_logits, torch.ones_like(self.GB_logits))
        self.LossCycA = self.L1(self.rec_A, self.real_A) * self.lambda_A
        self.LossCycB = self.L1(self.rec_B, self.real_B) * self.lambda_B
        self.LossG = self.LossGA + self.LossGB + self.LossCycA + self.LossCycB
        if self.lambda_idt > 0:
            self.LossIdtA = self.L1(self.idt_A, self.real_B) * self.lambda_B
            self.LossIdtB = self.L1(self.idt_B, self.real_A) * self.lambda_A
            self.LossG = self.LossG + self.lambda_idt * (self.LossIdtA + self.LossIdtB)
        else:
            self.LossIdtA = self.LossIdtB = torch.zeros([])

    def compute_d_loss(self):
        self.LossDRA = self.MSE(self.RB_logits, torch.ones_like(self.RB_logits))
        self.LossDFA = self.MSE(self.FB_logits, torch.zeros_like(self.FB_logits))
        self.LossDA = (self.LossDRA + self.LossDFA) * 0.5
        self.LossDRB = self.MSE(self.RA_logits, torch.ones_like(self.RA_logits))
        self.LossDFB = self.MSE(self.FA_logits, torch.zeros_like(self.FA_logits))
        self.LossDB = (self.LossDRB + self.LossDFB) * 0.5

    def update_lr(self):
        self.schedulerG.step()
        self.schedulerD.step()

    def update_g(self, data, update=True):
        self.forward_g(data)
        self.compute_g_loss()
        if update:
            self.optG.zero_grad()
            self.LossG.backward()
            self.optG.step()

    def update_d(self, data):
        self.forward_d(data)
        self.compute_d_loss()
        self.optD.zero_grad()
        self.LossDA.backward()
        self.LossDB.backward()
        self.optD.step()
