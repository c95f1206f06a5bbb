"""IPR protection wrappers with the reference's interface (models/wrappers.py:7-125).

BlackBoxWrapper: trigger input ``xwm = fn_inp(x)``, watermarked target ``ywm = fn_out(y)`` (no_grad, detached),
second generator forward on the trigger with BatchNorm running statistics frozen, reconstruction loss
``LossW = loss_fn(Gxwm, ywm)``; total ``LossG + lambda * LossW``.
WhiteBoxWrapper: adds the sign loss ``LossS`` over the target's normalisation gammas.
Both share ``_modules`` with the wrapped model so that ``fn_inp`` / ``fn_out`` / ``sign`` are checkpointed.
"""
import torch

import tools
from models.core import Wrapper
from models.util import DisableBatchNormStats, Replica


class BlackBoxWrapper(Wrapper):
    def __init__(self, model, config):
        super().__init__(model, config)
        self.configure()

    def configure(self):
        normalized = self.config.normalized
        ids = [d.index for d in self.device]
        dev0 = self.device[0]

        def build(spec):
            return Replica(getattr(tools, spec.type)(spec, normalized=normalized).to(dev0), device_ids=ids)

        self.fn_inp = build(self.config.fn_inp)
        self.fn_out = build(self.config.fn_out)
        self.Lambda = self.config["lambda"]
        self.loss_fn = getattr(tools, self.config.loss_fn)(normalized=normalized)

        self._modules = self.model._modules
        self._modules["fn_inp"] = self.fn_inp
        self._modules["fn_out"] = self.fn_out

    def compute_g_loss(self):
        self.LossG = self.model.LossG
        if self.inhibit:
            self.LossW = torch.zeros_like(self.LossG)
        else:
            self.LossW = self.loss_fn(self.Gxwm, self.ywm)

    def forward_g(self, data):
        self.inhibit = data.get("inhibit_bbox", False)
        if self.inhibit:
            return
        source = getattr(self.model, self.config.input_var)
        produced = getattr(self.model, self.config.output_var)
        net = getattr(self.model, self.config.target)
        fork = getattr(self, "_fork_event", None)
        self._fork_event = None
        if fork is not None:
            # the trigger pass is independent of the adversarial pass the inner model just enqueued: run it (forward
            # and, through autograd, backward) on the second stream, forked from where update_g started
            from ipr_gan_b200 import engine
            dev = self.device[0]
            main, aux = torch.cuda.current_stream(dev), engine.aux_stream(dev)
            aux.wait_event(fork)
            with torch.cuda.stream(aux):
                with torch.no_grad():
                    self.xwm = self.fn_inp(source.detach())
                    self.ywm = self.fn_out(produced.detach())
                with DisableBatchNormStats(net):
                    self.Gxwm = net(self.xwm)
            main.wait_stream(aux)
            for t in (self.xwm, self.ywm, self.Gxwm):
                t.record_stream(main)
            return
        with torch.no_grad():
            self.xwm = self.fn_inp(source.detach())
            self.ywm = self.fn_out(produced.detach())
        with DisableBatchNormStats(net):
            self.Gxwm = net(self.xwm)

    def get_metrics(self):
        metrics = self.model.get_metrics()
        if not self.inhibit:
            w = self.LossW.item()
            metrics[f"P/{self.config.loss_fn.upper()}"] = w
            metrics["G/Sum"] += self.Lambda * w
        return metrics

    def _mark_fork(self):
        """Record where update_g starts, so forward_g can fork the trigger pass from there (CUDA + native DCGAN nets;
        `_concurrent` falls through to the wrapped model and is None for models without concurrent passes)."""
        self._fork_event = None
        dev = self.device[0]
        concurrent = self._concurrent
        if dev.type == "cuda" and concurrent is not None and concurrent():
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            self._fork_event = ev

    def update_g(self, data, update=True):
        self._mark_fork()
        self.model.update_g(data, update=False)
        self.forward_g(data)
        self.compute_g_loss()
        if update:
            self.model.optG.zero_grad()
            (self.LossG + self.Lambda * self.LossW).backward()
            self.model.optG.step()


class WhiteBoxWrapper(Wrapper):
    def __init__(self, model, config):
        super().__init__(model, config)
        self.configure()

    def configure(self):
        target = getattr(self.model, self.config.target)
        self.loss_model = tools.SignLossModel(target, self.config).to(self.device[0])
        self._modules["sign"] = self.loss_model

    def compute_g_loss(self):
        target = getattr(self.model, self.config.target)
        self.LossG = self.model.LossG
        self.LossS = torch.zeros_like(self.LossG) if self.inhibit else self.loss_model(target)
        if hasattr(self.model, "LossW"):
            self.Lambda = self.model.Lambda
            self.LossW = self.model.LossW
        else:
            self.Lambda = 0
            self.LossW = torch.zeros_like(self.LossS)

    def forward_g(self, data):
        self.inhibit = data.get("inhibit_wbox", False)

    def get_metrics(self):
        metrics = self.model.get_metrics()
        if not self.inhibit:
            s = self.LossS.item()
            metrics["P/SignLoss"] = s
            metrics["G/Sum"] += s
        return metrics

    def update_g(self, data, update=True):
        self.model.update_g(data, update=False)
        self.forward_g(data)
        self.compute_g_loss()
        if update:
            self.model.optG.zero_grad()
            (self.LossG + self.Lambda * self.LossW + self.LossS).backward()
            self.model.optG.step()
