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

    def _fused(self):
        """True when the wrapped model back-propagates from seed gradients (``seeded``: the native DCGAN / SRGAN /
        CycleGAN steps) and the watermark loss is SSIM: forward + backward of the loss are then ONE launch."""
        return bool(self.seeded) and self.config.loss_fn == "ssim"

    def _wm_slot(self):
        """(1-element view the loss value is written to, scale): the inner model's metrics board when it has one."""
        board = self.board
        if board is not None:
            return board.slot(board.P_WM), board.loss_scale
        if getattr(self, "_w_slot", None) is None:
            self._w_slot = torch.zeros(1, device=self.device[0])
        return self._w_slot, 1.0

    def compute_g_loss(self):
        self.LossG = self.model.LossG
        if self._fused():
            slot, scale = self._wm_slot()
            if self.inhibit:
                slot.zero_()
            else:
                # forward + backward of the SSIM loss in one launch: the scalar goes to its slot, lambda * dLoss/dx
                # becomes the seed gradient of the trigger pass (no autograd node, no multiply by grad_output)
                from ipr_gan_b200 import ops
                _, dx = ops.ssim_loss_fwd_bwd(self.Gxwm.detach(), self.ywm, self.config.normalized, grad_scale=self.Lambda,
                                              loss_out=slot, loss_scale=scale)
                self.g_seeds.append((self.Gxwm, dx))
            self.LossW = slot[0]
            return
        if self.inhibit:
            self.LossW = torch.zeros_like(self.LossG)
        else:
            self.LossW = self.loss_fn(self.Gxwm, self.ywm)
            if self.seeded:                  # seeded inner model, other loss (l1 / mse): lambda * dLossW seeds the backward
                self.g_seeds.append((self.LossW, torch.full_like(self.LossW, float(self.Lambda))))

    def _triggers(self, source, produced):
        """xwm = fn_inp(source), ywm = fn_out(produced) (models/wrappers.py:48-51).  The protected-DCGAN pairing --
        TransformDist on the latents, a pasted patch on the images -- is ONE launch that reads each tensor once."""
        import tools
        from tools.triggers import _PatchTrigger
        f_in, f_out = self.fn_inp.module, self.fn_out.module
        if isinstance(f_in, tools.TransformDist) and isinstance(f_out, _PatchTrigger) and produced.is_cuda:
            from ipr_gan_b200 import ops
            z = source.detach().to(produced.device, torch.float32, non_blocking=True)
            return ops.trigger_pair(produced.detach(), f_out.fg, f_out.bg, f_out.position, f_out.config.size, z)
        return self.fn_inp(source.detach()), self.fn_out(produced.detach())

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
                    self.xwm, self.ywm = self._triggers(source, produced)
                with DisableBatchNormStats(net):
                    self.Gxwm = net(self.xwm)
            main.wait_stream(aux)
            for t in (self.xwm, self.ywm, self.Gxwm):
                t.record_stream(main)
            return
        pass_stream = getattr(net, "_ipr_pass_stream", None)
        if pass_stream is not None and produced.is_cuda:
            # the wrapped model keeps every pass of the target network on one stream (CycleGAN: a stream per
            # direction, models/cyclegan.py): the trigger pass joins them there, so that the target's gradient
            # accumulations stay ordered
            main = torch.cuda.current_stream(produced.device)
            pass_stream.wait_stream(main)
            with torch.cuda.stream(pass_stream):
                with torch.no_grad():
                    self.xwm, self.ywm = self._triggers(source, produced)
                with DisableBatchNormStats(net):
                    self.Gxwm = net(self.xwm)
            main.wait_stream(pass_stream)
            for t in (self.xwm, self.ywm, self.Gxwm):
                t.record_stream(main)
            return
        with torch.no_grad():
            self.xwm, self.ywm = self._triggers(source, produced)
        with DisableBatchNormStats(net):
            self.Gxwm = net(self.xwm)

    def get_metrics(self):
        metrics = self.model.get_metrics()
        if not self.inhibit:
            board = self.board if self._fused() else None
            w = board.fetch()[board.P_WM] if board is not None else self.LossW.item()
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
            root = self
            while isinstance(root, Wrapper):
                root = root.model
            # recorded by the inner model just before its optD.step(); valid for ONE generator step (a second
            # update_g without a new update_d must see the generator weights the first one wrote)
            ev = getattr(root, "pre_step_event", None)
            if ev is not None:
                root.pre_step_event = None
            else:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
            self._fork_event = ev

    def update_g(self, data, update=True):
        self._mark_fork()
        self.model.update_g(data, update=False)
        self.forward_g(data)
        self.compute_g_loss()
        if update:
            _backward_and_step(self, lambda: self.LossG + self.Lambda * self.LossW)


def _backward_and_step(wrapper, total):
    """optG.zero_grad(); total.backward(); optG.step() (models/wrappers.py:70-74, 121-125).  On the fused path the
    backward starts from the seed gradients the loss launches produced (`backward_g` of the innermost model)."""
    backward_g = wrapper.backward_g               # falls through to the wrapped model; None for other model types
    if backward_g is not None:
        backward_g(None if wrapper.seeded else total())
    else:
        wrapper.model.optG.zero_grad()
        total().backward()
    wrapper.model.optG.step()


class _SignHook(object):
    """Hands the sign vector of normalisation layer i to that layer's backward kernel exactly once per armed step
    (engine._GeneratorFn.backward): d(sign loss)/d(gamma) is added inside the BatchNorm backward launch."""

    def __init__(self, loss_model, signs):
        self.gamma_0, self.signs = loss_model.gamma_0, signs
        self.armed = [False] * len(signs)

    def arm(self):
        self.armed = [True] * len(self.signs)

    def __call__(self, i):
        if not self.armed[i]:
            return None, 0.0, 0.0
        self.armed[i] = False
        return self.signs[i], self.gamma_0, 1.0


class WhiteBoxWrapper(Wrapper):
    def __init__(self, model, config):
        super().__init__(model, config)
        self.configure()

    def configure(self):
        target = getattr(self.model, self.config.target)
        self.loss_model = tools.SignLossModel(target, self.config).to(self.device[0])
        self._modules["sign"] = self.loss_model
        self._sign_hook = None
        self._sign_slot = None
        import networks
        fused_dcgan = self.board is not None and isinstance(target.module, networks.ConvGenerator)
        if fused_dcgan or getattr(target.module, "_ipr_native_norms", False):
            # the target's normalisation layers run on this library: d(sign loss)/d(gamma) is added inside their
            # backward launches (BatchNorm and InstanceNorm alike), the value is one launch over all gammas
            _, signs = self.loss_model._collect(target)
            self._sign_hook = _SignHook(self.loss_model, signs)
            object.__setattr__(target.module, "_ipr_sign_hook", self._sign_hook)
            if not fused_dcgan:
                self._sign_slot = torch.zeros(1, device=self.device[0])

    def compute_g_loss(self):
        target = getattr(self.model, self.config.target)
        self.LossG = self.model.LossG
        board = self.board if (self._sign_hook is not None and self._sign_slot is None) else None
        if board is not None:
            if self.inhibit:
                board.slot(board.P_SIGN).zero_()
            else:
                # value: one launch over all gammas into its board slot; gradient: inside the BatchNorm backward
                self.loss_model.value_into(target, board.slot(board.P_SIGN), board.loss_scale)
                self._sign_hook.arm()
            self.LossS = board.scalar(board.P_SIGN)
        elif self._sign_hook is not None:
            if self.inhibit:
                self.LossS = torch.zeros_like(self.LossG)
            else:
                self.loss_model.value_into(target, self._sign_slot, 1.0)
                self._sign_hook.arm()
                self.LossS = self._sign_slot[0]      # a constant in the autograd graph: its gradient rides in the hook
        else:
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
            board = self.board if (self._sign_hook is not None and self._sign_slot is None) else None
            s = board.fetch()[board.P_SIGN] if board is not None else self.LossS.item()
            metrics["P/SignLoss"] = s
            metrics["G/Sum"] += s
        return metrics

    def update_g(self, data, update=True):
        self.model.update_g(data, update=False)
        self.forward_g(data)
        self.compute_g_loss()
        if update:
            _backward_and_step(self, lambda: self.LossG + self.Lambda * self.LossW + self.LossS)
