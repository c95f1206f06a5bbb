"""``Model`` / ``Wrapper`` bases with the reference's semantics (models/base.py:4-79):
* ``state_dict`` / ``load_state_dict`` walk ``_modules`` (networks, optimizers AND protection tools);
* ``load_state_dict(strict=False)`` silently skips missing entries, ``strict=True`` asserts;
* ``Wrapper.__getattr__`` falls through to the wrapped model and yields ``None`` for unknown names
  (so ``hasattr(wrapper, anything)`` is always true -- callers rely on it, models/wrappers.py:98)."""
from abc import ABC, abstractmethod
from collections import OrderedDict


class Seeds(list):
    """(tensor, gradient) pairs a fused step back-propagates from: every loss launch of this library returns the
    loss value AND d(loss)/d(its input), so the backward pass starts at the networks' outputs -- the loss arithmetic
    never becomes an autograd graph (``total.backward()`` of models/dcgan.py:76, models/wrappers.py:72,123)."""

    def add(self, tensor, grad):
        self.append((tensor, grad))

    def backward(self):
        import torch
        if self:
            torch.autograd.backward([t for t, _ in self], [g for _, g in self])
        del self[:]


class Model(ABC):
    def __init__(self):
        self._modules = OrderedDict()

    @abstractmethod
    def compute_g_loss(self): ...

    @abstractmethod
    def compute_d_loss(self): ...

    @abstractmethod
    def forward_d(self): ...

    @abstractmethod
    def forward_g(self): ...

    @abstractmethod
    def get_metrics(self): ...

    @abstractmethod
    def update_d(self): ...

    @abstractmethod
    def update_g(self): ...

    def state_dict(self):
        return OrderedDict((name, part.state_dict()) for name, part in self._modules.items())

    def load_state_dict(self, state_dict, strict=False):
        for name, part in self._modules.items():
            if strict:
                assert name in state_dict, f"Missing key: {name}"
            if name in state_dict:
                part.load_state_dict(state_dict[name])


class Wrapper(Model):
    def __init__(self, model, config):
        self.model = model
        self.config = config

    def __getattr__(self, key):
        own = self.__dict__
        if key in own:
            return own[key]
        inner = own.get("model")
        if inner is not None and hasattr(inner, key):
            return getattr(inner, key)
        return None

    @abstractmethod
    def configure(self): ...

    def compute_d_loss(self):
        self.model.compute_d_loss()

    def forward_d(self, data):
        self.model.forward_d(data)

    def update_d(self, data):
        self.model.update_d(data)
