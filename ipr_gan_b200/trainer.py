"""Protected IPR-DCGAN training step as a user of the drop-in API would run it
(``models.DCGAN`` -> ``BlackBoxWrapper`` -> ``WhiteBoxWrapper``, the stack built by
experiments/image_generation.py:38-84), plus the two B200-specific pieces around it:

* the whole step (update_d + update_g: 2 generator forwards, 3 discriminator forwards, all backwards,
  SSIM + sign loss, both Adam updates) is captured ONCE into a CUDA graph and replayed -- at 64-512
  samples per GPU the step is a few hundred short kernels, so launch latency would otherwise dominate;
* data parallelism is one process per GPU: gradients are all-reduced over NCCL inside
  ``optimizer.step()`` (``dist.AllReduceOptimizer``), BatchNorm statistics stay per rank exactly like the
  reference's ``nn.DataParallel`` replicas (experiments/base.py:36-39).
"""
import torch

import ipr_gan_b200

ipr_gan_b200.enable_dropin()

import models  # noqa: E402
from configs import presets  # noqa: E402
from ipr_gan_b200 import _lib, engine  # noqa: E402


class ProtectedDCGANTrainer(object):
    def __init__(self, batch, device, seed=1234, fn_inp="TransformDist", use_graph=True, size=32,
                 device_latents=False):
        self.batch, self.device, self.use_graph = batch, device, use_graph
        self.device_latents = device_latents
        torch.manual_seed(seed)                      # identical initial replicas on every rank
        mcfg = presets.dcgan_model(size)
        if use_graph:
            mcfg.opt_param["capturable"] = True
        model = models.DCGAN(mcfg, device=[device])
        model = models.BlackBoxWrapper(model, presets.dcgan_blackbox(fn_inp=fn_inp))
        model = models.WhiteBoxWrapper(model, presets.dcgan_whitebox())
        self.model = model
        self.real = torch.zeros(batch, 3, size, size, device=device)
        self.latent = torch.zeros(batch, 128, device=device)
        self.real_host = torch.zeros(batch, 3, size, size).pin_memory()
        self.latent_host = torch.zeros(batch, 128).pin_memory()
        self.graph = None
        self.launches_per_step = None
        # latents drawn on the device (Philox, position kept in device memory): no host randn + H2D copy per step
        # (experiments/image_generation.py:93-96); one stream per rank
        from ipr_gan_b200 import dist, ops
        self.normal = ops.DeviceNormal(device, seed * 7919 + dist.rank()) if device_latents else None

    # one reference-API step on whatever is in the static device buffers
    def _step(self):
        if self.normal is not None:
            self.normal.fill_(self.latent)
        self.model.update_d({"real_sample": self.real, "latent": self.latent})
        self.model.update_g({"fake_sample": self.model.fake_sample})

    def set_inputs(self, real, latent):
        """Stage a host (or device) batch into the static device buffers (async when the source is pinned)."""
        self.real.copy_(real, non_blocking=True)
        if latent is not None:
            self.latent.copy_(latent, non_blocking=True)

    def capture(self, warmup=3):
        """Warm up eagerly on a side stream, then capture the step into a CUDA graph."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                before = _lib.launch_count()
                self._step()
                self.launches_per_step = _lib.launch_count() - before
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        if not self.use_graph:
            return
        engine.reset_caches()                       # weight packing must be part of the captured step
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: NCCL's watchdog thread may query events while this thread captures (multi-GPU runs)
        with torch.cuda.graph(self.graph, stream=s, capture_error_mode="thread_local"):
            before = _lib.launch_count()
            self._step()
            self.launches_per_step = _lib.launch_count() - before
        torch.cuda.synchronize(self.device)

    def step(self):
        """One training step on the current contents of the device buffers."""
        if self.graph is not None:
            self.graph.replay()
            board = self.model.board
            if board is not None:
                board.touch()                       # the replay rewrote the loss slots behind the host's back
        else:
            self._step()

    def step_from_host(self, real_cpu, latent_cpu=None):
        """The call a user of the reference makes (experiments/image_generation.py:92-101): host tensors in,
        metrics dict (Python floats) out.  With ``device_latents`` the latent argument is not needed."""
        self.set_inputs(real_cpu, latent_cpu)
        self.step()
        return self.model.get_metrics()
