"""Flat Adam launch alone: time per launch and effective bandwidth (32 bytes per parameter: read p, g, m, v; write p, g=0, m, v)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ipr_gan_b200._lib import lib, check
L = lib()
dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for n in (2935873, 3811904, 16 * 1024 * 1024):
    p, g, m, v = (torch.randn(n + 3, device=dev)[:n] for _ in range(4))
    p, g, m, v = (torch.randn(n, device=dev) for _ in range(4))
    v.abs_()
    step = torch.zeros((), device=dev); ticket = torch.zeros(1, device=dev, dtype=torch.int32)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    def run():
        check(L.ipr_adam_flat_f32(ctypes.c_void_p(p.data_ptr()), ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(m.data_ptr()),
                                  ctypes.c_void_p(v.data_ptr()), n, 2e-4, 0.5, 0.999, 1e-8, 0.0, 1.0, 1,
                                  ctypes.c_void_p(step.data_ptr()), ctypes.c_void_p(ticket.data_ptr()), st), "adam")
    for _ in range(3): run()
    ts = []
    for cold in (True, False):
        ts = []
        for _ in range(10):
            if cold: flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        ts.sort(); t = ts[len(ts) // 2]
        print("n=%9d %s: %.1f us, %.2f TB/s" % (n, "cold (L2 flushed)" if cold else "warm", t * 1e3, 32.0 * n / (t * 1e-3) / 1e12))
