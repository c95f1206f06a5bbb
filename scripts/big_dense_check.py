"""Dense-layer parity at batch-512 tile counts (many tiles per persistent CTA; the pytest cases use small batches)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_dense as T
for name, args in [("test_conv3_forward_bias_lrelu", (512, 64, 128, 16)), ("test_conv3_forward_bias_lrelu", (512, 128, 256, 8)),
                   ("test_conv4s2_forward", (512, 64, 64, 32)), ("test_convT4s2_forward_with_stats", (512, 128, 64, 16)),
                   ("test_convT4s2_forward_with_stats", (512, 512, 256, 4)), ("test_tap_expanded_three_channel_layers", (512, 32))]:
    for rep in range(2):
        try:
            getattr(T, name)(*args)
            print("ok  ", name, args)
        except AssertionError as e:
            print("FAIL", name, args, str(e)[:100])
for kind, B, C, O, H in [("conv3_dgrad", 512, 128, 64, 16), ("conv4s2_dgrad", 512, 64, 64, 16), ("conv4s2_dgrad", 512, 128, 128, 8)]:
    for fn in [n for n in dir(T) if n.startswith("test_") and "dgrad" in n]:
        try:
            getattr(T, fn)(kind, B, C, O, H)
            print("ok  ", fn, kind, B, C, O, H)
        except TypeError:
            pass
        except AssertionError as e:
            print("FAIL", fn, kind, B, C, O, H, str(e)[:100])
