"""cProfile of eager protected DCGAN steps (host side): where the ~37 us per launch go."""
import cProfile, os, pstats, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ipr_gan_b200.trainer import ProtectedDCGANTrainer
dev = torch.device("cuda", 0)
tr = ProtectedDCGANTrainer(64, dev, use_graph=False)
tr.capture(3)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    tr.step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
