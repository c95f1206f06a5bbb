import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ipr_gan_b200; ipr_gan_b200.enable_dropin()
import torch, networks
from oracle import ipr_oracle as orc
def fro(a, b):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    return float((a - b).norm() / (b.norm() + 1e-12)), float((a - b).abs().max() / (b.abs().max() + 1e-12))
torch.manual_seed(0)
G, D = networks.ConvGenerator32(), networks.SNDiscriminator32()
Go, Do = orc.make_generator(), orc.make_discriminator()
Go.load_state_dict(G.state_dict()); Do.load_state_dict(D.state_dict())
G, D = G.cuda(), D.cuda()
for B in (8, 64):
    G.zero_grad(); Go.zero_grad()
    z = torch.randn(B, 128)
    out = G(z.cuda()); ref = Go(z)
    print('G out B=%d' % B, fro(out, ref))
    g = torch.randn_like(ref)
    out.backward(g.cuda()); ref.backward(g)
    for (n, p), (_, q) in zip(G.named_parameters(), Go.named_parameters()):
        print('  dG', n, fro(p.grad, q.grad))
    D.zero_grad(); Do.zero_grad()
    x = torch.randn(B, 3, 32, 32).clamp(-1, 1)
    xg = x.clone().cuda().requires_grad_(True); xo = x.clone().requires_grad_(True)
    lo, lr = D(xg), Do(xo)
    print('D out', fro(lo, lr))
    torch.relu(1 - lo).mean().backward(); torch.relu(1 - lr).mean().backward()
    print('  dx', fro(xg.grad, xo.grad))
    for (n, p), (_, q) in zip(D.named_parameters(), Do.named_parameters()):
        print('  dD', n, fro(p.grad, q.grad))
    for (n, b), (_, c) in zip(D.named_buffers(), Do.named_buffers()):
        print('  bufD', n, fro(b, c))
