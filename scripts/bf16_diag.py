import sys, random, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_model_gpu import *
nb, nt = 3, 16
sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=5)
m = build(nt, dtype=torch.bfloat16)
loss, diff, vis = run_ours(m, sig, 11)
rl, rd, rvis, sd, taps = run_oracle(sig, nt, 7, 11)
print('loss', float(loss), float(rl), 'pred rel', rel(vis['pred'].float().cpu(), rvis['pred']))
gmax = max(float(sd[k].grad.norm()) for k,_ in m.named_parameters())
errs = []
for k, p in m.named_parameters():
    g, r = p.grad.detach().cpu().double(), sd[k].grad.double()
    errs.append((float((g - r).norm()) / (float(r.norm()) + 1e-3 * gmax), k, float(r.norm())))
errs.sort(reverse=True)
for e in errs[:25]: print('%.4f %-90s %.3e' % e)
print('median', errs[len(errs)//2][0])
