import sys, random, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_model_gpu import *
nb, nt = 3, 16
sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=5)
rl, rd, rvis, sd, taps = run_oracle(sig, nt, 7, 11)
gmax = max(float(sd[k].grad.norm()) for k in sd if sd[k].grad is not None)
res = {}
for use_tc in (False, True):
    m = build(nt, dtype=torch.bfloat16)
    m._engine().k.use_tc = use_tc
    loss, diff, vis = run_ours(m, sig, 11)
    print('use_tc', use_tc, 'loss', float(loss), float(rl), 'pred rel', rel(vis['pred'].float().cpu(), rvis['pred']), 'tc launches', m.engine.k.tc_launches)
    res[use_tc] = {k: p.grad.detach().cpu().double().clone() for k, p in m.named_parameters()}
rows = []
for k in res[True]:
    r = sd[k].grad.double(); den = float(r.norm()) + 1e-3 * gmax
    rows.append((float((res[True][k]-r).norm())/den, float((res[False][k]-r).norm())/den, float((res[True][k]-res[False][k]).norm())/den, k))
rows.sort(reverse=True)
for e in rows[:14]: print('tc %.4f simt %.4f tc-vs-simt %.4f %s' % e)
import statistics
print('median tc %.4f simt %.4f' % (statistics.median(r[0] for r in rows), statistics.median(r[1] for r in rows)))
