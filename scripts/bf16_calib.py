"""Calibration (GPU box): how far does the *reference-style* torch autocast(bf16) run sit from fp32, per gradient tensor?
Uses the oracle's functional model on CUDA (library kernels) - informational only, not part of the product."""
import sys, random, torch
sys.path.insert(0, '/root/repo')
from oracle import sarssl_oracle as O
nb, nt = 3, 16
sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=5)
x = O.preprocess(sig)
def run(autocast):
    sd = {k: v.cuda() for k, v in O.synthetic_state_dict(7).items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k and not k.endswith(".pe"): v.requires_grad_(True)
    random.seed(11); pidx, cidx = O.draw_masks(nb, nt, nt // 2, 2)
    torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
        loss, diff, vis = O.pretrain_forward(x.cuda(), sd, pidx.cuda(), cidx.cuda(), training=True)
    loss.backward()
    return float(loss), {k: v.grad.float().cpu() for k, v in sd.items() if v.requires_grad}
l32, g32 = run(False)
l16, g16 = run(True)
gmax = max(float(g.norm()) for g in g32.values())
errs = sorted(((float((g16[k] - g32[k]).norm()) / (float(g32[k].norm()) + 1e-3 * gmax), k) for k in g32), reverse=True)
print('loss fp32 %.6f autocast-bf16 %.6f' % (l32, l16))
for e in errs[:8]: print('%.4f %s' % e)
print('median %.4f' % errs[len(errs)//2][0])
