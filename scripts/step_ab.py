"""ms per graph-replayed pre-training step (B = 256, nt = 256, bf16) - for same-box A/B runs of environment switches:
    for v in 0 1 0 1; do SARSSL_GEMM_WIDE=$v python scripts/step_ab.py; done"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sarssl_b200 import ops
from sarssl_b200.learner import STFTLearner
from sarssl_b200.model import SARSSL
from sarssl_b200.optim import FusedAdam
dev = torch.device("cuda", 0)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = SARSSL(sig_shape=(256, 256, 2, 2), device=dev)
model.set_compute_dtype(torch.bfloat16)
model.set_dropout(0.1)
model.rng_state = ops.mt_seed(400000001)
model.train()
learner = STFTLearner(model, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
learner.device = dev
sig = 0.1 * torch.randn(nb, 65792, 2, device=dev)
opt = FusedAdam(model, lr=1e-3)
x, = learner.data_preprocess(sig)
loss, _, _ = model(x)
loss.backward()
opt.step(1e-3)
learner._eager_pretrain_steps = 1
g = learner.graphed_pretrain_step(sig, opt)
for _ in range(3):
    g.run(sig, 1e-3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 12
e0.record()
for _ in range(n):
    loss, _, _ = g.run(sig, 1e-3)
e1.record(); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / n:.3f} ms/step  ({nb * n / e0.elapsed_time(e1) * 1e3:.0f} clips/s), loss {float(loss):.4f}")
