import sys, time, torch
sys.path.insert(0, '/root/repo')
from sarssl_b200 import ops
from sarssl_b200.learner import STFTLearner
from sarssl_b200.model import SARSSL
from sarssl_b200.optim import FusedAdam
dev = torch.device("cuda", 0)
torch.manual_seed(1)
model = SARSSL(sig_shape=(256, 256, 2, 2), device=dev); model.to(dev); model.set_compute_dtype(torch.bfloat16); model.set_dropout(0.1)
model.rng_state = ops.mt_seed(400000001); model.train()
learner = STFTLearner(model, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M"); learner.device = dev
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
sig = 0.1 * torch.randn(B, 65792, 2, device=dev)
opt = FusedAdam(model, lr=1e-3)
def step(tt):
    t0 = time.perf_counter(); x, = learner.data_preprocess(sig)
    t1 = time.perf_counter(); loss, diff, _ = model(x)
    t2 = time.perf_counter(); loss.backward()
    t3 = time.perf_counter(); opt.step(1e-3, grad_scale=1.0, zero_grad=True)
    t4 = time.perf_counter()
    tt.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
for _ in range(3): step([])
torch.cuda.synchronize()
tt = []
t0 = time.perf_counter()
for _ in range(5): step(tt)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue per step %.2f ms; total incl. drain %.2f ms/step" % ((t1 - t0) / 5 * 1e3, (t2 - t0) / 5 * 1e3))
for r in tt: print("  preprocess %.2f  forward %.2f  backward %.2f  adam %.2f ms" % tuple(1e3 * v for v in r))
