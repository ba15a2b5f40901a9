"""2-rank check of the graph-replayed data-parallel step (run under torchrun): replicas identical, weights equal to the eager epoch's."""
import faulthandler, os, sys
faulthandler.dump_traceback_later(50, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from oracle import sarssl_oracle as O
from sarssl_b200 import ops
from sarssl_b200.learner import STFTLearner
from sarssl_b200.model import SARSSL
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
nb, nt = 2, 16
finals = {}
for mode in (False, None):
    m3 = SARSSL(sig_shape=(256, nt, 2, 2), device=dev); m3.load_state_dict(O.synthetic_state_dict(7)); m3.to(dev)
    m3.set_dropout(0.1); m3.set_compute_dtype(torch.bfloat16); m3.rng_state = ops.mt_seed(99); m3.train()
    L3 = STFTLearner(m3, 512, 0.5, 512, 1, 16000); L3.device = dev; L3.mul_gpu()
    data = [[O.synthetic_waveforms(nb * world, (nt + 1) * 256, 2, seed=70 + i)[rank * nb:(rank + 1) * nb]] for i in range(4)]
    print(rank, "epoch start", mode, flush=True)
    L3.pretrain_epoch(data, lr=1e-3, epoch=1, use_graph=mode)
    torch.cuda.synchronize()
    print(rank, "epoch done", mode, flush=True)
    flat = [torch.empty_like(m3.store.flat) for _ in range(world)]
    dist.all_gather(flat, m3.store.flat)
    assert all(torch.equal(flat[0], f) for f in flat), mode
    finals[mode] = m3.store.flat.clone()
gerr = float((finals[None] - finals[False]).norm() / finals[False].norm())
print(rank, "graph vs eager DP weights rel diff", gerr, flush=True)
assert gerr < 2e-3
dist.destroy_process_group()
