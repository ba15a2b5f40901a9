"""Tiny fwd+bwd of the fused model on one GPU, checked against the CPU oracle (called from __graft_entry__.smoke).
Runs both precisions: fp32 mode (CUDA-core GEMM / conv, the 1e-4 forward gate) and bf16 mode (tcgen05 GEMM + tcgen05
implicit-GEMM conv, the benchmarked path), so the driver's launch record of smoke() lists the tensor-core kernels."""
import random

import torch


def _step(dev, sig, nt, dtype):
    from oracle import sarssl_oracle as O
    from sarssl_b200.learner import STFTLearner
    from sarssl_b200.model import SARSSL
    m = SARSSL(sig_shape=(256, nt, 2, 2), device=dev)
    m.load_state_dict(O.synthetic_state_dict(7))
    m.to(dev)
    m.set_dropout(0.0)
    m.set_compute_dtype(dtype)
    m.train()
    L = STFTLearner(m, 512, 0.5, 512, 1, 16000)
    L.device = dev
    x, = L.data_preprocess(sig.to(dev))
    random.seed(21)
    loss, diff, vis = m(x)
    loss.backward()
    return m, float(loss)


def run(dev):
    from oracle import sarssl_oracle as O
    nb, nt = 2, 8
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=2)
    sd = O.synthetic_state_dict(7)
    w = sd["decoder.proj.2.weight"].requires_grad_(True)
    random.seed(21)
    pidx, cidx = O.draw_masks(nb, nt, nt // 2, 2)
    rl, rd, _ = O.pretrain_forward(O.preprocess(sig), sd, pidx, cidx, training=True)
    rl.backward()
    # fp32 mode
    m, loss = _step(dev, sig, nt, torch.float32)
    err = abs(loss - float(rl)) / float(rl)
    gerr = float((m.store.p("decoder.proj.2.weight").grad.cpu() - w.grad).norm() / w.grad.norm())
    assert err < 1e-4 and gerr < 1e-3, (err, gerr)
    # bf16 mode: the tcgen05 kernels
    mb, lossb = _step(dev, sig, nt, torch.bfloat16)
    errb = abs(lossb - float(rl)) / float(rl)
    gerrb = float((mb.store.p("decoder.proj.2.weight").grad.cpu() - w.grad).norm() / w.grad.norm())
    assert mb.engine.k.tc_launches > 50, "bf16 mode did not run the tcgen05 kernels"
    assert errb < 2e-2 and gerrb < 2e-2, (errb, gerrb)
    print("smoke model ok: fp32 loss %.6f (oracle %.6f, decoder grad rel err %.2e); bf16/tcgen05 loss %.6f (rel err %.2e, decoder grad rel err %.2e, "
          "%d tensor-core launches)" % (loss, float(rl), gerr, lossb, errb, gerrb, mb.engine.k.tc_launches))
