"""Tiny fwd+bwd of the fused model on one GPU, checked against the CPU oracle (called from __graft_entry__.smoke)."""
import random

import torch


def run(dev):
    from oracle import sarssl_oracle as O
    from sarssl_b200.learner import STFTLearner
    from sarssl_b200.model import SARSSL
    nb, nt = 2, 8
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=2)
    m = SARSSL(sig_shape=(256, nt, 2, 2), device=dev)
    m.load_state_dict(O.synthetic_state_dict(7))
    m.to(dev)
    m.set_dropout(0.0)
    m.train()
    L = STFTLearner(m, 512, 0.5, 512, 1, 16000)
    L.device = dev
    x, = L.data_preprocess(sig.to(dev))
    random.seed(21)
    loss, diff, vis = m(x)
    loss.backward()
    sd = O.synthetic_state_dict(7)
    w = sd["decoder.proj.2.weight"].requires_grad_(True)
    random.seed(21)
    pidx, cidx = O.draw_masks(nb, nt, nt // 2, 2)
    rl, rd, _ = O.pretrain_forward(O.preprocess(sig), sd, pidx, cidx, training=True)
    rl.backward()
    err = abs(float(loss) - float(rl)) / float(rl)
    gerr = float((m.store.p("decoder.proj.2.weight").grad.cpu() - w.grad).norm() / w.grad.norm())
    assert err < 1e-4 and gerr < 1e-3, (err, gerr)
    print("smoke model ok: loss %.6f (oracle %.6f), decoder grad rel err %.2e" % (float(loss), float(rl), gerr))
