"""CPU: checkpoint interop.  File names / dictionary layout of the reference's Learner, DataParallel
`module.` prefixes, partial loads with a key prefix, epoch averaging; and - when the real reference is present (build
container) - round trips between the reference's Learner/model and ours in both directions."""
import os

import pytest
import torch

from oracle import ref_shim
from oracle import sarssl_oracle as O
from sarssl_b200.learner import STFTLearner
from sarssl_b200.model import SARSSL


def make_learner(seed):
    m = SARSSL(sig_shape=(256, 16, 2, 2), device="cpu")
    m.device = torch.device("cpu")
    m.load_state_dict(O.synthetic_state_dict(seed))
    L = STFTLearner(m, 512, 0.5, 512, 1, 16000)
    L.device = "cpu"
    return L


def test_save_resume_best_epoch_ensemble(tmp_path):
    d = str(tmp_path)
    L = make_learner(1)
    assert L.early_stopping(-2.0, patience=2) == (False, True)
    L.save_checkpoint(3, d, is_best_epoch=True, save_extra_hist=True)
    assert sorted(os.listdir(d)) == ["best_model.tar", "latest_model.tar", "model3.tar"]
    ck = torch.load(os.path.join(d, "latest_model.tar"), weights_only=False)
    assert set(ck) == {"epoch", "max_score", "model"} and len(ck["model"]) == 214 and ck["max_score"] == -2.0
    L2 = make_learner(2)
    L2.resume_checkpoint(d)
    assert L2.start_epoch == 4 and L2.max_score == -2.0
    for k, v in L.model.state_dict().items():
        assert torch.equal(v, L2.model.state_dict()[k]), k
    assert L2.model.store.p("decoder.proj.0.weight").data_ptr() == L2.model.store.flat.data_ptr() + 4 * L2.model.store.offsets["decoder.proj.0.weight"][0]
    # epoch averaging
    L3 = make_learner(5)
    L3.save_checkpoint(4, d, save_extra_hist=True)
    L2.ensembling(d, [3, 4])
    k = "spat_encoder.embed.layers.1.sequential.3.module.sequential.1.linear.weight"
    want = 0.5 * L.model.state_dict()[k] + 0.5 * L3.model.state_dict()[k]
    assert torch.allclose(L2.model.state_dict()[k], want)
    assert L2.load_checkpoint_ensemble(d) == [3, 4]
    L2.load_checkpoint_epoch(d, 4)
    assert torch.equal(L2.model.state_dict()[k], L3.model.state_dict()[k])
    with pytest.raises(ValueError):
        torch.save({"epoch": 9, "max_score": 0, "model": L.model.state_dict()}, os.path.join(d, "model7.tar"))
        L2.load_checkpoint_epoch(d, 7)
    L2.remove_checkpoint_epochs(d, [3, 4])
    assert not os.path.exists(os.path.join(d, "model3.tar"))
    assert L2.early_stopping(-3.0, patience=1) == (True, False)


def test_dataparallel_prefix_and_partial_load(tmp_path):
    d = str(tmp_path)
    L = make_learner(1)
    sd = {"module." + k: v for k, v in L.model.state_dict().items()}          # what nn.DataParallel would have written
    torch.save({"epoch": 1, "max_score": 0.5, "model": sd}, os.path.join(d, "best_model.tar"))
    L2 = make_learner(2)
    assert L2.load_checkpoint_best(d) == 1
    assert torch.equal(L2.model.state_dict()["decoder.proj.2.bias"], L.model.state_dict()["decoder.proj.2.bias"])
    # partial load: a checkpoint holding only the spatial encoder, keys relative to it
    part = {k[len("spat_encoder."):]: v for k, v in L.model.state_dict().items() if k.startswith("spat_encoder.")}
    torch.save({"epoch": 2, "max_score": 0.1, "model": part}, os.path.join(d, "best_model.tar"))
    L3 = make_learner(3)
    before = L3.model.state_dict()["spec_encoder.patch_embed.0.weight"].clone()
    L3.load_checkpoint_best(d, as_all_state=False, param_frozen=True, ex_key="spat_encoder.")
    assert torch.equal(L3.model.state_dict()["spat_encoder.patch_embed.3.weight"], L.model.state_dict()["spat_encoder.patch_embed.3.weight"])
    assert torch.equal(L3.model.state_dict()["spec_encoder.patch_embed.0.weight"], before)
    frozen = [n for n, p in L3.model.named_parameters() if not p.requires_grad]
    assert frozen and all(n.startswith("spat_encoder.") for n in frozen)


@pytest.mark.skipif(ref_shim.reference_root() is None, reason="reference checkout not present (GPU box)")
def test_round_trip_with_the_real_reference(tmp_path):
    rm, rl, rops, ru = ref_shim.load_reference()
    d = str(tmp_path)
    net = rm.SARSSL(sig_shape=(256, 16, 2, 2), pretrain=True, device="cpu")
    ref_learner = rl.STFTLearner(net, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
    ref_learner.cpu()
    ref_learner.save_checkpoint(epoch=7, checkpoints_dir=d, is_best_epoch=True)
    ours = make_learner(9)
    ours.resume_checkpoint(d, from_latest=False)                               # reference -> ours
    assert ours.start_epoch == 8
    for k, v in net.state_dict().items():
        assert torch.equal(v, ours.model.state_dict()[k]), k
    ours2 = make_learner(4)
    ours2.max_score = 1.5
    ours2.save_checkpoint(11, d, is_best_epoch=True)                            # ours -> reference
    ref_learner.resume_checkpoint(d, from_latest=True)
    assert ref_learner.start_epoch == 12 and ref_learner.max_score == 1.5
    for k, v in ours2.model.state_dict().items():
        assert torch.equal(v, net.state_dict()[k]), k
