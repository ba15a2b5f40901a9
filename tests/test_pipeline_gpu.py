"""GPU: the input pipeline step in front of the path (Learner.device_batches): side-stream H2D prefetch keeps order, structure and values."""
import pytest
import torch

from sarssl_b200.learner import Learner

pytestmark = pytest.mark.gpu


def test_device_batches_order_structure_values():
    L = Learner(None)
    L.device = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(3)
    host = []
    for i in range(7):
        sig = torch.randn(4, 4096, 2, generator=g)
        if i % 2 == 0:
            sig = sig.pin_memory()
        host.append([sig, {"TDOA": torch.full((4,), float(i)), "name": f"item{i}"}])
    busy = torch.randn(2048, 2048, device="cuda")
    got = []
    for sig, ann in L.device_batches(iter(host)):
        busy = busy @ busy * 1e-3                                  # device work queued while the next copy runs
        assert sig.is_cuda and ann["TDOA"].is_cuda and isinstance(ann["name"], str)
        got.append((sig.clone(), ann["TDOA"].clone(), ann["name"]))
    torch.cuda.synchronize()
    assert len(got) == 7
    for i, (sig, lab, name) in enumerate(got):
        assert torch.equal(sig.cpu(), host[i][0]) and torch.equal(lab.cpu(), host[i][1]["TDOA"]) and name == f"item{i}"
    assert list(L.device_batches([])) == []
    dev_item = [torch.ones(3, device="cuda")]
    out = list(L.device_batches([dev_item]))
    assert out[0][0].data_ptr() == dev_item[0].data_ptr()          # device tensors pass through untouched


def test_wav_loader_feeds_pretrain_epoch(tmp_path):
    """WAV files -> WaveformBatchLoader (pinned rows decoded by the native reader) -> pretrain_epoch == the same tensors passed directly."""
    import random

    import numpy as np
    import scipy.io.wavfile

    from sarssl_b200 import data as D
    from sarssl_b200.learner import STFTLearner
    from sarssl_b200.model import SARSSL

    rng = np.random.default_rng(0)
    for i in range(6):
        scipy.io.wavfile.write(tmp_path / f"c{i}.wav", 16000, np.round(rng.uniform(-0.3, 0.3, size=(17 * 256, 2)) * 32767).astype(np.int16))
    ds = D.FixMicSigDataset(str(tmp_path), fs=16000, load_anno=False, dataset_sz=None)
    loader = D.WaveformBatchLoader(ds, batch_size=3, shuffle=False, num_workers=2)
    direct = [[torch.stack([torch.from_numpy(ds[i][0]) for i in range(j, j + 3)])] for j in (0, 3)]
    assert loader.__len__() == 2 and next(iter(loader))[0].is_pinned()

    def run(dataset):
        torch.manual_seed(0)
        net = SARSSL(sig_shape=(256, 16, 2, 2), device="cuda:0")
        net.to("cuda:0")
        net.set_dropout(0.0)
        L = STFTLearner(net, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
        L.device = torch.device("cuda", 0)
        random.seed(5)
        return L.pretrain_epoch(dataset, lr=1e-4, epoch=1)[:2]

    a, b = run(loader), run(direct)
    assert abs(a[0] - b[0]) <= 1e-6 * abs(b[0]) and abs(a[1] - b[1]) <= 1e-6 * abs(b[1])
