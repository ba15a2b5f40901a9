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
