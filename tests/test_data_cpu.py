"""CPU: the input pipeline in front of the path - native WAV decode vs scipy / hand-built files, FixMicSigDataset listing and
return convention (code/dataset.py:107-178), batch loader order / sharding / pinned rows."""
import struct

import numpy as np
import pytest
import scipy.io.wavfile
import torch

from sarssl_b200 import data as D
from sarssl_b200._lib import SarsslError


def _write_pcm24(path, x, fs, extensible=False):
    """x float in [-1, 1) (n, nch) -> 24-bit PCM, optionally with a WAVE_FORMAT_EXTENSIBLE header and an odd-sized LIST chunk first."""
    q = np.clip(np.round(x * 8388608.0), -8388608, 8388607).astype(np.int32)
    raw = b"".join(struct.pack("<i", int(v))[:3] for v in q.reshape(-1))
    nch = x.shape[1]
    if extensible:
        fmt = struct.pack("<HHIIHHHHI", 0xFFFE, nch, fs, fs * nch * 3, nch * 3, 24, 22, 24, 0) + struct.pack("<H", 1) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    else:
        fmt = struct.pack("<HHIIHH", 1, nch, fs, fs * nch * 3, nch * 3, 24)
    junk = b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\x00"
    body = b"WAVE" + junk + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(raw)) + raw
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)
    return q.astype(np.float64) / 8388608.0


def test_wav_decode_formats(tmp_path):
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.9, 0.9, size=(1000, 2))
    i16 = np.round(x * 32767).astype(np.int16)
    scipy.io.wavfile.write(tmp_path / "a16.wav", 16000, i16)
    got, fs = D.wav_read(tmp_path / "a16.wav")
    assert fs == 16000 and got.dtype == np.float32 and got.shape == (1000, 2)
    assert np.array_equal(got, (i16.astype(np.float64) / 32768.0).astype(np.float32))            # == soundfile.read(...).astype(float32)
    i32 = np.round(x * 2147483000).astype(np.int32)
    scipy.io.wavfile.write(tmp_path / "a32.wav", 8000, i32)
    got, fs = D.wav_read(tmp_path / "a32.wav")
    assert fs == 8000 and np.allclose(got, i32 / 2147483648.0, atol=1e-7)
    f32 = x.astype(np.float32)
    scipy.io.wavfile.write(tmp_path / "f32.wav", 16000, f32)
    assert np.array_equal(D.wav_read(tmp_path / "f32.wav")[0], f32)
    scipy.io.wavfile.write(tmp_path / "f64.wav", 16000, x)
    assert np.array_equal(D.wav_read(tmp_path / "f64.wav")[0], f32)
    u8 = rng.integers(0, 256, size=(50, 1), dtype=np.uint8)
    scipy.io.wavfile.write(tmp_path / "u8.wav", 16000, u8)
    assert np.array_equal(D.wav_read(tmp_path / "u8.wav")[0], ((u8.astype(np.float32) - 128) / 128))
    for ext in (False, True):
        want = _write_pcm24(tmp_path / "p24.wav", x, 16000, extensible=ext)
        got, fs = D.wav_read(tmp_path / "p24.wav")
        assert D.wav_info(tmp_path / "p24.wav") == (16000, 2, 1000) and np.array_equal(got, want.astype(np.float32))
    # windowed read with zero fill past the end, decode into a caller buffer
    buf = torch.full((300, 2), 7.0)
    got, _ = D.wav_read(tmp_path / "a16.wav", first=900, count=300, out=buf)
    assert got is buf and np.array_equal(buf[:100].numpy(), (i16[900:] / 32768.0).astype(np.float32)) and float(buf[100:].abs().max()) == 0.0
    with pytest.raises(SarsslError):
        D.wav_read(tmp_path / "missing.wav")
    (tmp_path / "bad.wav").write_bytes(b"not a wav file at all")
    with pytest.raises(SarsslError):
        D.wav_info(tmp_path / "bad.wav")


def _make_set(root, n, ns=2048, with_anno=True):
    rng = np.random.default_rng(1)
    sigs = {}
    for i in range(n):
        x = np.round(rng.uniform(-0.5, 0.5, size=(ns, 2)) * 32767).astype(np.int16)
        scipy.io.wavfile.write(root / f"clip{i:02d}.wav", 16000, x)
        scipy.io.wavfile.write(root / f"clip{i:02d}_dp.wav", 16000, (x // 2).astype(np.int16))
        if with_anno:
            np.savez(root / f"clip{i:02d}_info.npz", room_sz=np.array([4.0, 5.0, 3.0]), TDOA=np.float64(i * 1e-4), T60_edc=np.float64(0.3 + 0.01 * i),
                     DRR=np.float64(1.0), C50=np.float64(2.0))
        sigs[f"clip{i:02d}.wav"] = (x / 32768.0).astype(np.float32)
    return sigs


def test_dataset_matches_reference_conventions(tmp_path):
    sigs = _make_set(tmp_path, 5)
    ds = D.FixMicSigDataset(str(tmp_path), fs=16000, load_anno=True, dataset_sz=None, load_dp=True)
    assert len(ds) == 5 and all(not f.name.endswith("_dp.wav") for f in ds.files)            # dp copies are not items (dataset.py:129)
    for i in range(len(ds)):
        sig, anno, dp = ds[i]
        name = ds.files[i].name
        assert sig.dtype == np.float32 and np.array_equal(sig, sigs[name]) and dp.shape == sig.shape
        k = int(name[4:6])
        assert set(anno) == {"TDOA", "T60", "DRR", "C50", "ABS"} and anno["TDOA"].dtype == np.float32 and abs(float(anno["TDOA"]) - k * 1e-4) < 1e-9
        vol, sur = 60.0, 4 * 5 + 4 * 3 + 5 * 3
        assert abs(float(anno["ABS"]) - 0.161 * vol / sur / (0.3 + 0.01 * k)) < 1e-6
    assert len(D.FixMicSigDataset(str(tmp_path), 16000, False, dataset_sz=3)) == 3
    ds8 = D.FixMicSigDataset(str(tmp_path), fs=8000, load_anno=False, dataset_sz=None)       # resampled on the host like the reference
    assert ds8[0][0].shape == (1024, 2)
    ds_t = D.FixMicSigDataset(str(tmp_path), 16000, False, None, transforms=[lambda s: s[:100] * 2])
    assert ds_t[0][0].shape == (100, 2)


def test_batch_loader_order_sharding_and_rows(tmp_path):
    sigs = _make_set(tmp_path, 10)
    ds = D.FixMicSigDataset(str(tmp_path), fs=16000, load_anno=True, dataset_sz=None)
    ld = D.WaveformBatchLoader(ds, batch_size=4, shuffle=False, num_workers=3, pin_memory=False)
    batches = list(ld)
    assert len(ld) == 3 and [b[0].shape[0] for b in batches] == [4, 4, 2]
    flat = torch.cat([b[0] for b in batches])
    for i in range(10):
        assert np.array_equal(flat[i].numpy(), sigs[ds.files[i].name])
    assert batches[0][1]["TDOA"].shape == (4,) and batches[0][1]["TDOA"].dtype == torch.float32
    # shuffle: reproducible per (seed, epoch), different between epochs, a permutation of the data set
    ld = D.WaveformBatchLoader(ds, 5, shuffle=True, seed=3, pin_memory=False, drop_last=True)
    a = torch.cat([b[0] for b in ld]); b_ = torch.cat([b[0] for b in ld])
    ld.set_epoch(1)
    c = torch.cat([b[0] for b in ld])
    assert torch.equal(a, b_) and not torch.equal(a, c) and sorted(float(v) for v in a.sum((1, 2))) == sorted(float(v) for v in c.sum((1, 2)))
    # data-parallel: the ranks' batches together are the global batches of the single-process loader
    whole = [b[0] for b in D.WaveformBatchLoader(ds, 4, shuffle=True, seed=5, pin_memory=False, drop_last=True)]
    parts = [[b[0] for b in D.WaveformBatchLoader(ds, 2, shuffle=True, seed=5, pin_memory=False, rank=r, world=2)] for r in range(2)]
    assert len(parts[0]) == len(whole) == 2
    for g, w in enumerate(whole):
        assert torch.equal(torch.cat([parts[0][g], parts[1][g]]), w)
    # crop / pad to a fixed length, early exit of the consumer does not hang the producer
    ld = D.WaveformBatchLoader(ds, 3, pin_memory=False, nsample=3000)
    first = next(iter(ld))
    assert first[0].shape == (3, 3000, 2) and float(first[0][:, 2048:].abs().max()) == 0.0
    scipy.io.wavfile.write(tmp_path / "short.wav", 16000, np.zeros((100, 2), dtype=np.int16))
    with pytest.raises(SarsslError):
        list(D.WaveformBatchLoader(D.FixMicSigDataset(str(tmp_path), 16000, False, None), 16, pin_memory=False))


def test_native_batch_reader_matches_per_file_reads_and_reports_the_bad_file(tmp_path):
    """sarssl_wav_read_batch_f32 (one call, native threads) == sarssl_wav_read_f32 per file; mismatching files fail loudly."""
    rng = np.random.default_rng(3)
    paths = []
    for i in range(9):
        p = tmp_path / f"c{i}.wav"
        scipy.io.wavfile.write(p, 16000, np.round(rng.uniform(-0.9, 0.9, size=(700, 2)) * 32767).astype(np.int16))
        paths.append(p)
    out = torch.empty(9, 700, 2)
    D.wav_read_batch(paths, 700, 2, out, fs=16000, exact=True, nthreads=4)
    for i, p in enumerate(paths):
        assert np.array_equal(out[i].numpy(), D.wav_read(p)[0])
    longer = torch.empty(9, 800, 2)                                     # crop / zero-pad mode
    D.wav_read_batch(paths, 800, 2, longer, fs=16000, exact=False, nthreads=3)
    assert torch.equal(longer[:, :700], out) and float(longer[:, 700:].abs().max()) == 0.0
    scipy.io.wavfile.write(tmp_path / "short.wav", 16000, np.zeros((650, 2), dtype=np.int16))
    with pytest.raises(SarsslError, match="short.wav"):
        D.wav_read_batch(paths[:4] + [tmp_path / "short.wav"] + paths[4:], 700, 2, torch.empty(10, 700, 2), fs=16000, exact=True, nthreads=4)
    with pytest.raises(SarsslError, match="8000|16000"):
        D.wav_read_batch(paths, 700, 2, out, fs=8000, exact=True)
    ds = D.FixMicSigDataset(data_dir=str(tmp_path), fs=16000, load_anno=False, dataset_sz=None)
    ds.files = paths
    ds.dataset_sz = len(paths)
    got = [b[0] for b in D.WaveformBatchLoader(ds, batch_size=4, num_workers=3, pin_memory=False)]
    assert [tuple(b.shape) for b in got] == [(4, 700, 2), (4, 700, 2), (1, 700, 2)] and torch.equal(torch.cat(got), out)
