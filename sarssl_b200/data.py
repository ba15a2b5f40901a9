"""Input pipeline in front of the hot path: presaved microphone signals -> pinned host batches.

Mirrors the reference's `FixMicSigDataset` (code/dataset.py:107-178) and the `torch.utils.data.DataLoader` wiring of
run_pretrain.py:191-199, re-designed for the GPU path:

  * WAV files are decoded by the native reader (`sarssl_wav_read_f32`, csrc/wav.cpp) straight into the batch's PINNED host
    buffer - no per-item numpy array, no collate copy; the reader runs in worker THREADS (ctypes releases the GIL), so no
    process pool and no pickling of 1 MB items;
  * batches are produced `prefetch` ahead by a background thread and handed to `Learner.device_batches`, which copies
    batch i+1 to the device on a side stream while batch i trains.

`FixMicSigDataset.__getitem__` keeps the reference's return convention ([mic_sig float32 (nsample, nch)] (+ annotation dict)
(+ direct-path signal)); resampling (`fs` mismatch) and user transforms fall back to scipy / the callables, like the reference.
"""
import ctypes as C
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import torch

from ._lib import SarsslError, check, lib


def wav_info(path):
    """(fs, nch, nsample) of a RIFF/WAVE file."""
    fs, nch, ns = C.c_int(0), C.c_int(0), C.c_longlong(0)
    check(lib().sarssl_wav_info(str(path).encode(), C.byref(fs), C.byref(nch), C.byref(ns)), "sarssl_wav_info")
    return fs.value, nch.value, ns.value


def wav_read(path, first=0, count=None, out=None):
    """Decode frames [first, first+count) as float32 (count, nch), scaled like soundfile.read (int16 / 2^15, ...).
    `out`: optional C-contiguous float32 array / CPU tensor to decode into (e.g. a row of a pinned batch); frames past the end
    of the file are zero-filled.  Returns (array, fs)."""
    fs, nch, ns = wav_info(path)
    if count is None:
        count = max(ns - first, 0)
    if out is None:
        out = np.empty((count, nch), dtype=np.float32)
    if torch.is_tensor(out):
        if out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != count * nch:
            raise SarsslError("wav_read: `out` must be a contiguous float32 tensor of count * nch elements")
        addr = out.data_ptr()
    else:
        if out.dtype != np.float32 or not out.flags.c_contiguous or out.size != count * nch:
            raise SarsslError("wav_read: `out` must be a C-contiguous float32 array of count * nch elements")
        addr = out.ctypes.data
    nread = C.c_longlong(0)
    check(lib().sarssl_wav_read_f32(str(path).encode(), first, count, C.c_void_p(addr), C.byref(nread)), "sarssl_wav_read_f32")
    return out, fs


def wav_read_batch(paths, count, nch, out, fs=0, first=0, exact=False, nthreads=8):
    """Decode `paths[i]` frames [first, first+count) into out[i] (float32 (len(paths), count, nch), e.g. a pinned batch) with `nthreads`
    native threads in ONE library call (no per-clip Python)."""
    if out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != (len(paths), count, nch):
        raise SarsslError("wav_read_batch: `out` must be a contiguous float32 tensor (len(paths), count, nch)")
    arr = (C.c_char_p * len(paths))(*[str(p).encode() for p in paths])
    bad = C.c_int(-1)
    check(lib().sarssl_wav_read_batch_f32(C.cast(arr, C.c_void_p), len(paths), first, count, nch, int(fs), int(exact), C.c_void_p(out.data_ptr()),
                                          int(nthreads), C.byref(bad)), "sarssl_wav_read_batch_f32")
    return out


class FixMicSigDataset:
    """dataset.py:107-178: presaved microphone signals (`*.wav`, direct-path copies `*_dp.wav`, annotations `*_info.npz`)."""

    def __init__(self, data_dir, fs, load_anno, dataset_sz, load_dp=False, transforms=None):
        if isinstance(data_dir, (list, tuple)):
            files, dp_files = [], []
            for d in data_dir:
                files += list(Path(d).rglob("*.wav"))
                dp_files += list(Path(d).rglob("*_dp.wav"))
            np.random.shuffle(files)                                   # like the reference (dataset.py:125)
        else:
            files = list(Path(data_dir).rglob("*.wav"))
            dp_files = list(Path(data_dir).rglob("*_dp.wav"))
        dp = set(dp_files)
        self.files = [f for f in files if f not in dp]
        self.dataset_sz = len(self.files) if dataset_sz is None else int(min(len(self.files), dataset_sz))
        self.fs, self.load_anno, self.load_dp, self.transforms = fs, load_anno, load_dp, transforms

    def __len__(self):
        return self.dataset_sz

    def _signal(self, file_name, out=None):
        fs, nch, ns = wav_info(file_name)
        direct = out is not None and fs == self.fs and self.transforms is None and out.shape[0] == ns
        sig, _ = wav_read(file_name, out=out if direct else None)
        if fs != self.fs:
            import scipy.signal
            sig = scipy.signal.resample_poly(sig, self.fs, fs).astype(np.float32)
        if self.transforms is not None:
            for t in self.transforms:
                sig = t(sig)
        return sig if torch.is_tensor(sig) else np.asarray(sig, dtype=np.float32)

    def annotations(self, idx):
        """dataset.py:156-168."""
        info = dict(np.load(str(self.files[idx]).replace(".wav", "_info.npz")))
        rs = info["room_sz"]
        vol = rs[0] * rs[1] * rs[2]
        sur = rs[0] * rs[1] + rs[0] * rs[2] + rs[1] * rs[2]
        return {"TDOA": info["TDOA"].astype(np.float32), "T60": info["T60_edc"].astype(np.float32), "DRR": info["DRR"].astype(np.float32),
                "C50": info["C50"].astype(np.float32), "ABS": np.array(0.161 * vol / sur / info["T60_edc"]).astype(np.float32)}

    def __getitem__(self, idx):
        file_name = str(self.files[idx])
        ret = [self._signal(file_name)]
        if self.load_anno:
            ret += [self.annotations(idx)]
        if self.load_dp:
            ret += [self._signal(file_name.replace(".wav", "_dp.wav"))]
        return ret


class WaveformBatchLoader:
    """DataLoader replacement for `FixMicSigDataset` (run_pretrain.py:191-199 uses batch_size, shuffle, num_workers, pin_memory).

    Yields `[sig_batch]` (pre-training) or `[sig_batch, {task: labels}]` (fine-tuning) with `sig_batch` a pinned float32 tensor
    (B, nsample, nch): every worker thread decodes its WAV straight into its row of the batch.  `prefetch` batches are built ahead
    by a background thread.  All clips must have the same length and channel count (the reference's default collate has the same
    requirement); `nsample` may be given to crop / zero-pad every clip instead.  `seed` + `set_epoch` make the shuffle
    reproducible and identical on every data-parallel rank; `rank` / `world` select this rank's share of each global batch."""

    def __init__(self, dataset, batch_size, shuffle=False, num_workers=8, pin_memory=True, drop_last=False, prefetch=2, seed=0, nsample=None,
                 rank=0, world=1):
        self.ds, self.bs, self.shuffle, self.drop_last = dataset, int(batch_size), shuffle, drop_last
        self.workers, self.pin, self.prefetch, self.seed, self.nsample = max(int(num_workers), 1), pin_memory, max(int(prefetch), 1), seed, nsample
        self.rank, self.world, self.epoch = rank, world, 0

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        n = len(self._order())                        # same truncation as iteration: whole global batches only when world > 1
        return n // self.bs if self.drop_last else (n + self.bs - 1) // self.bs

    def _order(self):
        idx = np.arange(len(self.ds))
        if self.shuffle:
            np.random.default_rng(self.seed + self.epoch).shuffle(idx)
        if self.world > 1:                                  # every rank walks the same permutation and keeps its slice of each global batch
            gb = self.bs * self.world
            idx = idx[: len(idx) // gb * gb].reshape(-1, self.world, self.bs)[:, self.rank].reshape(-1)
        return idx

    def _build(self, ids, pool):
        ds = self.ds
        fs0, nch, ns0 = wav_info(ds.files[ids[0]])
        ns = self.nsample if self.nsample is not None else ns0
        plain = ds.transforms is None and not ds.load_dp
        # pinned straight from torch's caching host allocator: after the first few batches no cudaHostAlloc and no staging copy happen
        batch = torch.empty((len(ids), ns, nch), dtype=torch.float32, pin_memory=bool(self.pin and torch.cuda.is_available()))
        if plain and not ds.load_anno:
            # the pre-training case (run_pretrain.py:191-199): the whole batch in one native call, `workers` decoder threads, no per-clip Python
            wav_read_batch([ds.files[i] for i in ids], ns, nch, batch, fs=ds.fs, exact=self.nsample is None, nthreads=self.workers)
            return [batch]

        def one(j):
            path = ds.files[ids[j]]
            fs, c, n = wav_info(path)
            if c != nch:
                raise SarsslError(f"{path}: {c} channels, the batch has {nch}")
            if plain and fs == ds.fs:
                if self.nsample is None and n != ns:
                    raise SarsslError(f"{path}: {n} samples, the batch has {ns} (pass nsample= to crop / pad)")
                wav_read(path, 0, ns, out=batch[j])          # decode straight into the pinned row (zero-filled past the end)
            else:
                sig = torch.as_tensor(ds._signal(str(path)))
                m = min(ns, sig.shape[0])
                batch[j, :m] = sig[:m]
                batch[j, m:] = 0
            return ds.annotations(ids[j]) if ds.load_anno else None

        annos = list(pool.map(one, range(len(ids))))
        ret = [batch]
        if ds.load_anno:
            keys = annos[0].keys()
            ret.append({k: torch.from_numpy(np.stack([np.asarray(a[k], dtype=np.float32) for a in annos])) for k in keys})      # stacked like default_collate
        if ds.load_dp:                                 # direct-path signals as the last item, like FixMicSigDataset.__getitem__ (dataset.py:170-176)
            dp = torch.empty_like(batch)

            def one_dp(j):
                sig = torch.as_tensor(ds._signal(str(ds.files[ids[j]]).replace(".wav", "_dp.wav")))
                m = min(ns, sig.shape[0])
                dp[j, :m] = sig[:m]
                dp[j, m:] = 0

            list(pool.map(one_dp, range(len(ids))))
            ret.append(dp)
        return ret

    def __iter__(self):
        order = self._order()
        chunks = [order[i:i + self.bs] for i in range(0, len(order), self.bs)]
        if self.drop_last and chunks and len(chunks[-1]) < self.bs:
            chunks.pop()
        q = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def producer():
            try:
                with ThreadPoolExecutor(self.workers) as pool:
                    for ids in chunks:
                        if stop.is_set():
                            return
                        q.put(self._build(ids, pool))
                q.put(None)
            except BaseException as e:                     # surface decode errors in the consumer
                q.put(e)

        th = threading.Thread(target=producer, daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            while th.is_alive():                            # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    th.join(timeout=0.05)
