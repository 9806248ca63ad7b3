"""Metrics and image output AFTER the sampler path, off the critical path.

Replaces what sample_image() and hmc() do with the samples in the reference:
  * main_sampling.py:735-742 -- a PNG of every accepted epoch + its PSNR, written synchronously inside the HMC loop
    (device -> host copy, PIL encode and file write between two leapfrog trajectories);
  * main_sampling.py:488-519 -- inverse_data_transform, the pixel-wise std-dev map over the collected samples
    (std over samples, mean over channels, min-max normalised, 'hot' colour map) and per-sample PSNR with its mean and
    sample standard deviation (ddof = 1).  SSIM / LPIPS need skimage / lpips, which are not vendored with the reference.

Here the PSNR values and the std-dev map are computed on the device in a handful of launches, and images leave through
`AsyncImageWriter`: the copy to pinned host memory is queued on a side stream behind an event, a worker thread waits for
that event, encodes the PNG and writes the file -- the sampler's stream never waits for any of it.
"""
import os
import queue
import struct
import threading
import zlib

import numpy as np
import torch


def inverse_data_transform(x):
    """datasets/__init__.py:214-223 for the FFHQ config: [-1, 1] -> [0, 1], clamped."""
    return ((x + 1.0) / 2.0).clamp(0.0, 1.0)


def psnr_per_sample(samples, x_orig):
    """main_sampling.py:517-519: 10 log10(1 / mse) of every sample against the ground truth, both mapped to [0, 1].
    samples: [..., C, H, W] in [-1, 1]; x_orig: [C, H, W] (or broadcastable).  Returns a device tensor [...]."""
    img = inverse_data_transform(samples.float())
    orig = inverse_data_transform(x_orig.float())
    mse = ((img - orig) ** 2).mean(dim=(-3, -2, -1))
    return 10.0 * torch.log10(1.0 / mse)


def std_dev_map(samples):
    """main_sampling.py:494-497: x.std(dim=0) (unbiased), mean over channels, min-max normalised.  samples: [S, C, H, W] in
    [-1, 1] (S >= 2).  Returns (normalised map [H, W], raw channel-mean std [H, W]) on the device."""
    x = inverse_data_transform(samples.float())
    std = x.std(dim=0)
    std_mean = std.mean(dim=0)
    lo, hi = std_mean.min(), std_mean.max()
    return (std_mean - lo) / (hi - lo), std_mean


def hot_colormap(v):
    """matplotlib's 'hot' colour map (the reference's `cmap='hot'`, main_sampling.py:500) for v in [0, 1]: [H, W] -> [3, H, W]."""
    v = v.clamp(0.0, 1.0)
    r = (v / 0.365079).clamp(0.0, 1.0)
    g = ((v - 0.365079) / (0.746032 - 0.365079)).clamp(0.0, 1.0)
    b = ((v - 0.746032) / (1.0 - 0.746032)).clamp(0.0, 1.0)
    return torch.stack([r, g, b])


def summarize(samples, x_orig):
    """Per-image summary the reference prints (main_sampling.py:510-551): PSNR of every sample, their mean and sample
    standard deviation (ddof = 1, only for S > 1).  One host read at the end."""
    ps = psnr_per_sample(samples, x_orig)
    out = {"psnr": ps.cpu().tolist(), "psnr_mean": float(ps.mean())}
    out["psnr_std"] = float(ps.std(unbiased=True)) if ps.numel() > 1 else 0.0
    return out


# ----------------------------------------------------------------------------------------------------------------
# PNG output without blocking the sampler
# ----------------------------------------------------------------------------------------------------------------
def encode_png(u8):
    """Minimal PNG encoder (8-bit RGB or grey, filter 0, zlib): u8 is a numpy uint8 array [H, W, 3] or [H, W]."""
    if u8.ndim == 2:
        u8 = u8[:, :, None]
    h, w, c = u8.shape
    assert c in (1, 3) and u8.dtype == np.uint8
    raw = np.concatenate([np.zeros((h, 1), np.uint8), u8.reshape(h, w * c)], axis=1).tobytes()

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    ihdr = struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 0, 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(raw, 3)) + chunk(b"IEND", b"")


def to_uint8_hwc(img01):
    """torchvision.utils.save_image's quantisation (mul 255, add 0.5, clamp, uint8), [C, H, W] in [0, 1] -> [H, W, C]."""
    return img01.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8)


class AsyncImageWriter:
    """save(img, path): img is a device tensor [C, H, W] in [0, 1] (or [H, W]).  The uint8 conversion is one small kernel on
    the caller's stream; the device -> pinned-host copy runs on this object's side stream behind an event; a worker
    thread waits for the copy, encodes and writes.  close() drains the queue (call once at the end of a run)."""

    def __init__(self, device, max_pending=64):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.q = queue.Queue(maxsize=max_pending)
        self.errors = []
        self.written = 0
        self._pool = {}
        self._lock = threading.Lock()
        self.worker = threading.Thread(target=self._run, daemon=True)
        self.worker.start()

    def _host_buffer(self, shape):
        with self._lock:
            free = self._pool.setdefault(shape, [])
            return free.pop() if free else torch.empty(shape, dtype=torch.uint8).pin_memory()

    def save(self, img01, path):
        if img01.dim() == 2:
            img01 = img01[None]
        u8 = to_uint8_hwc(img01.detach().float())                      # caller's stream
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        host = self._host_buffer(tuple(u8.shape))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            host.copy_(u8, non_blocking=True)
            u8.record_stream(self.stream)
            done = torch.cuda.Event()
            done.record(self.stream)
        self.q.put((host, done, path))

    def _run(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            host, done, path = item
            try:
                done.synchronize()
                arr = host.numpy()
                data = encode_png(arr[:, :, 0] if arr.shape[2] == 1 else arr)
                os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
                with open(path, "wb") as f:
                    f.write(data)
                self.written += 1
            except Exception as e:          # noqa: BLE001  -- reported at close()
                self.errors.append((path, repr(e)))
            finally:
                with self._lock:
                    self._pool.setdefault(tuple(host.shape), []).append(host)

    def close(self):
        self.q.put(None)
        self.worker.join()
        if self.errors:
            raise RuntimeError("AsyncImageWriter: %d image(s) failed: %s" % (len(self.errors), self.errors[:3]))
        return self.written
