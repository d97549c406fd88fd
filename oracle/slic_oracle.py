"""CPU oracle for favit_slic_segment (GPU SLIC).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's segmentation step is `skimage.segmentation.slic(image, n_segments, compactness=0.1, sigma=1.0,
start_label=0)` (/root/reference/models/sppp.py:61-73).  scikit-image is NOT installed in this image and the
reference pins no version and ships no fixture, so this oracle restates the algorithm as include/favit.h documents it
(Achanta et al.'s SLIC as scikit-image structures it: Gaussian pre-smoothing, regular grid of centres, k-means rounds
with the distance colour^2 / compactness^2 + space^2 / step^2) — PARITY UNPINNED against scikit-image; it pins the CUDA
kernel bit for bit: every float32 operation is done in the same order with numpy float32 scalars / arrays.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def slic_grid(H: int, W: int, n_segments: int):
    gx = int(round(math.sqrt(n_segments * W / H)))
    gx = max(1, min(W, gx))
    gy = int(round(n_segments / gx))
    gy = max(1, min(H, gy))
    return gy, gx


def _blur(img: np.ndarray, sigma: float) -> np.ndarray:
    """img [C,H,W] float32; separable, 'nearest' borders, taps in ascending offset order, x pass then y pass."""
    radius = min(8, int(4.0 * sigma + 0.5))
    if sigma <= 0 or radius <= 0:
        return img.copy()
    w = [math.exp(-0.5 * t * t / (sigma * sigma)) for t in range(-radius, radius + 1)]
    tot = sum(w)
    w = [F(v / tot) for v in w]
    C, H, W = img.shape
    out = img
    for axis in (2, 1):
        n = out.shape[axis]
        acc = np.zeros_like(out)
        for t in range(-radius, radius + 1):
            idx = np.clip(np.arange(n) + t, 0, n - 1)
            acc = (acc + w[t + radius] * np.take(out, idx, axis=axis)).astype(F)
        out = acc
    return out


def slic_oracle(images: np.ndarray, n_segments: int, compactness: float = 0.1, sigma: float = 1.0, iters: int = 10):
    """images float32 [B,C,H,W] -> int64 [B,H,W]."""
    images = np.asarray(images, dtype=F)
    B, C, H, W = images.shape
    gy, gx = slic_grid(H, W, n_segments)
    K = gy * gx
    step_y, step_x = F(H) / F(gy), F(W) / F(gx)
    step = max(step_y, step_x)
    inv_sy, inv_sx = F(1) / step_y, F(1) / step_x
    inv_step2 = F(1) / (step * step)
    inv_comp2 = F(1) / (F(compactness) * F(compactness))
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    cyi = np.minimum((ys.astype(F) * inv_sy).astype(np.int64), gy - 1)
    cxi = np.minimum((xs.astype(F) * inv_sx).astype(np.int64), gx - 1)
    out = np.empty((B, H, W), dtype=np.int64)
    for b in range(B):
        f = _blur(images[b], sigma)                                  # [C,H,W]
        cen = np.empty((K, 2 + C), dtype=F)
        for k in range(K):
            iy, ix = divmod(k, gx)
            cy, cx = (F(iy) + F(0.5)) * step_y, (F(ix) + F(0.5)) * step_x
            py, px = min(int(cy), H - 1), min(int(cx), W - 1)
            cen[k, 0], cen[k, 1] = cy, cx
            cen[k, 2:] = f[:, py, px]
        lab = np.zeros((H, W), dtype=np.int64)
        for it in range(iters + 1):
            best = np.full((H, W), F(3.402823466e38), dtype=F)
            lab = np.zeros((H, W), dtype=np.int64)
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    iy, ix = cyi + dy, cxi + dx
                    ok = (iy >= 0) & (iy < gy) & (ix >= 0) & (ix < gx)
                    k = np.where(ok, iy * gx + ix, 0)
                    c = cen[k]                                      # [H,W,2+C]
                    dc = np.zeros((H, W), dtype=F)
                    for ch in range(C):
                        d = (f[ch] - c[:, :, 2 + ch]).astype(F)
                        dc = (dc + (d * d).astype(F)).astype(F)
                    ey, ex = (ys.astype(F) - c[:, :, 0]).astype(F), (xs.astype(F) - c[:, :, 1]).astype(F)
                    ds = ((ey * ey).astype(F) + (ex * ex).astype(F)).astype(F)
                    dist = ((dc * inv_comp2).astype(F) + (ds * inv_step2).astype(F)).astype(F)
                    better = ok & (dist < best)
                    best = np.where(better, dist, best)
                    lab = np.where(better, k, lab)
            if it == iters:
                break
            flat = lab.reshape(-1)
            n = np.bincount(flat, minlength=K)
            sy = np.bincount(flat, weights=ys.reshape(-1).astype(np.float64), minlength=K)
            sx = np.bincount(flat, weights=xs.reshape(-1).astype(np.float64), minlength=K)
            for k in range(K):
                if n[k] == 0:
                    continue
                cen[k, 0] = F(sy[k] / n[k])
                cen[k, 1] = F(sx[k] / n[k])
            for ch in range(C):
                fx = np.rint((f[ch] * F(65536.0)).astype(F)).astype(np.int64).reshape(-1)
                s = np.zeros(K, dtype=np.int64)
                np.add.at(s, flat, fx)
                for k in range(K):
                    if n[k]:
                        cen[k, 2 + ch] = F(float(s[k]) / float(n[k]) / 65536.0)
        out[b] = lab
    return out
