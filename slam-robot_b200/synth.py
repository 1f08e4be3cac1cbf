"""Seeded synthetic inputs for tests and bench: textured BGR frames related by a
small affine motion, feature points, and 256-bit descriptors.  torch is used as plumbing only
(works on CPU and on CUDA); consumers always receive plain uint8 / float32 arrays, so the CPU
oracle and the CUDA path see identical bytes.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

PAD = 16


def _canvas(gen, n, H, W, device, cell=8, texture=30.0):
    """n smooth-plus-texture canvases of (H+2*PAD, W+2*PAD), float32 in [0,255], 3 channels."""
    Hc, Wc = H + 2 * PAD, W + 2 * PAD
    gh, gw = Hc // cell + 2, Wc // cell + 2
    low = torch.rand((n, 1, gh, gw), generator=gen, device=device) * 255.0
    smooth = F.interpolate(low, size=(Hc, Wc), mode="bicubic", align_corners=False)
    tex = (torch.rand((n, 1, Hc, Wc), generator=gen, device=device) * 2 - 1) * texture
    gray = smooth * 0.6 + 50.0 + tex
    tint = torch.tensor([0.95, 1.0, 1.05], device=device).view(1, 3, 1, 1)
    return (gray * tint).clamp_(0, 255)


def _warp_crop(canvas, angle, shift, H, W):
    """Sample the canvas under x' = R(angle) (x - c) + c + shift (bilinear), crop PAD."""
    n, _, Hc, Wc = canvas.shape
    dev = canvas.device
    ys, xs = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32),
                            torch.arange(W, device=dev, dtype=torch.float32), indexing="ij")
    cx, cy = (W - 1) * 0.5, (H - 1) * 0.5
    ca, sa = math.cos(angle), math.sin(angle)
    # frame-B pixel (x,y) shows canvas point A^-1: p_a = R^T (p_b - c - shift) + c
    dx, dy = xs - cx - shift[0], ys - cy - shift[1]
    ax = ca * dx + sa * dy + cx + PAD
    ay = -sa * dx + ca * dy + cy + PAD
    gx = (ax + 0.5) / Wc * 2 - 1
    gy = (ay + 0.5) / Hc * 2 - 1
    grid = torch.stack([gx, gy], -1).unsqueeze(0).expand(n, -1, -1, -1)
    return F.grid_sample(canvas, grid, mode="bilinear", padding_mode="border", align_corners=False)


def true_motion(pts, H, W, angle, shift):
    """Where a frame-A point lands in frame B under the synthetic motion."""
    pts = np.asarray(pts, np.float64).reshape(-1, 2)
    cx, cy = (W - 1) * 0.5, (H - 1) * 0.5
    ca, sa = math.cos(angle), math.sin(angle)
    dx, dy = pts[:, 0] - cx, pts[:, 1] - cy
    return np.stack([ca * dx - sa * dy + cx + shift[0], sa * dx + ca * dy + cy + shift[1]], 1)


def make_pairs(seed, n, H, W, device="cpu", angle=0.004, shift=(1.7, -2.3)):
    """n independent frame pairs. Returns (A, B) uint8 tensors of shape (n, H, W, 3), BGR."""
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    canvas = _canvas(gen, n, H, W, device)
    A = canvas[:, :, PAD:PAD + H, PAD:PAD + W]
    B = _warp_crop(canvas, angle, shift, H, W)
    to_u8 = lambda t: t.round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    return to_u8(A), to_u8(B)


def make_sequence(seed, nframes, H, W, stride=2, device="cpu", angle=0.004, shift=(1.7, -2.3), period=6, noise=2.0):
    """A replayed alternating-camera sequence as `./slam --load` reads it (main.cpp:503-519: `stride` cameras take
    turns, so frames i and i + stride belong to the same camera).  Every camera looks at its own scene, which moves
    back and forth: frame i shows it after m = tri(i // stride) steps of the synthetic motion (a triangle wave of
    `period` frames, so the view never leaves the canvas margin), i.e. consecutive frames of a camera are one motion
    step apart, forwards or backwards.  Per-frame sensor noise makes every frame unique.  uint8 (nframes, H, W, 3)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    canvas = _canvas(gen, stride, H, W, device)
    half = period // 2
    out = torch.empty((nframes, H, W, 3), dtype=torch.uint8, device=device)
    for c in range(stride):
        for m in range(half + 1):
            idx = [i for i in range(c, nframes, stride) if half - abs((i // stride) % period - half) == m]
            if not idx:
                continue
            fr = _warp_crop(canvas[c:c + 1], angle * m, (shift[0] * m, shift[1] * m), H, W)
            for j0 in range(0, len(idx), 64):
                sub = idx[j0:j0 + 64]
                nz = (torch.rand((len(sub), 1, H, W), generator=gen, device=device) * 2 - 1) * noise
                out[sub] = (fr + nz).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1)
    return out


def make_frames(seed, n, H, W, device="cpu"):
    """n unrelated textured frames, uint8 (n, H, W, 3)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    canvas = _canvas(gen, n, H, W, device)
    return canvas[:, :, PAD:PAD + H, PAD:PAD + W].round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def make_features(seed, n, H, W, margin=16.0, border_frac=0.0):
    """n seeded feature points (x,y) float32; a fraction may be placed in the `margin` band next to
    the image edges to exercise the clipping / replicate paths of GetPatch."""
    rng = np.random.default_rng(seed)
    pts = np.empty((n, 2), np.float32)
    pts[:, 0] = rng.uniform(margin, W - margin, n)
    pts[:, 1] = rng.uniform(margin, H - margin, n)
    nb = int(n * border_frac)
    if nb:
        side = rng.integers(0, 4, nb)
        t = rng.uniform(0.5, margin, nb).astype(np.float32)
        u = rng.uniform(0.5, 1.0, nb)
        bx = np.where(side == 0, t, np.where(side == 1, W - t, u * (W - 1)))
        by = np.where(side == 2, t, np.where(side == 3, H - t, u * (H - 1)))
        by = np.where(side < 2, rng.uniform(0.5, H - 0.5, nb), by)
        bx = np.where(side >= 2, rng.uniform(0.5, W - 0.5, nb), bx)
        pts[:nb, 0], pts[:nb, 1] = bx, by
    return pts


def make_descriptors(seed, n, dup_frac=0.0, source=None):
    """n seeded 256-bit descriptors as uint32 (n, 8); optionally plants copies of rows of
    `source` (ties / exact matches)."""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 2 ** 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    nd = int(n * dup_frac)
    if nd:
        src = d if source is None else np.asarray(source, np.uint32).reshape(-1, 8)
        rows = rng.integers(0, n, nd)
        pick = rng.integers(0, len(src), nd)
        d[rows] = src[pick]
    return d
