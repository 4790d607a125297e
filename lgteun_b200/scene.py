"""Tiled full-scene inference: an overlap-tile driver around the forward.

The reference has no scene driver: its loaders cut fixed 64x64 / 256x256 patches offline (dataset/ps_dataset.py) and the
network is applied patch by patch.  A scene larger than the largest supported tile (PAN 1024 x 1024) is therefore processed
here the same way, as a batch of square tiles, but with overlapping tiles whose borders are discarded: every output pixel
comes from a tile in which it lies at least `halo` LrMS pixels away from the tile border (except at the scene border
itself), so the border effects of the bicubic resizes, the zero-padded depthwise convs and the window grid stay out of the
stitched result.

Parity is defined PER TILE: each tile is one ordinary `forward(ms_tile, pan_tile)` — because of the FFT branch a tile is
not equivalent to the same region of a whole-image forward, and no such claim is made.  `plan_tiles` is pure host logic
(unit-tested on CPU); `forward_scene` gathers the tiles on the device, runs them in batches through the module and scatters
the kept regions."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch

UP = 4   # PAN / LrMS resolution ratio (models/unlg_former.py:26)


@dataclass(frozen=True)
class Tile:
    y0: int      # LrMS origin of the tile
    x0: int
    ky0: int     # kept LrMS region [ky0, ky1) x [kx0, kx1) (scene coordinates); the kept regions partition the scene
    ky1: int
    kx0: int
    kx1: int


def _axis(n: int, tile: int, halo: int):
    """Origins and kept intervals along one axis: stride = tile - 2*halo, last origin clamped to n - tile."""
    if n < tile:
        raise ValueError(f"scene side {n} is smaller than the tile side {tile}")
    core = tile - 2 * halo
    if core <= 0:
        raise ValueError("halo too large for the tile")
    out = []
    covered = 0
    while covered < n:
        o = min(max(covered - halo, 0), n - tile)
        end = n if o + tile >= n else o + tile - halo
        out.append((o, covered, end))
        covered = end
    return out


def plan_tiles(h: int, w: int, tile: int = 64, halo: int = 8) -> List[Tile]:
    """Tiles of `tile` x `tile` LrMS pixels covering an h x w LrMS scene; `tile` must be a power of two in [4, 256]
    (PAN side 16..1024, the sizes the kernels support)."""
    if tile < 4 or tile > 256 or tile & (tile - 1):
        raise ValueError("tile must be a power of two in [4, 256] LrMS pixels")
    if halo < 0:
        raise ValueError("halo must be non-negative")
    return [Tile(y0, x0, ky0, ky1, kx0, kx1) for (y0, ky0, ky1) in _axis(h, tile, halo) for (x0, kx0, kx1) in _axis(w, tile, halo)]


@torch.no_grad()
def forward_scene(net, ms: torch.Tensor, pan: torch.Tensor, tile: int = 64, halo: int = 8, batch: int = 64) -> torch.Tensor:
    """ms [B,h,w] (or [1,B,h,w]), pan [1,4h,4w] (or [1,1,4h,4w]) CUDA tensors of one scene -> HrMS [B,4h,4w].
    Tiles are gathered on the device, run `batch` at a time through `net` and their kept regions scattered into the
    output; the same call sequence on the same tiles reproduces every output pixel bit for bit."""
    if ms.dim() == 4:
        if ms.shape[0] != 1:
            raise ValueError("forward_scene handles one scene per call")
        ms = ms[0]
    if pan.dim() == 4:
        pan = pan[0]
    if ms.dim() != 3 or pan.dim() != 3 or pan.shape[0] != 1:
        raise ValueError("expected ms [B,h,w] and pan [1,4h,4w]")
    bands, h, w = ms.shape
    if pan.shape[1] != UP * h or pan.shape[2] != UP * w:
        raise ValueError("pan must be 4x the LrMS size")
    if ms.device.type != "cuda" or pan.device != ms.device:
        raise RuntimeError("forward_scene needs CUDA tensors on one device (there is no CPU path)")
    tiles = plan_tiles(h, w, tile, halo)
    out = torch.empty((bands, UP * h, UP * w), dtype=torch.float32, device=ms.device)
    T = tile
    for lo in range(0, len(tiles), batch):
        chunk = tiles[lo:lo + batch]
        ms_b = torch.stack([ms[:, t.y0:t.y0 + T, t.x0:t.x0 + T] for t in chunk]).contiguous()
        pan_b = torch.stack([pan[:, UP * t.y0:UP * (t.y0 + T), UP * t.x0:UP * (t.x0 + T)] for t in chunk]).contiguous()
        res = net(ms_b, pan_b)
        for i, t in enumerate(chunk):
            ys, ye = UP * (t.ky0 - t.y0), UP * (t.ky1 - t.y0)
            xs, xe = UP * (t.kx0 - t.x0), UP * (t.kx1 - t.x0)
            out[:, UP * t.ky0:UP * t.ky1, UP * t.kx0:UP * t.kx1] = res[i, :, ys:ye, xs:xe]
    return out
