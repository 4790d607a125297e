"""Tiled full-scene driver: host logic on CPU, per-tile parity on the GPU."""
import numpy as np
import pytest
import torch

from conftest import load_weights


@pytest.mark.parametrize("h,w,tile,halo", [(64, 64, 64, 8), (96, 160, 64, 8), (100, 73, 32, 4), (256, 300, 128, 16),
                                           (40, 40, 32, 0), (33, 64, 32, 8)])
def test_plan_tiles_partitions_the_scene(h, w, tile, halo):
    from lgteun_b200.scene import plan_tiles
    tiles = plan_tiles(h, w, tile, halo)
    cover = np.zeros((h, w), dtype=np.int32)
    for t in tiles:
        assert 0 <= t.y0 <= h - tile and 0 <= t.x0 <= w - tile
        cover[t.ky0:t.ky1, t.kx0:t.kx1] += 1
        # kept pixels are at least `halo` away from the tile border unless that border is the scene border
        assert t.ky0 == 0 or t.ky0 - t.y0 >= halo
        assert t.kx0 == 0 or t.kx0 - t.x0 >= halo
        assert t.ky1 == h or (t.y0 + tile) - t.ky1 >= halo
        assert t.kx1 == w or (t.x0 + tile) - t.kx1 >= halo
        assert t.y0 <= t.ky0 < t.ky1 <= t.y0 + tile and t.x0 <= t.kx0 < t.kx1 <= t.x0 + tile
    assert (cover == 1).all()


def test_plan_tiles_errors():
    from lgteun_b200.scene import plan_tiles
    with pytest.raises(ValueError):
        plan_tiles(64, 64, 48, 8)          # not a power of two
    with pytest.raises(ValueError):
        plan_tiles(16, 64, 32, 4)          # scene smaller than the tile
    with pytest.raises(ValueError):
        plan_tiles(64, 64, 16, 8)          # nothing left between the halos


@pytest.mark.gpu
def test_forward_scene_matches_per_tile_oracle():
    """Every stitched pixel equals the oracle's forward of the tile it was taken from (tolerance of the forward: 1e-3)."""
    import lgteun_b200
    from lgteun_b200.scene import UP, forward_scene, plan_tiles
    from oracle import lgteun_oracle as O
    from types import SimpleNamespace
    sd = load_weights(4)
    net = lgteun_b200.Pansharpening(SimpleNamespace(ms_chans=4), None, stage=2)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    g = torch.Generator().manual_seed(23)
    h, w, tile, halo = 40, 72, 32, 4
    ms = torch.rand(4, h, w, generator=g)
    pan = torch.rand(1, UP * h, UP * w, generator=g)
    out = forward_scene(net, ms.cuda(), pan.cuda(), tile=tile, halo=halo, batch=4).cpu()
    assert out.shape == (4, UP * h, UP * w)
    tiles = plan_tiles(h, w, tile, halo)
    assert len(tiles) > 4                                     # several batches
    for t in tiles:
        ref = O.forward(sd, ms[None, :, t.y0:t.y0 + tile, t.x0:t.x0 + tile],
                        pan[None, :, UP * t.y0:UP * (t.y0 + tile), UP * t.x0:UP * (t.x0 + tile)])[0]
        ys, ye, xs, xe = UP * (t.ky0 - t.y0), UP * (t.ky1 - t.y0), UP * (t.kx0 - t.x0), UP * (t.kx1 - t.x0)
        got = out[:, UP * t.ky0:UP * t.ky1, UP * t.kx0:UP * t.kx1]
        assert (got - ref[:, ys:ye, xs:xe]).abs().max().item() <= 1e-3
    # a scene of exactly one tile is the plain forward
    one = forward_scene(net, ms[:, :32, :32].cuda(), pan[:, :128, :128].cuda(), tile=32, halo=4)
    with torch.no_grad():
        plain = net(ms[None, :, :32, :32].cuda(), pan[None, :, :128, :128].cuda())[0]
    assert torch.equal(one, plain)
    with pytest.raises(RuntimeError):
        forward_scene(net, ms, pan, tile=32, halo=4)          # CPU tensors: no fallback
