"""SURVEY 8f rank 2 on the GPU: s3od_b200.visualizer (CUDA library through the C ABI) against the reference's golden outputs
and against the numpy oracle at full size.  Everything is integer / fixed-order float32 arithmetic: bit-exact."""
import os
import time

import numpy as np
import pytest
import torch

from oracle import visualizer as ov

pytestmark = pytest.mark.gpu


class _Res:
    def __init__(self, masks):
        self.predicted_mask, self.all_masks = masks[0], masks


@pytest.mark.parametrize("name", ["vis_36x52_k3", "vis_31x45_k3", "vis_20x24_k1", "vis_16x16_k4"])
def test_visualizer_matches_reference_golden(golden_dir, name):
    from s3od_b200 import visualizer as vis
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    image, masks = g["image"], g["masks"]
    res = _Res(masks)
    assert np.array_equal(np.array(vis.visualize_removal(image, res)), g["green"])
    assert np.array_equal(np.array(vis.visualize_removal(image, res, (255, 255, 255))), g["white"])
    assert np.array_equal(np.array(vis.visualize_removal(image, res, (13, 77, 201))), g["odd"])
    grid = vis.visualize_all_masks(image, res)
    assert grid.size == (g["grid"].shape[1], g["grid"].shape[0])
    assert np.array_equal(np.array(grid), g["grid"])
    assert vis.is_ambiguous(masks) == bool(g["ambiguous"])
    assert vis.is_ambiguous(masks, 0.95) == bool(g["ambiguous_095"])
    if len(masks) >= 2:
        ious = [vis.compute_mask_iou(masks[i], masks[j]) for i in range(len(masks)) for j in range(i + 1, len(masks))]
        assert np.array_equal(np.array(ious, np.float64), g["ious"])
        assert vis.is_ambiguous(masks[:2], 0.5) == bool(g["ambiguous_pair"])


def test_visualizer_fullsize_matches_oracle_and_bandwidth(capsys):
    """2048 x 2048 (configs[2] source size), device-resident in and out: bit-exact against the numpy oracle; prints the
    achieved bandwidth of the three kernels (algorithmic bytes / CUDA-event time)."""
    from s3od_b200 import visualizer as vis
    rng = np.random.default_rng(3)
    H = W = 2048
    image = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    masks = np.clip(rng.random((3, H, W), dtype=np.float32) * 1.4 - 0.2, 0, 1).astype(np.float32)
    masks[1] = np.clip(masks[0] + 0.02 * rng.standard_normal((H, W)).astype(np.float32), 0, 1)
    d_img, d_masks = torch.from_numpy(image).cuda(), torch.from_numpy(masks).cuda()
    comp = vis.composite_device(d_img, d_masks[0], (0, 255, 0))
    grid = vis.mask_grid_device(d_img, d_masks)
    counts = vis.pair_counts_device(d_masks)
    torch.cuda.synchronize()
    assert np.array_equal(comp.cpu().numpy(), ov.visualize_removal(image, masks[0]))
    assert np.array_equal(grid.cpu().numpy(), ov.visualize_all_masks(image, masks))
    ref = [[int(np.logical_and(masks[i] > 0.5, masks[j] > 0.5).sum()), int(np.logical_or(masks[i] > 0.5, masks[j] > 0.5).sum())]
           for i in range(3) for j in range(i + 1, 3)]
    assert counts.cpu().tolist() == ref
    assert vis.is_ambiguous(d_masks) == ov.is_ambiguous(masks)

    def timed(fn, n=20):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    npix = H * W
    rows = [("composite", timed(lambda: vis.composite_device(d_img, d_masks[0], (0, 255, 0))), npix * (3 + 4 + 3)),
            ("mask grid (3)", timed(lambda: vis.mask_grid_device(d_img, d_masks)), npix * 3 * (3 + 4 + 3)),
            ("pair counts (3)", timed(lambda: vis.pair_counts_device(d_masks)), npix * 3 * 4)]
    with capsys.disabled():
        for name, ms, nbytes in rows:
            print(f"\n[visualizer 2048x2048] {name}: {ms * 1e3:.1f} us, {nbytes / ms / 1e6:.0f} GB/s (incl. output allocation)")
