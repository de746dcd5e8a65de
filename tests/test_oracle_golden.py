"""The oracle against the reference's own outputs (tests/golden, written by oracle/make_golden.py from
the unmodified /root/reference code) and against the installed third-party libraries it restates."""
import os

import numpy as np
import pytest
import torch

from oracle import model as om
from oracle import prepost as pp
from oracle.pipeline import remove_background
from s3od_b200.arch import VITB


@pytest.mark.parametrize("S", [64, 128])
def test_model_matches_reference(vitb_sd, golden_dir, S):
    g = np.load(os.path.join(golden_dir, f"model_s{S}.npz"))
    st = {}
    out = om.forward(vitb_sd, torch.from_numpy(g["x"]), VITB, st)
    np.testing.assert_allclose(st["embed"].numpy(), g["embed"], atol=2e-5)
    np.testing.assert_allclose(st["layer0"].numpy(), g["layer0"], atol=5e-5)
    for i in range(4):
        np.testing.assert_allclose(st[f"tap{i}"].numpy(), g[f"tap{i}"], atol=2e-4)
    np.testing.assert_allclose(out["pred_iou"].numpy(), g["pred_iou"], atol=1e-4)
    np.testing.assert_allclose(out["pred_masks"].numpy(), g["pred_masks"], atol=5e-4)
    np.testing.assert_allclose(out["features"][:, ::8, ::2, ::2].numpy(), g["features_sub"], atol=2e-4)


PIPE = ["pipe_s128_120x160", "pipe_s128_160x120", "pipe_s64_128x128", "pipe_s128_50x50", "pipe_s128_128x128"]


@pytest.mark.parametrize("name", PIPE)
def test_pipeline_matches_reference(vitb_sd, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    res = remove_background(vitb_sd, g["image"], VITB, int(g["image_size"]))
    np.testing.assert_allclose(res["all_ious"], g["all_ious"], atol=1e-5)
    np.testing.assert_allclose(res["all_masks"], g["all_masks"], atol=2e-5)
    assert int(res["all_ious"].argmax()) == int(g["all_ious"].argmax())
    np.testing.assert_array_equal(res["predicted_mask"], res["all_masks"][res["best_idx"]])
    # alpha = trunc(mask*255): a 1e-5 mask difference can flip a truncation, nothing more
    d = np.abs(res["rgba"][..., 3].astype(int) - g["alpha"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 2e-3
    np.testing.assert_array_equal(res["rgba"][..., :3], g["image"])


def test_odd_padding_raises_like_reference(vitb_sd, golden_dir):
    from s3od_b200.synth import synth_image
    g = np.load(os.path.join(golden_dir, "pipe_odd_padding.npz"))
    assert str(g["error"]) == "ValueError"
    with pytest.raises(ValueError):
        pp.preprocess(synth_image(300, 200, seed=1), 128)


def test_cv2_resize_restatement_bit_exact():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    shapes = [(480, 640, 768, 1024), (100, 100, 256, 256), (400, 200, 200, 100), (37, 53, 64, 92), (300, 200, 128, 85),
              (64, 128, 32, 64), (77, 130, 64, 108), (2, 2, 64, 64), (1, 5, 7, 9), (640, 480, 512, 384)]
    for _ in range(10):
        shapes.append(tuple(int(v) for v in rng.integers(2, 300, 4)))
    for h, w, nh, nw in shapes:
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        np.testing.assert_array_equal(pp.cv_resize_linear_u8(img, nh, nw), cv2.resize(img, (nw, nh)))


def test_aa_resize_restatement():
    import torch.nn.functional as F
    rng = np.random.default_rng(1)
    for h, w, oh, ow in [(96, 128, 60, 80), (128, 128, 13, 13), (64, 64, 128, 128), (128, 128, 250, 250), (96, 64, 51, 37)]:
        x = rng.random((3, h, w), dtype=np.float32)
        ref = F.interpolate(torch.from_numpy(x)[None], size=(oh, ow), mode="bilinear", align_corners=False,
                            antialias=True)[0].numpy()
        np.testing.assert_allclose(pp.aa_resize(x, oh, ow), ref, atol=3e-7)


def test_pad_info_matches_reference_cases():
    # values quoted in SURVEY 8(a) a1 from utils.py:6-29
    assert pp.get_pad_info(2048, 2048, 1024)["resized_size"] == (1024, 1024)
    p = pp.get_pad_info(480, 640, 1024)
    assert p["resized_size"] == (768, 1024) and p["height_pad"] == 128 and p["width_pad"] == 0
    p = pp.get_pad_info(800, 400, 1024)
    assert p["resized_size"] == (1024, 512) and p["width_pad"] == 256


def test_full_size_golden_subsample(vitb_sd, golden_dir):
    """Default image_size=1024 (config 0 of BASELINE.json): ~15 s of CPU."""
    from s3od_b200.synth import synth_image
    g = np.load(os.path.join(golden_dir, "full_s1024.npz"))
    res = remove_background(vitb_sd, synth_image(1024, 1024, seed=int(g["image_seed"])), VITB, 1024)
    np.testing.assert_allclose(res["pred_iou"][None], g["pred_iou"], atol=2e-4)
    np.testing.assert_allclose(res["pred_masks"][None][:, :, 5::16, 3::16], g["pred_masks_sub"], atol=2e-3)
    np.testing.assert_allclose(res["all_masks"][:, 5::16, 3::16], g["all_masks_sub"], atol=1e-4)


@pytest.mark.parametrize("name", ["vis_36x52_k3", "vis_31x45_k3", "vis_20x24_k1", "vis_16x16_k4"])
def test_visualizer_oracle_matches_reference_golden(golden_dir, name):
    """oracle/visualizer.py against the outputs of the unmodified reference (oracle/make_golden_vis.py): bit-exact."""
    from oracle import visualizer as ov
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    image, masks = g["image"], g["masks"]
    assert np.array_equal(ov.visualize_removal(image, masks[0]), g["green"])
    assert np.array_equal(ov.visualize_removal(image, masks[0], (255, 255, 255)), g["white"])
    assert np.array_equal(ov.visualize_removal(image, masks[0], (13, 77, 201)), g["odd"])
    assert np.array_equal(ov.visualize_all_masks(image, masks), g["grid"])
    assert bool(ov.is_ambiguous(masks)) == bool(g["ambiguous"])
    assert bool(ov.is_ambiguous(masks, 0.95)) == bool(g["ambiguous_095"])
    if len(masks) >= 2:
        ious = [ov.compute_mask_iou(masks[i], masks[j]) for i in range(len(masks)) for j in range(i + 1, len(masks))]
        assert np.array_equal(np.array(ious, np.float64), g["ious"])


def test_metrics_oracle_matches_reference_golden(golden_dir):
    """oracle/metrics.py against the values of the unmodified reference (oracle/make_golden_metrics.py)."""
    from oracle import metrics as omt
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    for name in g["names"]:
        r = omt.step(torch.from_numpy(g[name + "_pred"]), torch.from_numpy(g[name + "_mask"]))
        got = np.array([r["mae"], r["max_f"], r["avg_f"], r["s_score"]])
        assert np.abs(got - g[name + "_vals"]).max() <= 1e-7, name
        # weighted F-measure: the scipy-free restatement reproduces the unmodified reference (scipy EDT + convolve) to the last bit
        from oracle.wfm import weighted_f
        assert abs(weighted_f(g[name + "_pred"], g[name + "_mask"]) - float(g[name + "_wfm"])) <= 1e-12, name


def test_feature_transform_ties_match_scipy():
    """oracle/wfm.py::feature_transform against scipy.ndimage.distance_transform_edt(return_indices=True) - the call the
    reference makes (metrics.py:161) - on tie-heavy masks: same distances AND the same nearest pixel for every background pixel."""
    from scipy.ndimage import distance_transform_edt
    from oracle.wfm import feature_transform
    rng = np.random.default_rng(4)
    for trial in range(8):
        H, W = int(rng.integers(9, 45)), int(rng.integers(9, 45))
        gt = rng.random((H, W)) > rng.choice([0.5, 0.85, 0.97])
        if trial % 3 == 0:
            gt = np.zeros((H, W), bool)
            gt[H // 4:H // 2, W // 3:W // 2] = True
            gt[1, 1] = gt[H - 2, W - 3] = True
        if not gt.any():
            continue
        dist, idx = distance_transform_edt(gt == 0, return_indices=True)
        d2, iy, ix = feature_transform(gt)
        bg = ~gt
        assert np.array_equal(iy[bg], idx[0][bg]) and np.array_equal(ix[bg], idx[1][bg])
        assert np.abs(np.sqrt(d2[bg]) - dist[bg]).max() <= 1e-12
