"""the training-side `SODPredictor` (synth_sod/src/synth_sod/model_training/predictor.py:330-477).

CPU tests: the oracle's albumentations front end against the installed cv2 / a direct numpy statement, the host geometry,
the checkpoint formats.  GPU tests: `s3od_b200.sod_predictor.SODPredictor.predict` against the CPU oracle.
The front end is parity-UNPINNED (albumentations is not installed, see oracle/sod_predictor.py)."""
import numpy as np
import pytest
import torch

from oracle import sod_predictor as osp
from s3od_b200.arch import VITB
from s3od_b200.sod_predictor import PredictionResult, SODPredictor, transform_placement
from s3od_b200.synth import synth_image, synth_noise_image, synth_state_dict
from s3od_b200.weights import albumentations_affine, albumentations_lut

SHAPES = [(96, 128), (128, 96), (64, 64), (50, 70), (33, 64), (200, 120), (40, 40)]


@pytest.mark.parametrize("h,w", SHAPES)
@pytest.mark.parametrize("S", [64, 72])
def test_transform_matches_cv2_and_placement(h, w, S):
    """LongestMaxSize + PadIfNeeded + Normalize restated with the oracle's integer resize == the same steps with the
    installed cv2.resize; `transform_placement` (the product's geometry) says where the image lands."""
    cv2 = pytest.importorskip("cv2")
    img = synth_noise_image(h, w, seed=h * 100 + w)
    got = osp.transform(img, S)
    scale = S / float(max(w, h))
    res = img if scale == 1.0 else cv2.resize(img, (int(round(w * scale)), int(round(h * scale))), interpolation=cv2.INTER_LINEAR)
    pl = transform_placement(h, w, S)
    assert pl["resized_size"] == res.shape[:2]
    canvas = np.zeros((S, S, 3), np.uint8)
    t, l = pl["height_pad"], pl["width_pad"]
    canvas[t:t + res.shape[0], l:l + res.shape[1]] = res
    mean = np.array([0.485, 0.456, 0.406], np.float32) * np.float32(255)
    den = np.reciprocal(np.array([0.229, 0.224, 0.225], np.float32) * np.float32(255))
    np.testing.assert_array_equal(got, (canvas.astype(np.float32) - mean) * den)
    assert t == (S - res.shape[0]) // 2 and l == (S - res.shape[1]) // 2           # centred, the extra pixel below / right


def test_normalisation_table_and_affine_hint():
    lut = albumentations_lut().float().numpy().reshape(3, 256)
    ramp = np.ascontiguousarray(np.broadcast_to(np.arange(256, dtype=np.uint8)[None, :, None], (256, 256, 3)))
    x = osp.transform(ramp, 256)[0]                                                                   # (256, 3): no resize, no padding
    np.testing.assert_array_equal(lut.T, torch.from_numpy(np.ascontiguousarray(x)).bfloat16().float().numpy())
    ab = albumentations_affine().numpy()
    assert ab.shape == (6,) and np.all(ab[:3] > 0) and np.all(ab[3:] < 0)


def test_prediction_result_surface():
    r = PredictionResult(binary_mask=np.zeros((2, 2)), soft_mask=np.zeros((2, 2)))
    assert not r.has_multiple_masks and r.num_masks == 1
    r = PredictionResult(np.zeros((2, 2)), np.zeros((2, 2)), all_masks=np.zeros((3, 2, 2)), all_ious=np.zeros(3))
    assert r.has_multiple_masks and r.num_masks == 3


def test_no_cpu_path():
    with pytest.raises(RuntimeError):
        SODPredictor("/nonexistent.pt", device="cpu")


# ---------------------------------------------------------------------------------------------------------------- GPU
def _ckpt(tmp_path, sd, lightning):
    p = tmp_path / "sod.ckpt"
    if lightning:         # Lightning layout: 'model.' prefix + hyper_parameters.config.model (predictor.py:360-369)
        torch.save({"state_dict": {"model." + k: v for k, v in sd.items()},
                    "hyper_parameters": {"config": {"model": {"encoder_name": "dinov3_base", "num_outputs": 3}}}}, p)
    else:
        torch.save({"state_dict": sd}, p)
    return str(p)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,S,lightning", [(96, 128, 64, True), (128, 96, 64, False), (64, 64, 64, False), (50, 70, 64, True),
                                             (33, 64, 72, False), (200, 120, 72, True), (40, 40, 72, False), (128, 128, 128, False)])
def test_predict_matches_oracle(tmp_path, vitb_sd, h, w, S, lightning):
    """image_size 72 is the reference's 840 case in small: 4 x 4 patches, 64^2 masks, 72-based padding."""
    img = synth_image(h, w, seed=7 * h + w)
    ref = osp.predict(vitb_sd, img, VITB, S, 0.5)
    pred = SODPredictor(_ckpt(tmp_path, vitb_sd, lightning), image_size=S, device="cuda:0")
    assert pred.model.preprocess_mode() == 1
    got = pred.predict(img, 0.5)
    assert got.soft_mask.shape == (h, w) and got.soft_mask.dtype == np.float32
    assert got.all_masks.shape == (3, h, w) and got.all_masks.dtype == np.float32
    assert set(np.unique(got.all_masks)) <= {0.0, 1.0}
    # bf16 tolerances of tests/test_gpu_parity.py
    assert np.abs(got.soft_mask - ref["soft_mask"]).max() <= 4e-2
    np.testing.assert_allclose(got.all_ious, ref["all_ious"], atol=1e-2)
    assert int(got.all_ious.argmax()) == ref["best_idx"]
    np.testing.assert_array_equal(got.binary_mask, (got.soft_mask > 0.5).astype(np.float32))
    np.testing.assert_array_equal(got.binary_mask, got.all_masks[ref["best_idx"]])
    confident = np.abs(ref["soft_mask"] - 0.5) > 0.06
    assert np.array_equal(got.binary_mask[confident], ref["binary_mask"][confident])
    pred.model.close()


@pytest.mark.gpu
def test_single_output_model_returns_no_all_masks(tmp_path):
    """num_outputs == 1 branch (predictor.py:451-456); the shapes take the K = 1 instantiations of the tile, identity, exact-2x
    and per-pixel post-process kernels."""
    from dataclasses import replace
    arch = replace(VITB, num_outputs=1)
    sd = synth_state_dict(arch, 3)
    p = tmp_path / "one.pt"
    torch.save({"state_dict": sd}, p)
    pred = SODPredictor(str(p), image_size=64, device="cuda:0")
    for h, w in [(80, 64), (64, 64), (128, 128), (40, 30), (50, 70)]:
        img = synth_image(h, w, seed=5 + h)
        ref = osp.predict(sd, img, arch, 64, 0.4)
        got = pred.predict(img, 0.4)
        assert got.all_masks is None and got.all_ious is None and not got.has_multiple_masks
        assert got.soft_mask.shape == (h, w)
        assert np.abs(got.soft_mask - ref["soft_mask"]).max() <= 4e-2
        np.testing.assert_array_equal(got.binary_mask, (got.soft_mask > 0.4).astype(np.float32))
    pred.model.close()


@pytest.mark.gpu
def test_odd_patch_grid_is_rejected(tmp_path, vitb_sd):
    with pytest.raises(ValueError):
        SODPredictor(_ckpt(tmp_path, vitb_sd, False), image_size=88, device="cuda:0")      # 5 x 5 patches
