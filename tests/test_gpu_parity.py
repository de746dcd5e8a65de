"""Parity of the CUDA path (through the C ABI) with the reference's outputs (tests/golden) and with the CPU oracle.

Tolerances (bf16 operands, fp32 accumulation; BASELINE.json north_star "stated bf16 tolerance").  Rounding only the
WEIGHTS to bf16 already moves sigmoid masks by up to 2e-2 on these seeded weights (tests/test_host_cpu.py), so:
    stage tensors     rel-L2 <= 2.5e-2
    sigmoid masks     max-abs <= 4e-2 (<= 5e-2 over all 3.1 M pixels of the 1024^2 case), mean-abs <= 6e-3; masks thresholded at 0.5: IoU >= 0.985 and NOT ONE flipped pixel
                      among those whose reference logit is further than 0.25 from the threshold.  (The seeded weights give
                      soft masks - logit std ~3 - so ~1 % of all pixels sit within the ~1.3 % relative logit error of the
                      threshold; that band alone costs up to 1.5 % IoU whatever the logit scale.)
    IoU logits        max-abs <= 3e-2;  sigmoid(IoU) max-abs <= 1e-2;  best-mask index exact
    preprocess        bit-exact (integer resize + LUT);  postprocess on given logits <= 5e-6;  alpha exact w.r.t. own mask
"""
import os

import numpy as np
import pytest
import torch

from oracle import model as om
from oracle import prepost as opp
from oracle.pipeline import remove_background as oracle_remove_background
from s3od_b200.arch import VITB
from s3od_b200.synth import synth_image, synth_noise_image

pytestmark = pytest.mark.gpu

CKPT = "/tmp/s3od_synth_vitb_seed0.pt"


def synth_sd():
    from s3od_b200.synth import synth_state_dict
    return synth_state_dict(VITB, 0)


def _rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _mask_metrics(logits, ref_logits):
    a, b = torch.sigmoid(logits.float()), torch.sigmoid(ref_logits.float())
    inter = float(((a > 0.5) & (b > 0.5)).sum())
    union = float(((a > 0.5) | (b > 0.5)).sum())
    confident = ref_logits.float().abs() > 0.25
    flips = int((((a > 0.5) != (b > 0.5)) & confident).sum())
    return float((a - b).abs().max()), float((a - b).abs().mean()), inter / max(union, 1.0), flips


def _assert_masks(logits, ref_logits, max_abs=4e-2):
    mx, mean, iou, flips = _mask_metrics(logits, ref_logits)
    msg = f"sigmoid max-abs {mx:.4f} mean-abs {mean:.5f} thresholded IoU {iou:.4f} confident flips {flips}"
    print(msg)
    assert mx <= max_abs, msg
    assert mean <= 6e-3, msg
    assert iou >= 0.985, msg
    assert flips == 0, msg


@pytest.fixture(scope="module")
def models(vitb_sd):
    from s3od_b200.engine import B200DPTSegmentation
    cache = {}

    def get(S, max_batch=2, micro_batch=None):
        key = (S, max_batch, micro_batch)
        if key not in cache:
            cache[key] = B200DPTSegmentation(vitb_sd, VITB, S, "cuda:0", max_batch=max_batch, micro_batch=micro_batch)
        return cache[key]

    yield get
    for m in cache.values():
        m.close()


@pytest.fixture(scope="module")
def predictors():
    from s3od_b200 import BackgroundRemoval
    from s3od_b200.synth import save_checkpoint
    if not os.path.exists(CKPT):
        save_checkpoint(CKPT, VITB, 0)
    cache = {}

    def get(S, **kw):
        key = (S, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = BackgroundRemoval(model_id=CKPT, image_size=S, device="cuda:0", **kw)
        return cache[key]

    return get


@pytest.mark.parametrize("S", [64, 128])
def test_model_matches_reference_golden(models, vitb_sd, golden_dir, S):
    g = np.load(os.path.join(golden_dir, f"model_s{S}.npz"))
    x = torch.from_numpy(g["x"])
    B = x.shape[0]
    m = models(S, max_batch=2)
    out = m(x.cuda())
    torch.cuda.synchronize()
    gp = S // 16
    P = gp * gp
    for j in range(4):
        tap = m.stage(f"tap{j}", torch.bfloat16, (B, P, 768)).float().cpu()
        assert _rel(tap, torch.from_numpy(g[f"tap{j}"])) <= 2.5e-2
    st = {}
    ref = om.forward(vitb_sd, x, VITB, st)
    for name, C, R in (("l1", 256, 4 * gp), ("l2", 256, 2 * gp), ("l3", 256, gp), ("l4", 256, gp // 2), ("p4", 256, gp),
                       ("p3", 256, 2 * gp), ("p2", 256, 4 * gp), ("p1", 256, 8 * gp), ("feat", 64, S)):
        t = m.stage(name, torch.bfloat16, (B, R, R, C)).float().cpu().permute(0, 3, 1, 2)
        assert _rel(t, st[name]) <= 2.5e-2, name
    pm, pi = out["pred_masks"].cpu(), out["pred_iou"].cpu()
    gm, gi = torch.from_numpy(g["pred_masks"]), torch.from_numpy(g["pred_iou"])
    _assert_masks(pm, gm)
    assert float((pi - gi).abs().max()) <= 3e-2
    assert torch.equal(pi.argmax(1), gi.argmax(1))
    assert float((ref["pred_masks"] - gm).abs().max()) < 1e-3          # the oracle agrees with the reference here too


def test_model_matches_oracle_on_fresh_input(models, vitb_sd):
    S = 96                                                            # odd patch grid (6x6), 3x3 coarsest level
    x = torch.from_numpy(np.concatenate([opp.preprocess(synth_image(S, S, seed=500 + i), S)[0] for i in range(2)], 0))
    m = models(S, max_batch=2)
    out = m(x.cuda())
    ref = om.forward(vitb_sd, x, VITB)
    # the maximum over 3.1 M pixels is an extreme value: 5e-2 here, 4e-2 on the <= 50 k-pixel cases
    _assert_masks(out["pred_masks"].cpu(), ref["pred_masks"], max_abs=5e-2)
    assert float((out["pred_iou"].cpu() - ref["pred_iou"]).abs().max()) <= 3e-2


def test_vitl_single_output_matches_oracle():
    """BASELINE.json config 4 architecture: ViT-L/16 backbone (D=1024, 16 heads, 24 layers, taps 4/11/17/23), one mask."""
    from s3od_b200.arch import VITL
    from s3od_b200.engine import B200DPTSegmentation
    from s3od_b200.synth import synth_state_dict
    sd = synth_state_dict(VITL, 3)
    S = 64
    x = torch.from_numpy(np.concatenate([opp.preprocess(synth_image(S, S, seed=900 + i), S)[0] for i in range(2)], 0))
    m = B200DPTSegmentation(sd, VITL, S, "cuda:0", max_batch=2)
    try:
        out = m(x.cuda())
        ref = om.forward(sd, x, VITL)
        assert out["pred_masks"].shape == (2, 1, S, S) and out["pred_iou"].shape == (2, 1)
        _assert_masks(out["pred_masks"].cpu(), ref["pred_masks"], max_abs=5e-2)
        assert float((out["pred_iou"].cpu() - ref["pred_iou"]).abs().max()) <= 5e-2
    finally:
        m.close()


def test_micro_batching_is_deterministic(models):
    """B=3 pushed through micro-batches of 2 (+1 ragged) gives bit-identical logits to the same images run alone."""
    S = 64
    x = torch.from_numpy(np.concatenate([opp.preprocess(synth_image(S, S, seed=40 + i), S)[0] for i in range(3)], 0)).cuda()
    big = models(S, max_batch=3, micro_batch=2)
    one = models(S, max_batch=2)
    out = big(x)
    for i in range(3):
        o = one(x[i:i + 1])
        assert torch.equal(o["pred_masks"][0], out["pred_masks"][i])
        assert torch.equal(o["pred_iou"][0], out["pred_iou"][i])
    again = big(x)
    assert torch.equal(again["pred_masks"], out["pred_masks"])


def test_forward_rejects_wrong_shape(models):
    with pytest.raises(ValueError):
        models(64)(torch.zeros(1, 3, 32, 32).cuda())
    with pytest.raises(ValueError):
        models(64, max_batch=2)(torch.zeros(3, 3, 64, 64).cuda())     # batch > max_batch


@pytest.mark.parametrize("h,w,S", [(128, 128, 64), (64, 64, 64), (120, 160, 128), (160, 120, 128), (50, 50, 128), (37, 91, 64),
                                   (480, 640, 256), (512, 512, 256), (100, 100, 128), (256, 64, 128)])
def test_preprocess_bit_exact(models, h, w, S):
    """uint8 letterbox resize + normalise on the device == bf16(reference float tensor), bit for bit."""
    img = synth_noise_image(h, w, seed=h * 1000 + w)
    try:
        ref, _ = opp.preprocess(img, S)
    except ValueError:
        with pytest.raises(ValueError):
            models(S).geometry(h, w)
        return
    m = models(S)
    assert m.preprocess_mode() == 1          # the FMA form reproduces the normalisation table, so the library selected it
    m.preprocess([torch.from_numpy(img).cuda()])
    gp = S // 16
    patches = m.stage("patches", torch.bfloat16, (gp * gp, 768)).float().cpu()
    got = patches.reshape(gp, gp, 3, 16, 16).permute(2, 0, 3, 1, 4).reshape(3, S, S)
    assert torch.equal(got, torch.from_numpy(ref[0]).bfloat16().float())


@pytest.mark.parametrize("H,W,S,hp,wp", [(128, 128, 64, 0, 0), (120, 160, 128, 16, 0), (160, 120, 128, 0, 16), (50, 50, 128, 0, 0),
                                         (300, 300, 128, 0, 0), (64, 64, 64, 0, 0),
                                         # shared-memory tile kernel: several 16 x 512 tiles, ragged edges, 2x / identity / odd ratios
                                         (1024, 1024, 512, 0, 0), (512, 512, 512, 0, 0), (520, 1040, 256, 64, 0),
                                         (96, 128, 128, 16, 0), (128, 96, 128, 0, 16),         # identity kernel with padding
                                         (192, 256, 128, 16, 0), (256, 192, 128, 0, 16), (1160, 1240, 1024, 222, 202),   # exact-2x kernel
                                         (1100, 600, 512, 0, 116), (700, 1028, 512, 82, 0), (2048, 2048, 1024, 0, 0)])
def test_postprocess_matches_oracle_on_given_logits(models, H, W, S, hp, wp):
    rng = np.random.default_rng(H + W)
    logits = (rng.standard_normal((1, 3, S, S)) * 3).astype(np.float32)
    ioul = rng.standard_normal((1, 3)).astype(np.float32)
    img = synth_noise_image(H, W, seed=3)
    pad = dict(height_pad=hp, width_pad=wp, original_size=(H, W), resized_size=(S - 2 * hp, S - 2 * wp))
    ref = opp.postprocess(logits[0], ioul[0], pad, img)
    m = models(S)
    d_img = torch.from_numpy(img).cuda()
    outs, ious, best = m.postprocess(torch.from_numpy(logits).cuda(), torch.from_numpy(ioul).cuda(), [d_img], [pad])
    am, rgba = outs[0][0].cpu().numpy(), outs[0][1].cpu().numpy()
    np.testing.assert_allclose(am, ref["all_masks"], atol=5e-6)
    np.testing.assert_allclose(ious.cpu().numpy()[0], ref["all_ious"], atol=1e-6)
    assert int(best[0]) == ref["best_idx"]
    np.testing.assert_array_equal(rgba[..., :3], img)
    np.testing.assert_array_equal(rgba[..., 3], (am[int(best[0])] * 255).astype(np.uint8))      # predictor.py:130, exact
    assert np.abs(rgba[..., 3].astype(int) - ref["rgba"][..., 3].astype(int)).max() <= 1


PIPE = ["pipe_s128_120x160", "pipe_s128_160x120", "pipe_s64_128x128", "pipe_s128_50x50", "pipe_s128_128x128"]


@pytest.mark.parametrize("name", PIPE)
def test_remove_background_matches_reference_golden(predictors, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    br = predictors(int(g["image_size"]))
    res = br.remove_background(g["image"])
    assert res.all_masks.shape == g["all_masks"].shape and res.all_masks.dtype == np.float32
    assert res.predicted_mask.shape == g["image"].shape[:2] and res.all_ious.shape == (3,)
    d = np.abs(res.all_masks - g["all_masks"])
    assert d.max() <= 4e-2 and d.mean() <= 6e-3
    a, b = res.all_masks > 0.5, g["all_masks"] > 0.5
    assert (a & b).sum() / max((a | b).sum(), 1) >= 0.985
    confident = np.abs(g["all_masks"] - 0.5) > 0.06                      # |logit| > 0.25
    assert not ((a != b) & confident).any()
    assert np.abs(res.all_ious - g["all_ious"]).max() <= 1e-2
    assert int(res.all_ious.argmax()) == int(g["all_ious"].argmax())
    # reference contract tests (tests/test_fixture_inference.py:92-116, 73-89)
    np.testing.assert_array_equal(res.predicted_mask, res.all_masks[res.all_ious.argmax()])
    assert res.rgba_image.mode == "RGBA" and res.rgba_image.size == (g["image"].shape[1], g["image"].shape[0])
    rgba = np.array(res.rgba_image)
    np.testing.assert_array_equal(rgba[..., :3], g["image"])
    np.testing.assert_array_equal(rgba[..., 3], (res.predicted_mask * 255).astype(np.uint8))
    assert np.abs(rgba[..., 3].astype(int) - g["alpha"].astype(int)).max() <= 11        # 4e-2 * 255


def test_remove_background_inputs_and_errors(predictors):
    """Reference contract tests (tests/test_inference_package.py:49-122): numpy and PIL inputs, rectangles, ValueError."""
    from PIL import Image
    br = predictors(128)
    img = synth_noise_image(100, 100, seed=1)
    r1 = br.remove_background(img)
    r2 = br.remove_background(Image.fromarray(img))
    assert r1.predicted_mask.shape == (100, 100) and r1.all_masks.ndim == 3 and len(r1.all_masks) == len(r1.all_ious) == 3
    np.testing.assert_array_equal(r1.all_masks, r2.all_masks)
    assert (r1.all_ious >= 0).all() and (r1.all_ious <= 1).all() and r1.all_masks.min() >= 0 and r1.all_masks.max() <= 1
    for h, w in ((64, 128), (128, 64)):
        assert br.remove_background(synth_noise_image(h, w, seed=2)).predicted_mask.shape == (h, w)
    with pytest.raises(ValueError):
        br.remove_background(synth_noise_image(300, 200, seed=1))    # odd padding: the reference raises ValueError (F11)


def test_batch_api_matches_single_calls(predictors):
    br = predictors(64, max_batch=4, micro_batch=2)
    imgs = [synth_image(128, 128, seed=70), synth_image(48, 64, seed=71), synth_image(64, 64, seed=72), synth_image(64, 32, seed=73),
            synth_image(32, 32, seed=74)]
    batch = br.remove_background_batch(imgs)
    assert len(batch) == 5
    for im, rb in zip(imgs, batch):
        rs = br.remove_background(im)
        np.testing.assert_array_equal(rs.all_masks, rb.all_masks)
        np.testing.assert_array_equal(rs.all_ious, rb.all_ious)
        np.testing.assert_array_equal(np.array(rs.rgba_image), np.array(rb.rgba_image))


def test_batch_api_chunk_schedule_with_rotating_slots(predictors):
    """16 images at micro-batch 16 run as chunks of 8, 4, 4 (lead + tails) over three rotating output slots; a second call
    reuses the slots while the first results are still referenced."""
    br = predictors(64, max_batch=16, micro_batch=16)
    imgs = [synth_image((64, 128, 96)[i % 3], 64, seed=100 + i) for i in range(16)]
    first = br.remove_background_batch(imgs)
    second = br.remove_background_batch(imgs[::-1])
    for i in (0, 5, 9, 13, 15):
        single = br.remove_background(imgs[i])
        for got in (first[i], second[15 - i]):
            np.testing.assert_array_equal(single.all_masks, got.all_masks)
            np.testing.assert_array_equal(single.all_ious, got.all_ious)
            np.testing.assert_array_equal(np.array(single.rgba_image), np.array(got.rgba_image))


def test_c_abi_rejects_misaligned_buffers(models):
    """The vector paths of the pre / post-process kernels need aligned buffers: the C ABI refuses anything else (ValueError)."""
    m = models(64)
    raw = torch.zeros(64 * 64 * 3 + 1, dtype=torch.uint8, device="cuda")
    skewed = raw[1:].view(64, 64, 3)                                   # contiguous, data_ptr % 16 == 1
    with pytest.raises(ValueError):
        m.preprocess([skewed])
    logits = torch.zeros(1, 3, 64, 64, device="cuda")
    pad = dict(height_pad=0, width_pad=0, original_size=(64, 64), resized_size=(64, 64))
    with pytest.raises(ValueError):
        m.postprocess(logits, torch.zeros(1, 3, device="cuda"), [skewed], [pad])
    good = raw[:64 * 64 * 3].view(64, 64, 3)
    m.preprocess([good])
    m.postprocess(logits, torch.zeros(1, 3, device="cuda"), [good], [pad])
    torch.cuda.synchronize()


def test_shared_instance_from_several_threads(predictors):
    """The reference's demo calls one shared predictor from worker threads (demo/app.py:18-25): calls are serialised."""
    import threading
    br = predictors(64, max_batch=4, micro_batch=2)
    imgs = [synth_image((64, 128, 96)[i % 3], 64, seed=80 + i) for i in range(6)]
    want = [br.remove_background(im) for im in imgs]
    got = [None] * len(imgs)

    def work(i):
        for _ in range(3):
            got[i] = br.remove_background(imgs[i])
    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(imgs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for w, g in zip(want, got):
        np.testing.assert_array_equal(w.all_masks, g.all_masks)
        np.testing.assert_array_equal(w.all_ious, g.all_ious)
        np.testing.assert_array_equal(np.array(w.rgba_image), np.array(g.rgba_image))


def test_full_size_against_reference_golden(predictors, golden_dir):
    """BASELINE.json config 0/1 shape: image_size 1024, one 1024x1024 image; sub-sampled reference outputs."""
    g = np.load(os.path.join(golden_dir, "full_s1024.npz"))
    br = predictors(1024, max_batch=2)
    img = synth_image(1024, 1024, seed=int(g["image_seed"]))
    x, _ = opp.preprocess(img, 1024)
    out = br.model(torch.from_numpy(x).cuda())
    _assert_masks(out["pred_masks"].cpu()[:, :, 5::16, 3::16], torch.from_numpy(g["pred_masks_sub"]))
    assert float((out["pred_iou"].cpu() - torch.from_numpy(g["pred_iou"])).abs().max()) <= 3e-2
    res = br.remove_background(img)
    d = np.abs(res.all_masks[:, 5::16, 3::16] - g["all_masks_sub"])
    assert d.max() <= 4e-2 and d.mean() <= 6e-3
    # every pixel, against the oracle run on the GPU box's CPU (the fixture only holds a 1/256 sub-sample)
    ref = om.forward(synth_sd(), torch.from_numpy(x), VITB)
    # the maximum over 3.1 M pixels is an extreme value: 5e-2 here, 4e-2 on the <= 50 k-pixel cases
    _assert_masks(out["pred_masks"].cpu(), ref["pred_masks"], max_abs=5e-2)
    assert int(res.all_ious.argmax()) == int(g["all_ious"].argmax())
    assert np.abs(res.all_ious - g["all_ious"]).max() <= 1e-2
    # size-independent properties at full size
    np.testing.assert_array_equal(res.predicted_mask, res.all_masks[res.all_ious.argmax()])
    rgba = np.array(res.rgba_image)
    np.testing.assert_array_equal(rgba[..., :3], img)
    np.testing.assert_array_equal(rgba[..., 3], (res.predicted_mask * 255).astype(np.uint8))


def test_config3_shape_2048_source(predictors):
    """BASELINE.json config 2 shape: 2048x2048 source -> exact 2x integer preprocess, 2x mask up-sample, RGBA composite."""
    br = predictors(1024, max_batch=2)
    img = synth_image(2048, 2048, seed=9)
    res = br.remove_background_batch([img, img[::-1].copy()])
    for r, im in zip(res, (img, img[::-1])):
        assert r.all_masks.shape == (3, 2048, 2048) and r.rgba_image.size == (2048, 2048)
        rgba = np.array(r.rgba_image)
        np.testing.assert_array_equal(rgba[..., :3], im)
        np.testing.assert_array_equal(rgba[..., 3], (r.predicted_mask * 255).astype(np.uint8))
        assert 0.0 <= r.all_masks.min() and r.all_masks.max() <= 1.0
    x, _ = opp.preprocess(img, 1024)
    gp = 64
    br.model.preprocess([torch.from_numpy(img).cuda()])
    patches = br.model.stage("patches", torch.bfloat16, (gp * gp, 768)).float().cpu()
    got = patches.reshape(gp, gp, 3, 16, 16).permute(2, 0, 3, 1, 4).reshape(3, 1024, 1024)
    assert torch.equal(got, torch.from_numpy(x[0]).bfloat16().float())


def test_fullsize_1024_matches_oracle_on_device(models, vitb_sd, capsys):
    """Full configs[1] resolution: the CUDA path against the oracle's fp32 arithmetic run with torch on the SAME GPU
    (TF32 off), two noise images.  Also prints the eager timings (fp32 and bf16 autocast) of that torch path - the
    'PyTorch on B200' figure quoted in DESIGN.md; informational, not a benchmark."""
    import time
    S = 1024
    x = torch.from_numpy(np.concatenate([opp.preprocess(synth_noise_image(S, S, seed=900 + i), S)[0] for i in range(2)], 0))
    m = models(S, max_batch=2)
    out = m(x.cuda())
    torch.cuda.synchronize()
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        sd_dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in vitb_sd.items()}
        xd = x.cuda()
        with torch.no_grad():
            ref = om.forward(sd_dev, xd, VITB)
            torch.cuda.synchronize()
            timings = {}
            for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16 autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
                with ctx:
                    om.forward(sd_dev, xd, VITB)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(2):
                        om.forward(sd_dev, xd, VITB)
                    torch.cuda.synchronize()
                    timings[name] = 2 * 2 / (time.perf_counter() - t0)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    t0 = time.perf_counter()
    for _ in range(4):
        m(xd)
    torch.cuda.synchronize()
    ours = 4 * 2 / (time.perf_counter() - t0)
    with capsys.disabled():
        print(f"\n[eager torch on this GPU, model forward only, batch 2 @1024] fp32 {timings['fp32']:.2f} img/s, "
              f"bf16 autocast {timings['bf16 autocast']:.2f} img/s; this library (same call, batch 2) {ours:.1f} img/s")
    _assert_masks(out["pred_masks"].cpu(), ref["pred_masks"].cpu(), max_abs=5e-2)
    assert float((out["pred_iou"].cpu() - ref["pred_iou"].cpu()).abs().max()) <= 3e-2
    assert torch.equal(out["pred_iou"].cpu().argmax(1), ref["pred_iou"].cpu().argmax(1))
