"""Single-process multi-GPU entry of the public API (one host thread + stream + weight replica per device):
`BackgroundRemoval(devices=[0, 1])` splits a batch across the GPUs; every result equals the single-device result bit for bit.
Needs two GPUs (skipped on a one-GPU box)."""
import numpy as np
import pytest
import torch

from s3od_b200.arch import VITB
from s3od_b200.synth import save_checkpoint, synth_image

pytestmark = pytest.mark.gpu

CKPT = "/tmp/s3od_synth_vitb_seed0.pt"


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_device_predictor_matches_single_device():
    import os
    from s3od_b200 import BackgroundRemoval
    if not os.path.exists(CKPT):
        save_checkpoint(CKPT, VITB, 0)
    one = BackgroundRemoval(model_id=CKPT, image_size=128, device="cuda:0", max_batch=4, micro_batch=2)
    two = BackgroundRemoval(model_id=CKPT, image_size=128, devices=[0, 1], max_batch=4, micro_batch=2)
    assert two.devices == ["cuda:0", "cuda:1"] and two.model is two.models[0]
    imgs = [synth_image((128, 96, 160)[i % 3], 128, seed=500 + i) for i in range(7)]
    want = one.remove_background_batch(imgs)
    for _ in range(2):                                           # second call reuses every replica's slots and staging ring
        got = two.remove_background_batch(imgs)
        assert len(got) == len(imgs)
        for w, g in zip(want, got):
            np.testing.assert_array_equal(w.all_masks, g.all_masks)
            np.testing.assert_array_equal(w.all_ious, g.all_ious)
            np.testing.assert_array_equal(np.array(w.rgba_image), np.array(g.rgba_image))
    single = two.remove_background(imgs[3])                      # one image: the first device only
    np.testing.assert_array_equal(single.all_masks, want[3].all_masks)
    two.close()
    one.close()


def test_pageable_and_pinned_inputs_and_pageable_results():
    """Pageable numpy inputs go through the pinned staging ring, pinned ones skip it; result_memory='pageable' returns plain arrays."""
    import os
    from s3od_b200 import BackgroundRemoval
    if not os.path.exists(CKPT):
        save_checkpoint(CKPT, VITB, 0)
    br = BackgroundRemoval(model_id=CKPT, image_size=64, device="cuda:0", max_batch=4, micro_batch=2)
    br2 = BackgroundRemoval(model_id=CKPT, image_size=64, device="cuda:0", max_batch=4, micro_batch=2, result_memory="pageable")
    imgs = [synth_image(64, (64, 48)[i % 2], seed=900 + i) for i in range(9)]
    pinned = []
    for im in imgs:
        t = torch.empty(im.shape, dtype=torch.uint8, pin_memory=True)
        t.numpy()[...] = im
        pinned.append(t.numpy())
    a = br.remove_background_batch(imgs)
    b = br.remove_background_batch(pinned)
    c = br2.remove_background_batch([imgs[i] if i % 2 else pinned[i] for i in range(9)])
    for ra, rb, rc in zip(a, b, c):
        np.testing.assert_array_equal(ra.all_masks, rb.all_masks)
        np.testing.assert_array_equal(ra.all_masks, rc.all_masks)
        np.testing.assert_array_equal(np.array(ra.rgba_image), np.array(rc.rgba_image))
        assert not torch.from_numpy(rc.all_masks).is_pinned() and rc.predicted_mask.base is not None
    # the bounded caches: many geometries do not grow the device-side cache beyond 3 slots x the largest call
    before = br.model.cache_bytes()
    for s in range(40):
        br.remove_background(synth_image(32 + 2 * (s % 9), 64, seed=s))
    assert br.model.cache_bytes() <= before          # nine new (smaller) geometries: carved from the existing per-slot arenas
    assert len(br.model._tab_cache) <= br.model.max_tables
    br.close()
    br2.close()
