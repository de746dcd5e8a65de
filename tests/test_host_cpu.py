"""CPU-side tests: host logic, weight packing, the C-ABI library's exported surface, sharding (gloo, world_size 2)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import prepost as opp                      # noqa: E402
from s3od_b200 import geometry, utils                  # noqa: E402
from s3od_b200.arch import VITB                        # noqa: E402


def test_library_builds_and_exports_every_declared_symbol():
    from s3od_b200.build import build_lib
    path = build_lib()
    lib = ctypes.CDLL(path)                            # loads without a GPU (static cudart, no compute calls)
    header = open(os.path.join(ROOT, "include", "s3od_b200.h")).read()
    names = sorted(set(re.findall(r"\b(s3od_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/s3od_b200.h but not exported"
    lib.s3od_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.s3od_version()


def test_create_rejects_bad_arguments_without_gpu():
    from s3od_b200.engine import load_library
    lib = load_library()
    ctx = ctypes.c_void_p()
    assert lib.s3od_create(ctypes.byref(ctx), 0, 7, 3, 1024, 1, 1) == -1          # unknown arch
    assert lib.s3od_create(ctypes.byref(ctx), 0, 0, 2, 1024, 1, 1) == -1          # num_outputs
    assert lib.s3od_create(ctypes.byref(ctx), 0, 0, 3, 1000, 1, 1) == -1          # image_size % 32
    assert b"image_size" in lib.s3od_last_error()


def test_pad_info_matches_oracle_and_reference_cases():
    for h, w, S in [(2048, 2048, 1024), (480, 640, 1024), (800, 400, 1024), (100, 100, 1024), (2000, 2000, 1024), (120, 160, 128)]:
        a = utils.get_pad_info(np.empty((h, w, 3), np.uint8), S)
        assert a == opp.get_pad_info(h, w, S)
    assert utils.get_pad_info(np.empty((480, 640, 3), np.uint8), 1024)["height_pad"] == 128


def test_odd_padding_raises_value_error():
    pad = utils.get_pad_info(np.empty((300, 200, 3), np.uint8), 128)
    with pytest.raises(ValueError):
        utils.check_padding(pad, 128)
    utils.check_padding(utils.get_pad_info(np.empty((120, 160, 3), np.uint8), 128), 128)


def test_linear_tables_match_oracle():
    for dst, src in [(1024, 640), (768, 480), (64, 37), (92, 53), (1024, 100), (85, 200), (682, 1999)]:
        for vertical in (False, True):
            np.testing.assert_array_equal(geometry.linear_tables(dst, src, vertical), np.stack(opp._linear_coeffs(dst, src, vertical)))


def test_aa_tables_match_oracle():
    for i, o in [(1024, 640), (768, 480), (1024, 100), (64, 128), (1024, 2000), (512, 400), (96, 51), (128, 50)]:
        s, w = geometry.aa_tables(i, o)
        os_, ow = opp._aa_weights(i, o)
        np.testing.assert_array_equal(s, np.array(os_))
        for r in range(o):
            n = len(ow[r])
            np.testing.assert_array_equal(w[r, :n], ow[r])
            assert (w[r, n:] == 0).all()


def test_resize_mode():
    assert geometry.resize_mode(1024, 1024, 1024, 1024) == 0
    assert geometry.resize_mode(2048, 2048, 1024, 1024) == 1
    assert geometry.resize_mode(480, 640, 768, 1024) == 2


def test_packed_weights_reproduce_the_oracle(vitb_sd):
    """weights.pack_weights (BN folding, QKV fusion, tap-major convs, sub-pixel ConvT, merged heads, commuted out_conv)
    pushed through the CUDA data flow in fp32 on CPU must equal the oracle up to bf16 weight rounding."""
    from packed_ref import forward_packed
    from oracle import model as om
    from s3od_b200.weights import pack_weights
    torch.manual_seed(0)
    S = 64
    x = torch.randn(1, 3, S, S)
    pw = pack_weights(vitb_sd, VITB, S)
    # undo the bf16 rounding of the operands so only the packing logic is tested
    exact = pack_weights({k: v.double() if v.is_floating_point() else v for k, v in vitb_sd.items()}, VITB, S)
    assert set(exact) == set(pw)
    masks, iou = forward_packed(pw, x, VITB)
    ref = om.forward(vitb_sd, x.bfloat16().float(), VITB)
    assert float((masks - ref["pred_masks"]).abs().max()) < 0.15          # bf16 weights only
    assert float((iou - ref["pred_iou"]).abs().max()) < 0.03
    assert float((torch.sigmoid(masks) - torch.sigmoid(ref["pred_masks"])).abs().mean()) < 4e-3


def test_packed_names_cover_the_library_requirements(vitb_sd):
    from s3od_b200.weights import pack_weights
    pw = pack_weights(vitb_sd, VITB, 64)
    assert pw["enc.0.qkv.w"].shape == (2304, 768) and pw["enc.0.qkv.w"].dtype == torch.bfloat16
    assert float(pw["enc.0.qkv.b"][768:1536].abs().max()) == 0.0           # key_bias = false
    assert "enc.11.qkv.w" not in pw                                          # dead layer 12
    assert "head.ref4.rcu1.c1.w" not in pw                                   # unused unit
    assert pw["head.mh.up.w"].shape == (256, 512) and pw["head.mh.heads.w"].shape == (96, 576)
    assert pw["head.rs0.w"].shape == (4096, 256) and pw["head.rs1.w"].shape == (2048, 512)
    # both checkpoint key layouts are accepted
    old = {k.replace("encoder.model.layer.", "encoder.layer."): v for k, v in vitb_sd.items()}
    pw2 = pack_weights(old, VITB, 64)
    assert torch.equal(pw2["enc.3.up.w"], pw["enc.3.up.w"])


def test_normalisation_lut_is_bf16_of_reference_values():
    from s3od_b200.weights import normalisation_lut
    lut = normalisation_lut().float().reshape(3, 256)
    img = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, axis=2)
    ref = ((img.astype(np.float32) / 255.0 - opp.MEAN) / opp.STD).astype(np.float32)      # predictor.py:91
    np.testing.assert_array_equal(lut.numpy(), torch.from_numpy(ref[0].T.copy()).bfloat16().float().numpy())


def test_result_dataclass_and_bad_model_path():
    """Reference tests/test_inference_package.py:24-47."""
    import s3od_b200
    from PIL import Image
    r = s3od_b200.RemovalResult(predicted_mask=np.zeros((4, 4), np.float32), all_masks=np.zeros((3, 4, 4), np.float32),
                                all_ious=np.zeros(3, np.float32), rgba_image=Image.new("RGBA", (4, 4)))
    assert r.all_masks.shape == (3, 4, 4) and r.rgba_image.mode == "RGBA"
    with pytest.raises(ValueError):
        s3od_b200.BackgroundRemoval(model_id="/nonexistent/path/model.pt")
    assert s3od_b200.BackgroundRemoval.DEFAULT_MODEL_ID == "okupyn/s3od"
    assert s3od_b200.BackgroundRemoval.DEFAULT_CHECKPOINT_NAME == "s3od.pt"


def test_no_cpu_fallback():
    import s3od_b200
    with pytest.raises(RuntimeError):
        s3od_b200.BackgroundRemoval(model_id="/nonexistent/path/model.pt", device="cpu")


def test_shard_range_partitions_exactly():
    from s3od_b200.sharder import shard_range
    for n in (0, 1, 7, 32, 256, 257):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                b, e = shard_range(n, r, world)
                cover.extend(range(b, e))
                assert 0 <= e - b <= n // world + 1
            assert cover == list(range(n))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from s3od_b200 import sharder
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, lr, w = sharder.init_from_env(backend="gloo")
    b, e = sharder.shard_range(37, r, w)
    sharder.barrier()
    t_max = sharder.max_over_ranks(10.0 + r)
    n_sum = sharder.sum_over_ranks(e - b)
    q.put((r, b, e, t_max, n_sum))
    dist.destroy_process_group()


def test_sharder_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1:3] == (0, 19) and res[1][1:3] == (19, 37)
    assert all(r[3] == 11.0 and r[4] == 37.0 for r in res)               # max over ranks, total units


def test_chunk_schedule_covers_the_batch_and_ends_small():
    """remove_background_batch: full micro-batches, then the last one halved down to <= 4 images (short copy-out tail)."""
    from s3od_b200.predictor import chunk_schedule
    assert chunk_schedule(32, 16) == [(0, 8), (8, 24), (24, 28), (28, 32)]
    assert chunk_schedule(64, 32) == [(0, 8), (8, 40), (40, 56), (56, 60), (60, 64)]
    assert chunk_schedule(16, 16) == [(0, 8), (8, 12), (12, 16)]
    assert chunk_schedule(1, 16) == [(0, 1)]
    assert chunk_schedule(0, 16) == []
    for n in range(1, 70):
        for step in (1, 3, 8, 16):
            b = chunk_schedule(n, step)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
            assert all(0 < e - s0 <= step for s0, e in b)
            assert b[-1][1] - b[-1][0] <= 4


def test_strict_checkpoint_ingest_like_load_state_dict(vitb_sd):
    """predictor.py:76 loads strictly: a checkpoint of another architecture, a missing or an unexpected key raises
    RuntimeError before anything is packed (no silent garbage masks / out-of-bounds reads)."""
    from s3od_b200.arch import VITB, VITL
    from s3od_b200.weights import expected_shapes, pack_weights, validate_state_dict
    assert len(expected_shapes(VITB)) == 371 == len(vitb_sd)                       # 355 fp32 + 16 int64 entries
    validate_state_dict(vitb_sd, VITB)
    with pytest.raises(RuntimeError, match="Missing key"):
        pack_weights(vitb_sd, VITL, 64)                                            # ViT-B weights, ViT-L model
    bad = dict(vitb_sd)
    bad.pop("seg_head.classifier_head.4.bias")
    with pytest.raises(RuntimeError, match="Missing key"):
        validate_state_dict(bad, VITB)
    bad = dict(vitb_sd)
    bad["seg_head.extra.weight"] = torch.zeros(3)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        validate_state_dict(bad, VITB)
    bad = dict(vitb_sd)
    bad["seg_head.scratch.layer1_rn.weight"] = torch.zeros(128, 256, 3, 3)         # features != 256
    with pytest.raises(RuntimeError, match="size mismatch"):
        validate_state_dict(bad, VITB)
    from dataclasses import replace
    with pytest.raises(RuntimeError, match="size mismatch|Missing key"):
        validate_state_dict(vitb_sd, replace(VITB, num_outputs=1))
    # the older transformers layout (encoder.layer.N.*) is the same checkpoint under other names
    old = {k.replace("encoder.model.layer.", "encoder.layer."): v for k, v in vitb_sd.items()}
    validate_state_dict(old, VITB)


def test_zero_pad_non_square_geometry_is_refused():
    """A source one pixel off square (1025 x 1024) resizes to 1024 x 1023 with BOTH pads 0: the reference then feeds the
    network a non-square input (predictor.py:88-89).  The square launch plan cannot reproduce that, so it raises."""
    from s3od_b200.utils import check_padding, get_pad_info
    for hw in ((1025, 1024), (1024, 1023)):
        pad = get_pad_info(np.empty(hw + (0,), np.uint8), 1024)
        assert pad["height_pad"] == 0 and pad["width_pad"] == 0 and pad["resized_size"] != (1024, 1024)
        with pytest.raises(ValueError, match="non-square"):
            check_padding(pad, 1024)
    check_padding(get_pad_info(np.empty((2048, 2048, 0), np.uint8), 1024), 1024)
