"""Host logic of the training-step slice (BASELINE.json configs[3]): the flat parameter layout, the bucketed gradient
all-reduce over a 2-rank gloo group on the CPU (the N > 1 path), and the oracle restatement of the reference loss against the
golden vectors produced by the UNMODIFIED reference loss.py (oracle/make_golden_loss.py)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import loss as ol                                   # noqa: E402
from s3od_b200.arch import VITB, VITL                          # noqa: E402
from s3od_b200.training import GradientAllReduce, ParameterLayout, gradless_keys, trainable_keys   # noqa: E402


def test_oracle_loss_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss.npz"))
    for n in g["names"]:
        z = torch.from_numpy(g[n + "_logits"]).requires_grad_(True)
        q = torch.from_numpy(g[n + "_iou_logits"]).requires_grad_(True)
        loss, parts = ol.loss_module(z, q, torch.from_numpy(g[n + "_masks"]), int(g[n + "_epoch"]))
        loss.backward()
        assert abs(float(loss) - float(g[n + "_loss"])) <= 1e-6
        assert float((z.grad - torch.from_numpy(g[n + "_grad_logits"])).abs().max()) <= 1e-9
        assert float((q.grad - torch.from_numpy(g[n + "_grad_iou"])).abs().max()) <= 1e-9
        for k in ("best_iou", "focal_loss_best", "iou_loss_best", "mse_ious_loss"):
            assert abs(float(parts[k]) - float(g[n + "_" + k])) <= 1e-6
        assert abs(float(parts["focal_loss_full"].mean()) - float(g[n + "_focal_loss_full"])) <= 1e-6
    for tag, lr in (("enc", 1e-5), ("head", 1e-4)):              # torch.optim.AdamW, three steps
        p = torch.from_numpy(g[f"adamw_{tag}_p0"])
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        for s in range(3):
            p, m, v = ol.adamw_step(p, torch.from_numpy(g[f"adamw_{tag}_grads"][s]), m, v, s + 1, lr)
            assert float((p - torch.from_numpy(g[f"adamw_{tag}_p{s + 1}"])).abs().max()) <= 2e-7


def test_parameter_layout_matches_the_reference_parameter_census():
    """28 tensors get no gradient; the all-reduce payload is 107.8 M fp32 values (431 MB)."""
    lay = ParameterLayout(VITB, bucket_bytes=64 << 20)
    assert len(lay.gradless) == 28 == len(gradless_keys(VITB))
    assert lay.numel_with_grad() == 107_825_862
    assert len(trainable_keys(VITB)) == len(lay.segments) + 28
    # reverse-autograd order: mask head / classifier first, patch embedding and tokens last
    assert lay.segments[0].name.startswith("seg_head.") and lay.segments[-1].name.startswith("encoder.embeddings.")
    offs = [s.offset for s in lay.segments]
    assert offs == sorted(offs) and all(o % 4 == 0 for o in offs)
    for a, b in zip(lay.segments, lay.segments[1:]):
        assert a.offset + a.numel <= b.offset                      # no overlap
    assert lay.bucket_offsets[0] == 0 and lay.bucket_offsets[-1] == lay.total
    sizes = np.diff(lay.bucket_offsets) * 4
    assert (sizes[:-1] <= (64 << 20) + 4 * 3072 * 768 * 4).all() and lay.num_buckets >= 6
    (h0, h1), (e0, e1) = lay.group_ranges[1], lay.group_ranges[0]
    assert h0 == 0 and h1 == e0 and e1 == lay.total                # two contiguous optimiser groups: head (lr x 10), encoder
    assert all((s.group == 1) == s.name.startswith("seg_head.") for s in lay.segments)
    big = ParameterLayout(VITL)
    assert len(big.gradless) == 28 and big.numel_with_grad() > 300_000_000


def _allreduce_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dataclasses import replace
        tiny = replace(VITB, hidden=64, layers=3, heads=1, mlp=128, taps=(1, 1, 2, 2), features=8, out_channels=(8, 8, 16, 16), inter_features=4)
        lay = ParameterLayout(tiny, bucket_bytes=16 << 10)
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.zeros(lay.total)
        grads = {s.name: torch.randn(s.numel, generator=g) for s in lay.segments}
        red = GradientAllReduce(lay, flat)
        for s in lay.segments:                                     # the backward pass produces gradients in this order
            flat[s.offset:s.offset + s.numel] = grads[s.name]
            red.mark_ready(s.name)
        red.finish()
        # expected: the sum over both ranks' seeded gradients
        want = torch.zeros(lay.total)
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            for s in lay.segments:
                want[s.offset:s.offset + s.numel] += torch.randn(s.numel, generator=gr)
        q.put((rank, bool(torch.allclose(flat, want, atol=1e-6)), red.launch_order, lay.num_buckets, red.world))
    finally:
        dist.destroy_process_group()


def test_bucketed_gradient_allreduce_two_ranks_gloo():
    """The N > 1 path on the CPU: every bucket is reduced exactly once, in the order its gradients complete."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, order, nb, world in res:
        assert ok, f"rank {rank}: reduced gradients differ from the sum"
        assert order == list(range(nb)) and nb >= 3 and world == 2


def test_single_process_allreduce_is_a_no_op():
    from dataclasses import replace
    tiny = replace(VITB, hidden=64, layers=3, heads=1, mlp=128, taps=(1, 1, 2, 2), features=8, out_channels=(8, 8, 16, 16), inter_features=4)
    lay = ParameterLayout(tiny, bucket_bytes=8 << 10)
    flat = torch.arange(lay.total, dtype=torch.float32)
    red = GradientAllReduce(lay, flat)
    for s in reversed(lay.segments):                               # any completion order launches each bucket once
        red.mark_ready(s.name)
    red.finish()
    assert sorted(red.launch_order) == list(range(lay.num_buckets)) and red.world == 1
    assert torch.equal(flat, torch.arange(lay.total, dtype=torch.float32))


def test_weight_gradient_split_plans():
    """Host logic of the split-K weight-gradient GEMMs (training.wgrad_plan / wgrad_plan_tn): every split keeps whole 64-row blocks,
    the (split, tile) items never exceed one wave of the 148 SMs, and short contractions are not split at all."""
    from s3od_b200.training import wgrad_plan, wgrad_plan_tn
    for n_out, n_in, rows in ((768, 768, 4101), (2304, 768, 16404), (768, 3072, 4101), (64, 640, 262144), (256, 2304, 65536), (96, 128, 700), (768, 768, 100)):
        splits, kpad = wgrad_plan(n_out, n_in, rows, 148)
        assert splits >= 1 and kpad >= rows and kpad % (64 * splits) == 0 and kpad - rows < 64 * splits
        tiles = ((n_out + 127) // 128 + 1) // 2 * (n_in // (256 if n_in % 256 == 0 else 128))
        assert splits == 1 or splits * tiles <= 74
        s_tn = wgrad_plan_tn(n_out, n_in, rows, 148)
        tiles_tn = ((n_out + 127) // 128) * ((n_in + 255) // 256)
        assert s_tn >= 1 and (s_tn == 1 or (s_tn * tiles_tn <= 148 and rows // s_tn >= 256))
    assert wgrad_plan(768, 768, 100, 148)[0] == 1 and wgrad_plan_tn(768, 768, 100, 148) == 1
    assert wgrad_plan_tn(768, 768, 16404, 148) == 8          # 18 output tiles -> 8 splits = 144 items
