"""SURVEY 8f rank 4 on the GPU: s3od_b200.metrics.EvaluationMetrics against the reference's golden values
(tests/golden/metrics.npz, written by oracle/make_golden_metrics.py from the unmodified reference) and against the CPU oracle
at 1024 x 1024.  Tolerance: 2e-6 absolute - the reference sums in float32, the device path in double."""
import os
import time

import numpy as np
import pytest
import torch

from oracle import metrics as om

pytestmark = pytest.mark.gpu
TOL = 2e-6


def test_metrics_match_reference_golden(golden_dir):
    from s3od_b200.metrics import EvaluationMetrics
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    em = EvaluationMetrics("cuda:0")
    sm = EvaluationMetrics("cuda:0", sm_only=True)
    for i, name in enumerate(g["names"]):
        pred, mask = torch.from_numpy(g[name + "_pred"]), torch.from_numpy(g[name + "_mask"])
        keep = mask.clone()
        em.step(pred, mask)
        sm.step(pred, mask)
        assert torch.equal(mask, keep)                                  # the caller's mask is left alone
        got = np.array([em.metrics[k][i] for k in ("mae", "max_f", "avg_f", "s_score")])
        assert np.abs(got - g[name + "_vals"]).max() <= TOL, (name, got, g[name + "_vals"])
        assert abs(sm.metrics["s_score"][i] - g[name + "_vals"][3]) <= TOL
        # E-measure: integer histograms + the reference's float64 formulas -> equal to rounding
        np.testing.assert_allclose(em.changeable_ems[i], g[name + "_ems"], rtol=1e-12, atol=1e-12)
        # weighted F-measure: exact feature transform (scipy's tie-breaking), double-accumulated filter and sums
        assert abs(em.weighted_fms[i] - float(g[name + "_wfm"])) <= 1e-9, (name, em.weighted_fms[i], float(g[name + "_wfm"]))
    out = em.compute_metrics()
    assert set(out) == {"MAE", "MaxF", "AvgF", "Sm", "Em", "wF"} and set(sm.compute_metrics()) == {"Sm"}
    assert abs(out["Em"] - float(g["em_all_cases"])) <= 1e-12
    assert abs(out["wF"] - float(g["wf_all_cases"])) <= 1e-9
    vals = np.stack([g[n + "_vals"] for n in g["names"]])
    assert abs(out["MAE"] - vals[:, 0].mean()) <= TOL and abs(out["Sm"] - vals[:, 3].mean()) <= TOL
    em.reset()
    assert em.metrics["mae"] == []


def test_metrics_fullsize_matches_oracle(capsys):
    from s3od_b200.metrics import EvaluationMetrics
    g = torch.Generator().manual_seed(11)
    H = W = 1024
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    mask = (((yy - 420) / 300) ** 2 + ((xx - 560) / 260) ** 2 < 1).float()
    pred = (0.75 * mask + 0.3 * torch.rand(H, W, generator=g)).clamp(0, 1)
    ref = om.step(pred, mask)
    em = EvaluationMetrics("cuda:0")
    d_pred, d_mask = pred.cuda(), mask.cuda()
    em.step(d_pred, d_mask)
    got = {k: em.metrics[k][0] for k in ("mae", "max_f", "avg_f", "s_score")}
    for k in got:
        assert abs(got[k] - ref[k]) <= 2e-5, (k, got[k], ref[k])      # float32 sums over 1 M pixels in the oracle
    from oracle.wfm import weighted_f
    wf_ref = weighted_f(pred[::4, ::4].numpy(), mask[::4, ::4].numpy())               # the numpy oracle is O(H W^2): a 256 x 256 sub-sample
    em_small = EvaluationMetrics("cuda:0")
    em_small.step(pred[::4, ::4].contiguous().cuda(), mask[::4, ::4].contiguous().cuda())
    assert abs(em_small.weighted_fms[0] - wf_ref) <= 1e-9, (em_small.weighted_fms[0], wf_ref)
    assert 0.0 < em.weighted_fms[0] < 1.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        em.step(d_pred, d_mask)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    with capsys.disabled():
        print(f"\n[metrics 1024x1024] step (stats + region + weighted-F passes, 3 small D2H): {dt * 1e3:.2f} ms per image")
