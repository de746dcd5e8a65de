"""Generate tests/golden/metrics.npz from the REAL reference metrics (TEST INFRASTRUCTURE - see oracle/__init__.py).

Run in the build container only:   python -m oracle.make_golden_metrics
Loads /root/reference/synth_sod/src/synth_sod/model_training/metrics.py (unmodified) by path and records, per case, the
values EvaluationMetrics.step appends to its lists (mae, max_f, avg_f, s_score) and the 256 changeable E-measure values
of EMeasure.step (metrics.py:23-33)."""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference/synth_sod/src/synth_sod/model_training/metrics.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cases():
    g = torch.Generator().manual_seed(5)
    out = []
    for (h, w) in ((64, 64), (96, 128), (50, 70)):
        yy, xx = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")
        blob = (((yy - 0.45 * h) / (0.3 * h)) ** 2 + ((xx - 0.55 * w) / (0.25 * w)) ** 2 < 1).float()
        pred = (0.8 * blob + 0.25 * torch.rand(h, w, generator=g)).clamp(0, 1)
        out.append((f"blob_{h}x{w}", pred, blob))
        out.append((f"noise_{h}x{w}", torch.rand(h, w, generator=g), (torch.rand(h, w, generator=g) > 0.6).float()))
    h, w = 40, 56
    out.append(("empty_mask", torch.rand(h, w, generator=g), torch.zeros(h, w)))
    out.append(("full_mask", torch.rand(h, w, generator=g), torch.ones(h, w)))
    out.append(("binary_pred", (torch.rand(h, w, generator=g) > 0.5).float(), (torch.rand(h, w, generator=g) > 0.5).float()))
    return out


def main():
    spec = importlib.util.spec_from_file_location("ref_metrics", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rec = {}
    names = []
    for name, pred, mask in cases():
        em = ref.EvaluationMetrics(device=None)
        em.step(pred.clone(), mask.clone())
        names.append(name)
        rec[name + "_pred"], rec[name + "_mask"] = pred.numpy(), mask.numpy()
        rec[name + "_vals"] = np.array([em.metrics["mae"][0], em.metrics["max_f"][0], em.metrics["avg_f"][0], em.metrics["s_score"][0]], np.float64)
        rec[name + "_ems"] = np.asarray(em.emeasure.metrics["changeable_ems"][0], np.float64)
        rec[name + "_wfm"] = np.float64(em.weighted_fmeasure.metrics["weighted_fms"][0])        # WeightedFMeasure.step (metrics.py:146-153)
        em2 = ref.EvaluationMetrics(device=None, sm_only=True)
        em2.step(pred.clone(), mask.clone())
        assert abs(em2.metrics["s_score"][0] - em.metrics["s_score"][0]) < 1e-7
        print(name, rec[name + "_vals"])
    full = ref.EvaluationMetrics(device=None)
    for name, pred, mask in cases():
        full.step(pred.clone(), mask.clone())
    allm = full.compute_metrics()
    rec["em_all_cases"] = np.float64(allm["Em"])
    rec["wf_all_cases"] = np.float64(allm["wF"])
    rec["names"] = np.array(names)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics.npz"), **rec)


if __name__ == "__main__":
    main()
