"""Generate tests/golden/vis_*.npz from the REAL reference visualizer (TEST INFRASTRUCTURE - see oracle/__init__.py).

Run in the build container only (needs /root/reference):   python -m oracle.make_golden_vis
Imports s3od.visualizer (unmodified) from /root/reference/src, and the two pure functions of demo/app.py:38-56 by
executing just their source lines (the module itself needs gradio, which is not installed).
"""
import ast
import os
import sys

import numpy as np

REF_SRC = "/root/reference/src"
REF_DEMO = "/root/reference/demo/app.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def _demo_functions():
    src = open(REF_DEMO).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("compute_mask_iou", "is_ambiguous")]
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), REF_DEMO, "exec"), ns)     # the reference's own code objects
    return ns["compute_mask_iou"], ns["is_ambiguous"]


def main():
    sys.path.insert(0, REF_SRC)
    from s3od.visualizer import visualize_removal, visualize_all_masks              # the reference, unmodified
    from s3od.predictor import RemovalResult
    compute_mask_iou, is_ambiguous = _demo_functions()
    rng = np.random.default_rng(7)
    for name, (h, w), k in (("vis_36x52_k3", (36, 52), 3), ("vis_31x45_k3", (31, 45), 3), ("vis_20x24_k1", (20, 24), 1),
                            ("vis_16x16_k4", (16, 16), 4)):
        image = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        # soft masks with exact 0 / 1 plateaus and values straddling 0.5; mask 1 is a perturbed copy of mask 0
        base = rng.random((h, w), dtype=np.float32)
        masks = np.stack([np.clip(base * 1.4 - 0.2 + 0.05 * i * rng.standard_normal((h, w)).astype(np.float32), 0, 1)
                          for i in range(k)]).astype(np.float32)
        if k >= 3:
            masks[2] = (rng.random((h, w)) > 0.5).astype(np.float32)               # an unrelated mask -> ambiguous
        res = RemovalResult(predicted_mask=masks[0], all_masks=masks, all_ious=np.zeros(k, np.float32), rgba_image=None)
        rec = dict(image=image, masks=masks,
                   green=np.array(visualize_removal(image, res)), white=np.array(visualize_removal(image, res, (255, 255, 255))),
                   odd=np.array(visualize_removal(image, res, (13, 77, 201))), grid=np.array(visualize_all_masks(image, res)),
                   ambiguous=np.array(is_ambiguous(masks)), ambiguous_095=np.array(is_ambiguous(masks, 0.95)),
                   ambiguous_pair=np.array(is_ambiguous(masks[:2], 0.5) if k >= 2 else False))
        if k >= 2:
            rec["ious"] = np.array([compute_mask_iou(masks[i], masks[j]) for i in range(k) for j in range(i + 1, k)], np.float64)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print("wrote", name, {kk: getattr(v, "shape", v) for kk, v in rec.items()})


if __name__ == "__main__":
    main()
