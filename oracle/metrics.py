"""CPU restatement of the torch half of the reference's EvaluationMetrics (TEST INFRASTRUCTURE - see oracle/__init__.py).

Follows /root/reference/synth_sod/src/synth_sod/model_training/metrics.py:213-421: MAE, the 255-threshold precision /
recall sweep (MaxF, AvgF, beta^2 = 0.3) and the S-measure (object + region terms).  The E-measure and weighted F-measure of
that class are numpy / scipy CPU code in the reference as well and are not part of the device path.
One function per image: returns dict(mae, max_f, avg_f, s_score) as Python floats, like the values the reference appends
to `self.metrics[...]` in `step`.
"""
import torch


def _eval_pr(pred, mask, num=255):                                      # metrics.py:316-327
    th = torch.linspace(0, 1 - 1e-10, num)
    prec, rec = torch.zeros(num), torch.zeros(num)
    for i in range(num):
        sel = (pred >= th[i]).float()
        tp = (sel * mask).sum()
        prec[i], rec[i] = tp / (sel.sum() + 1e-20), tp / (mask.sum() + 1e-20)
    return prec, rec


def _object(values):                                                    # metrics.py:338-344 on the selected pixels
    x = values.mean()
    return 2.0 * x / (x * x + 1.0 + values.std() + 1e-20)


def _s_object(pred, mask):                                              # metrics.py:329-336
    u = mask.mean()
    return u * _object(pred[mask == 1]) + (1 - u) * _object((1 - pred)[mask == 0])


def _ssim(p, m):                                                        # metrics.py:405-421
    n = p.numel()
    x, y = p.mean(), m.mean()
    sx = ((p - x) ** 2).sum() / (n - 1 + 1e-20)
    sy = ((m - y) ** 2).sum() / (n - 1 + 1e-20)
    sxy = ((p - x) * (m - y)).sum() / (n - 1 + 1e-20)
    alpha = 4 * x * y * sxy
    beta = (x * x + y * y) * (sx + sy)
    if alpha != 0:
        return alpha / (beta + 1e-20)
    return 1.0 if beta == 0 else 0


def centroid(mask):                                                     # metrics.py:358-378
    rows, cols = mask.shape
    if mask.sum() == 0:
        return round(cols / 2), round(rows / 2)
    total = mask.sum()
    X = torch.round((mask.sum(dim=0) * torch.arange(cols).float()).sum() / total)
    Y = torch.round((mask.sum(dim=1) * torch.arange(rows).float()).sum() / total)
    return int(X), int(Y)


def _s_region(pred, mask):                                              # metrics.py:346-356, 380-403
    h, w = mask.shape
    X, Y = centroid(mask)
    area = h * w
    w1, w2, w3 = X * Y / area, (w - X) * Y / area, X * (h - Y) / area
    w4 = 1 - w1 - w2 - w3
    q = [_ssim(pred[ys, xs], mask[ys, xs]) for ys, xs in ((slice(0, Y), slice(0, X)), (slice(0, Y), slice(X, w)),
                                                          (slice(Y, h), slice(0, X)), (slice(Y, h), slice(X, w)))]
    return w1 * q[0] + w2 * q[1] + w3 * q[2] + w4 * q[3]


def s_measure(pred, mask):                                              # metrics.py:259-273
    y = mask.mean()
    if y == 0:
        return float(1.0 - pred.mean())
    if y == 1:
        return float(pred.mean())
    m = (mask >= 0.5).float()
    q = 0.5 * _s_object(pred, m) + 0.5 * _s_region(pred, m)
    return max(float(q), 0.0)


def step(pred: torch.Tensor, mask: torch.Tensor) -> dict:
    pred, mask = pred.float(), mask.float()
    prec, rec = _eval_pr(pred, mask)
    f = 1.3 * prec * rec / (0.3 * prec + rec)
    f[f != f] = 0
    return dict(mae=float((pred - mask).abs().mean()), max_f=float(f.max()), avg_f=float(f.mean()), s_score=s_measure(pred, mask))
