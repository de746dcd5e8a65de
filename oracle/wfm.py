"""numpy restatement of the reference's weighted F-measure (TEST INFRASTRUCTURE - see oracle/__init__.py).

Follows WeightedFMeasure.step / cal_wfm / matlab_style_gauss2D
(/root/reference/synth_sod/src/synth_sod/model_training/metrics.py:146-204) WITHOUT scipy: the exact Euclidean feature
transform `bwdist(gt == 0, return_indices=True)` is restated as the separable lower-envelope minimum, with the tie-breaking
scipy 1.x shows (among equidistant foreground pixels: the smallest column, then the smallest row - pinned against scipy in
tests/test_oracle_golden.py and against the unmodified reference through tests/golden/metrics.npz), and the 7 x 7 convolution
(`scipy.ndimage.convolve(..., mode="constant", cval=0)`: double accumulation, result cast to the float32 input type) as an
explicit zero-padded sum."""
import numpy as np

_EPS = np.spacing(1)


def feature_transform(gt: np.ndarray):
    """For every pixel the nearest pixel where gt is True: (squared distance, row index, column index)."""
    H, W = gt.shape
    rows = np.arange(H)[:, None]
    up = np.maximum.accumulate(np.where(gt, rows, -1), axis=0)
    dn = np.flipud(np.minimum.accumulate(np.flipud(np.where(gt, rows, 10 ** 9)), axis=0))
    dn = np.where(dn >= 10 ** 9, -1, dn)
    take_dn = (up < 0) | ((dn >= 0) & ((dn - rows) < (rows - up)))            # the upper one on a tie
    near = np.where(take_dn, dn, up)
    d2col = np.where(near >= 0, (near - rows).astype(np.int64) ** 2, 1 << 40)
    xs = np.arange(W)
    dx2 = (xs[:, None] - xs[None, :]).astype(np.int64) ** 2                   # [x, x']
    d2 = np.empty((H, W), np.int64)
    ix = np.empty((H, W), np.int64)
    for y in range(H):
        c = dx2 + d2col[y][None, :]
        k = np.argmin(c, axis=1)                                              # first minimum = smallest x'
        ix[y] = k
        d2[y] = c[xs, k]
    iy = near[rows.repeat(W, 1), ix]
    return d2, iy, ix


def gauss7(sigma: float = 5.0) -> np.ndarray:
    """matlab_style_gauss2D((7, 7), sigma=5), metrics.py:192-204."""
    y, x = np.ogrid[-3:4, -3:4]
    h = np.exp(-(x * x + y * y) / (2.0 * sigma * sigma))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    return h / h.sum()


def weighted_f(pred: np.ndarray, mask: np.ndarray, beta: float = 1.0) -> float:
    """WeightedFMeasure.step on one image (pred float32 in [0, 1]; gt = mask >= 0.5)."""
    pred = np.asarray(pred, np.float32)
    gt = np.asarray(mask) >= 0.5
    if not gt.any():
        return 0.0
    d2, iy, ix = feature_transform(gt)
    E = np.abs(pred - gt.astype(np.float32))
    Et = np.where(gt, E, E[iy, ix])
    K = gauss7()
    pad = np.pad(Et.astype(np.float64), 3)
    EA = np.zeros(Et.shape, np.float64)
    H, W = Et.shape
    for a in range(7):
        for b in range(7):
            EA += K[a, b] * pad[a:a + H, b:b + W]
    EA = EA.astype(np.float32)
    mn = np.where(gt & (EA < E), EA, E)
    B = np.where(~gt, 2 - np.exp(np.log(0.5) / 5 * np.sqrt(d2.astype(np.float64))), 1.0)
    Ew = mn * B
    n_fg = float(gt.sum())
    TPw = n_fg - Ew[gt].sum()
    FPw = Ew[~gt].sum()
    R = 1 - Ew[gt].mean()
    P = TPw / (TPw + FPw + _EPS)
    return float((1 + beta) * R * P / (R + beta * P + _EPS))
