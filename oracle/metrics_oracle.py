"""numpy restatement of the three full-reference quality metrics named by the parity bar
(PSNR / SAM / ERGAS within 0.01 of the reference).  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/models/base/metrics.py: sam :22-35, psnr :39-48, ergas :166-182.
Inputs are HWC arrays in the de-normalised range (value * 2047.5, dataset/utils.py:252-263).
Pinned by tests/golden (metrics computed by the reference's own functions on the same arrays)."""
import numpy as np

DYNAMIC_RANGE = 2047.5            # metrics.py:19
_EPS = np.finfo(np.float64).eps


def sam(a: np.ndarray, b: np.ndarray) -> float:
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    dot = (a * b).sum(axis=2)
    na = np.sqrt((a * a).sum(axis=2))
    nb = np.sqrt((b * b).sum(axis=2))
    cos = (dot / (na * nb + _EPS)).clip(min=0, max=1)
    return float(np.mean(np.arccos(cos)))


def psnr(a: np.ndarray, b: np.ndarray, dynamic_range: float = DYNAMIC_RANGE) -> float:
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = np.mean(d * d)
    if mse <= 1e-10:
        return float("inf")
    return float(20 * np.log10(dynamic_range / (np.sqrt(mse) + _EPS)))


def ergas(fake: np.ndarray, real: np.ndarray, scale: int = 4) -> float:
    fake = fake.astype(np.float64)
    real = real.astype(np.float64)
    c = real.shape[2]
    means = real.reshape(-1, c).mean(axis=0)
    mses = ((fake - real) ** 2).reshape(-1, c).mean(axis=0)
    return float(100 / scale * np.sqrt((mses / (means ** 2 + _EPS)).mean()))


def evaluate(pred_nchw: np.ndarray, gt_nchw: np.ndarray, bit_depth: int = 11) -> np.ndarray:
    """Mean [PSNR, SAM, ERGAS] over a batch of NCHW images normalised to [0,1): de-normalised with 2**bit_depth - .5
    (dataset/utils.py:252-263); the PSNR peak stays the module constant 2047.5 (metrics.py:19,39)."""
    rows = []
    scale = np.float32(2 ** bit_depth - 0.5)
    for p, g in zip(pred_nchw, gt_nchw):
        p = np.transpose(p, (1, 2, 0)) * scale
        g = np.transpose(g, (1, 2, 0)) * scale
        rows.append([psnr(p, g), sam(p, g), ergas(p, g)])
    return np.asarray(rows).mean(axis=0)
