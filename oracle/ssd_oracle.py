"""CPU oracle for the SSD head path -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference's SSD-side algorithms (SURVEY.md 8a rows 12, 13 and the grid
assignment of ``dataset_ssd.py``): multi-scale grid encoding, ``ReduceSSDBoundingBoxes``, hard-negative
mining and ``ssd_loss`` (value + gradients).  Only ``tests/`` and ``__graft_entry__.smoke()`` may import
it; the product package never does.

Parity status: PINNED against outputs of the real reference run in the build container
(``tests/golden/make_golden_extra.py`` -> ``ssd_encode.npz``, ``ssd_decode.npz``, ``ssd_loss.npz``; checked by
``tests/test_oracle_golden.py``): encode / decode rows / kept order / mining mask bit-exact, loss and gradients
to 1e-6 relative.

Arithmetic is IEEE binary32 with one rounding per operation in the reference's operation order, as in
``yolo_oracle``.  Reference citations are ``file:line`` under ``/root/reference``.
"""
from __future__ import annotations

import math

import numpy as np

from .yolo_oracle import nms

F32 = np.float32
PATCH_SIZES = (60, 30, 15, 7)


def num_priors(patch_sizes=PATCH_SIZES) -> int:
    return sum(ps * ps for ps in patch_sizes)


def ssd_priors(patch_sizes=PATCH_SIZES) -> np.ndarray:
    """datasets/utils.py:35-48: per scale ``[ps*ps, 4]`` rows ``(fl32(1/ps)*i, fl32(1/ps)*j, 0, 0)`` for the
    cell (i, j) in row-major order (``1 / ps * i`` is python-float x int64 tensor = an f32 multiply)."""
    out = []
    for ps in patch_sizes:
        inv = F32(1 / ps)
        ii, jj = np.meshgrid(np.arange(ps), np.arange(ps), indexing="ij")
        pr = np.zeros((ps * ps, 4), F32)
        pr[:, 0] = (inv * ii.reshape(-1).astype(F32)).astype(F32)
        pr[:, 1] = (inv * jj.reshape(-1).astype(F32)).astype(F32)
        out.append(pr)
    return np.concatenate(out, axis=0)


def ssd_grid_encode(boxes: np.ndarray, width: int, height: int, patch_sizes=PATCH_SIZES) -> np.ndarray:
    """datasets/WIDERFace/dataset_ssd.py:36-76 for every scale, concatenated like :134-139 -> ``[P,5]``.

    :41-43  ``x, w /= width``; ``y, h /= height`` (f32);  :46-49 patch size = python float ``1/ps``
    :52     ``i = floor(bx[1] / (1/ps))`` (f32 division by the f32-rounded python float)
    :59     score = ``1 - 0.001*ps`` (f32 tensor minus python double, rounded to f32)
    :63-68  offsets from the UN-clamped cell: ``(x - i*(1/ps)) / (1/ps)``
    :73-75  clamp the cell, later boxes overwrite
    """
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 5)
    maps = []
    for ps in patch_sizes:
        fm = np.zeros((5, ps, ps), F32)
        if boxes.shape[0]:
            b = boxes.copy()
            b[:, 1] = (b[:, 1] / F32(width)).astype(F32); b[:, 3] = (b[:, 3] / F32(width)).astype(F32)
            b[:, 2] = (b[:, 2] / F32(height)).astype(F32); b[:, 4] = (b[:, 4] / F32(height)).astype(F32)
            pz = 1 / ps
            for bx in b:
                i = math.floor(float(F32(bx[1]) / F32(pz)))
                j = math.floor(float(F32(bx[2]) / F32(pz)))
                nb = bx.copy()
                nb[0] = F32(F32(nb[0]) - F32(0.001 * ps))
                nb[1] = F32(F32(nb[1]) - F32(i * pz))
                nb[2] = F32(F32(nb[2]) - F32(j * pz))
                nb[1] = F32(nb[1] / F32(pz))
                nb[2] = F32(nb[2] / F32(pz))
                fm[:, min(max(i, 0), ps - 1), min(max(j, 0), ps - 1)] = nb
        maps.append(fm.transpose(1, 2, 0).reshape(-1, 5))
    return np.concatenate(maps, axis=0)


def reduce_ssd_bounding_boxes(x: np.ndarray, probability_threshold: float, iou_threshold: float, input_shape,
                              patch_sizes=PATCH_SIZES, with_priors: bool = False) -> np.ndarray:
    """ReduceSSDBoundingBoxes.forward, datasets/utils.py:56-92.  ``x[P,5]`` rows (score, x, y, w, h).

    :59-64  with priors: ``x, y *= fl32(1/ps)`` then ``x..h += priors``
    :65-66  ``x, w *= width``; ``y, h *= height``   (input_shape = (C, width, height), :23)
    :51     keep rows with ``score > thr``;  :69-70 ``x2 = w + x``, ``y2 = h + y``
    :82     ``round`` (half to even) of the four corners;  :85 nms;  :87 back to (x, y, x2-x, y2-y)
    """
    _, width, height = input_shape
    x = np.asarray(x, dtype=F32).copy()
    if with_priors:
        mult = np.concatenate([np.full(ps * ps, F32(1 / ps), F32) for ps in patch_sizes])
        pr = ssd_priors(patch_sizes)
        x[:, 1] = (x[:, 1] * mult).astype(F32)
        x[:, 2] = (x[:, 2] * mult).astype(F32)
        x[:, 1:5] = (x[:, 1:5] + pr).astype(F32)
    x[:, 1] = (x[:, 1] * F32(width)).astype(F32); x[:, 3] = (x[:, 3] * F32(width)).astype(F32)
    x[:, 2] = (x[:, 2] * F32(height)).astype(F32); x[:, 4] = (x[:, 4] * F32(height)).astype(F32)
    sel = np.nonzero(x[:, 0] > F32(probability_threshold))[0]
    if sel.size == 0:
        return np.zeros((0, 5), F32)
    c = x[sel]
    c[:, 3] = (c[:, 3] + c[:, 1]).astype(F32)
    c[:, 4] = (c[:, 4] + c[:, 2]).astype(F32)
    bbx = np.rint(c[:, 1:]).astype(F32)
    keep = nms(bbx, c[:, 0], iou_threshold)
    b = bbx[keep]
    return np.stack([c[keep, 0], b[:, 0], b[:, 1], (b[:, 2] - b[:, 0]).astype(F32),
                     (b[:, 3] - b[:, 1]).astype(F32)], axis=1).astype(F32)


def hard_negative_mining(loss: np.ndarray, labels: np.ndarray, neg_pos_ratio: int) -> np.ndarray:
    """losses/SSDLoss.py:27-54.  Per row: positives (label > 0) plus the ``ratio * num_pos`` negatives of
    LARGEST ``loss`` (stable descending order: ties -> lower index first, like ``Tensor.sort`` on CPU)."""
    loss = np.asarray(loss, dtype=F32).copy()
    pos = np.asarray(labels) > 0
    num_neg = pos.sum(axis=1, keepdims=True) * neg_pos_ratio
    loss[pos] = -np.inf
    idx = np.argsort(-loss.astype(np.float64), axis=1, kind="stable")
    orders = np.argsort(idx, axis=1, kind="stable")
    return pos | (orders < num_neg)


def ssd_loss(conf, loc, labels, gt_loc, neg_pos_ratio: int, dtype=np.float64):
    """losses/SSDLoss.py:57-86.  Returns (loss, dloss/dconf, dloss/dloc, mask).

    :69-70  mining on ``-log(conf)`` (f32);  :72-77 BCE summed over the mined set with ``clamp(conf, 1e-7, 1-1e-7)``
    and ``round(labels)`` as target;  :78-83 smooth-L1 (beta 1) summed over the positive priors;  :85-86 both
    divided by the number of positive priors of the WHOLE batch."""
    conf32 = np.asarray(conf, dtype=F32)
    labels = np.asarray(labels, dtype=F32)
    mask = hard_negative_mining(-np.log(conf32), labels, neg_pos_ratio)
    c = np.asarray(conf, dtype=dtype)
    t = np.rint(labels).astype(dtype)
    eps = dtype(F32(10 ** -7))
    lo, hi = eps, dtype(F32(1) - F32(10 ** -7))
    cc = np.clip(c, lo, hi)
    cls = np.where(mask, -(t * np.log(cc) + (1 - t) * np.log(1 - cc)), 0.0)
    inside = (c >= lo) & (c <= hi)
    dconf = np.where(mask & inside, -(t / cc) + (1 - t) / (1 - cc), 0.0)
    pos = labels > 0
    diff = np.asarray(loc, dtype=dtype) - np.asarray(gt_loc, dtype=dtype)
    ad = np.abs(diff)
    l1 = np.where(ad < 1, 0.5 * diff * diff, ad - 0.5) * pos[..., None]
    dloc = np.where(ad < 1, diff, np.sign(diff)) * pos[..., None]
    n_pos = pos.sum()
    with np.errstate(invalid="ignore", divide="ignore"):
        loss = (l1.sum() + cls.sum()) / dtype(n_pos)
        return loss, dconf / dtype(n_pos), dloc / dtype(n_pos), mask
