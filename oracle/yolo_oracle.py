"""CPU oracle for the YOLO grid path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's algorithm for the
non-convolutional part of the detection hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package never does
(it fails loudly if the CUDA library is missing).

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the *real*
reference modules (``/root/reference``: ``datasets/utils.py``,
``losses/YoloLoss.py``, ``datasets/WIDERFace/dataset.py``) plus the installed
``torchvision.ops.nms`` (0.26.0 CPU; the reference pins 0.11.2 whose source is
not available here) and stores their outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against them
bit-for-bit (integer work) / to 1e-6 (loss).

All arithmetic is carried out in IEEE binary32 with one rounding per
operation, in the operation order of the reference (no fused multiply-add),
because the decoded box corners are rounded half-to-even afterwards and a
1-ulp difference can flip a rounded coordinate.

Reference citations are ``file:line`` under ``/root/reference``.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------
# grid-cell assignment          datasets/WIDERFace/dataset.py:32-64
# --------------------------------------------------------------------------
def grid_encode(boxes: np.ndarray, num_of_patches: int, width: int, height: int) -> np.ndarray:
    """boxes ``[K,5]`` = (1, x, y, w, h) f32 -> feature map ``[5,S,S]`` f32.

    dataset.py:35-39  patch sizes are python floats ``width/S``, ``height/S``.
    dataset.py:43     ``i = floor(bx[1] / ps_x)`` -- f32 tensor / python float is an
                      f32 division, ``math.floor`` of the f32 result.
    dataset.py:51-52  offsets use the UN-clamped i, j: ``bx - i*ps`` where ``i*ps`` is a
                      python double that torch rounds to f32 before the f32 subtract.
    dataset.py:55-59  ``/ps`` (f32), ``w/width``, ``h/height`` (f32).
    dataset.py:61-63  clamp i, j to [0,S-1]; ``fm[:, i, j] = box`` -- later boxes overwrite.
    """
    S = int(num_of_patches)
    fm = np.zeros((5, S, S), dtype=F32)
    psx = width / S
    psy = height / S
    for bx in np.asarray(boxes, dtype=F32).reshape(-1, 5):
        i = math.floor(float(F32(bx[1]) / F32(psx)))
        j = math.floor(float(F32(bx[2]) / F32(psy)))
        nb = bx.copy()
        nb[1] = F32(F32(nb[1]) - F32(i * psx))
        nb[2] = F32(F32(nb[2]) - F32(j * psy))
        nb[1] = F32(nb[1] / F32(psx))
        nb[2] = F32(nb[2] / F32(psy))
        nb[3] = F32(nb[3] / F32(width))
        nb[4] = F32(nb[4] / F32(height))
        ci = min(max(i, 0), S - 1)
        cj = min(max(j, 0), S - 1)
        fm[:, ci, cj] = nb
    return fm


# --------------------------------------------------------------------------
# decode + threshold            datasets/utils.py:111-126,152-163
# --------------------------------------------------------------------------
def decode_candidates(x: np.ndarray, probability_threshold: float, width: int, height: int,
                      num_of_patches: int):
    """``x[5,S,S]`` -> (scores[K], rounded xyxy[K,4], cell index[K]) in row-major (i,j) order.

    utils.py:108-109  ``ps = width / num_of_patches`` (python float; NOT derived from x.shape --
                      SeparableCNN decodes a 10x10 map with ps=30, reproduce that).
    utils.py:112,119  ``where(x[0] > thr)``: thr is rounded to f32 by torch, row-major order.
    utils.py:122-125  ``X = x1*ps_x + i*ps_x`` (first spatial index pairs with x), two f32
                      multiplies then one f32 add; ``W = x3*width``; ``H = x4*height``.
    utils.py:153-154  ``x2 = W + X``; ``y2 = H + Y`` (un-rounded operands).
    utils.py:162      ``torch.round`` = half-to-even on [X, Y, x2, y2].
    """
    x = np.asarray(x, dtype=F32)
    S1, S2 = x.shape[1], x.shape[2]
    psx = F32(width / num_of_patches)
    psy = F32(height / num_of_patches)
    thr = F32(probability_threshold)
    ii, jj = np.nonzero(x[0] > thr)          # row-major, like torch.where
    conf = x[0, ii, jj]
    X = (x[1, ii, jj] * psx).astype(F32) + (ii.astype(F32) * psx).astype(F32)
    Y = (x[2, ii, jj] * psy).astype(F32) + (jj.astype(F32) * psy).astype(F32)
    W = (x[3, ii, jj] * F32(width)).astype(F32)
    H = (x[4, ii, jj] * F32(height)).astype(F32)
    X = X.astype(F32)
    Y = Y.astype(F32)
    x2 = (W + X).astype(F32)
    y2 = (H + Y).astype(F32)
    boxes = np.stack([X, Y, x2, y2], axis=1).astype(F32) if len(ii) else np.zeros((0, 4), F32)
    boxes = np.rint(boxes).astype(F32)       # rint == round-half-even
    cells = (ii * S2 + jj).astype(np.int32)
    return conf.astype(F32), boxes, cells


# --------------------------------------------------------------------------
# NMS   torchvision.ops.nms (torchvision/csrc/ops/cpu/nms_kernel.cpp, 0.26.0 installed;
#       called from datasets/utils.py:87,164)
# --------------------------------------------------------------------------
def nms(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    """Greedy NMS, returns kept indices (int64) in descending-score order.

    * order = stable sort of scores, descending (ties: lower index first)
    * areas = (x2-x1)*(y2-y1), no +1
    * inter = max(0, xx2-xx1) * max(0, yy2-yy1); ovr = inter / (area_i + area_j - inter)
    * j is suppressed iff ovr > thr (strict).  0/0 = NaN compares false -> kept.
    All f32, one rounding per op; thr is rounded to f32 (the C++ op takes a double and
    compares ``ovr > iou_threshold`` with ovr a float -> promoted compare; using the f32-rounded
    threshold differs only when ovr lies strictly between thr and f32(thr), see note below).
    """
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=F32).reshape(-1)
    n = boxes.shape[0]
    order = np.argsort(-scores.astype(np.float64), kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = ((x2 - x1).astype(F32) * (y2 - y1).astype(F32)).astype(F32)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = float(iou_threshold)               # compare in double like the C++ op
    with np.errstate(invalid="ignore", divide="ignore"):
        for _i in range(n):
            i = order[_i]
            if suppressed[i]:
                continue
            keep.append(i)
            for _j in range(_i + 1, n):
                j = order[_j]
                if suppressed[j]:
                    continue
                xx1 = max(x1[i], x1[j]); yy1 = max(y1[i], y1[j])
                xx2 = min(x2[i], x2[j]); yy2 = min(y2[i], y2[j])
                w = max(F32(0), F32(xx2 - xx1)); h = max(F32(0), F32(yy2 - yy1))
                inter = F32(w * h)
                ovr = F32(inter / F32(F32(areas[i] + areas[j]) - inter))
                if float(ovr) > thr:
                    suppressed[j] = True
    return np.asarray(keep, dtype=np.int64)


def reduce_bounding_boxes(x: np.ndarray, probability_threshold: float, iou_threshold: float,
                          input_shape, num_of_patches: int, return_index: bool = False):
    """ReduceBoundingBoxes.forward  datasets/utils.py:157-170.

    Returns ``[K',5]`` f32 rows (score, X, Y, x2-X, y2-Y) from the ROUNDED corners
    (utils.py:165-166), in NMS keep order; empty -> shape (0,5) (utils.py:170).
    ``input_shape`` = (C, width, height) -- utils.py:107 unpacks it in that order.
    """
    _, width, height = input_shape
    scores, boxes, cells = decode_candidates(x, probability_threshold, width, height, num_of_patches)
    if scores.shape[0] == 0:
        out = np.zeros((0, 5), dtype=F32)
        return (out, np.zeros((0,), np.int32)) if return_index else out
    keep = nms(boxes, scores, iou_threshold)
    b = boxes[keep]
    out = np.stack([scores[keep], b[:, 0], b[:, 1], (b[:, 2] - b[:, 0]).astype(F32),
                    (b[:, 3] - b[:, 1]).astype(F32)], axis=1).astype(F32)
    if return_index:
        return out, cells[keep].astype(np.int32)
    return out


# --------------------------------------------------------------------------
# YoloLoss                      losses/YoloLoss.py:4-44
# --------------------------------------------------------------------------
def yolo_loss(pred: np.ndarray, gt: np.ndarray, dtype=np.float64):
    """Returns (loss, dloss/dpred) for one image, ``pred``/``gt`` = ``[5,S,S]``.

    YoloLoss.py:6      S = pred.shape[1]
    YoloLoss.py:8-9    if nansum(pred) != 0: NaNs in pred are replaced by 0.1 (grad 0 there)
    YoloLoss.py:17-18  gt_x,gt_y = gt[1],gt[2] but pred_y,pred_x = pred[1],pred[2]  (swapped!)
    YoloLoss.py:25     no-object weight = 1/S
    YoloLoss.py:27-38  L = sum 3 g0 [(g1-p2)^2 + (g2-p1)^2] + 3 g0 [(sqrt g3 - sqrt p3)^2 +
                       (sqrt g4 - sqrt p4)^2] + (g0 + (1-g0)/S)(g0-p0)^2
    The default float64 evaluation is the "exact" value; the f32 CUDA kernel and the f32
    reference are both compared against it with a stated relative tolerance.
    """
    S = pred.shape[1]
    p = np.asarray(pred, dtype=dtype).reshape(5, -1).copy()
    g = np.asarray(gt, dtype=dtype).reshape(5, -1)
    nanmask = np.isnan(p)
    if np.nansum(p) != 0:
        p[nanmask] = 0.1
    else:
        nanmask = np.zeros_like(nanmask)
    g0 = g[0]
    w_no = dtype(1.0) / dtype(S)
    with np.errstate(invalid="ignore", divide="ignore"):
        sp3, sp4 = np.sqrt(p[3]), np.sqrt(p[4])
        sg3, sg4 = np.sqrt(g[3]), np.sqrt(g[4])
        xy = 3 * g0 * ((g[1] - p[2]) ** 2 + (g[2] - p[1]) ** 2)
        wh = 3 * g0 * ((sg3 - sp3) ** 2 + (sg4 - sp4) ** 2)
        cw = g0 + (1 - g0) * w_no
        cf = cw * (g0 - p[0]) ** 2
        loss = np.sum(xy + wh + cf)
        d = np.zeros_like(p)
        d[0] = -2 * cw * (g0 - p[0])
        d[1] = -6 * g0 * (g[2] - p[1])
        d[2] = -6 * g0 * (g[1] - p[2])
        d[3] = -3 * g0 * (sg3 - sp3) / sp3
        d[4] = -3 * g0 * (sg4 - sp4) / sp4
    d[nanmask] = 0
    return loss, d.reshape(np.asarray(pred).shape)


def step_metrics(gt_rows: np.ndarray, pred_rows: np.ndarray, iou_thr: float = 0.5):
    """Per-image detection metrics of the reference's train/validation step (models/ModelMeta.py:184-214).

    gt_rows / pred_rows: ``[K,5]`` rows (score, x, y, w, h) as returned by ``reduce_bounding_boxes``.  Restates
    ``torchvision.ops.box_iou`` on the xyxy boxes built at ModelMeta.py:201-205 (x2 = w + x, y2 = h + y), with
    ``nan_to_num(., 0)`` (:206), in IEEE f32 like the tensor ops.  Returns (hits, iou_sum): the number of
    (gt, pred) pairs with IoU > iou_thr (:210-212) and the sum of all IoUs (:214).  The caller applies the
    reference's bookkeeping (nothing is accumulated for an image without predictions, :200).
    """
    f = np.float32
    g = gt_rows.astype(f).reshape(-1, 5)
    p = pred_rows.astype(f).reshape(-1, 5)
    if g.shape[0] == 0 or p.shape[0] == 0:
        return 0, f(0.0)
    gx1, gy1 = g[:, 1], g[:, 2]
    gx2, gy2 = (g[:, 3] + gx1).astype(f), (g[:, 4] + gy1).astype(f)
    px1, py1 = p[:, 1], p[:, 2]
    px2, py2 = (p[:, 3] + px1).astype(f), (p[:, 4] + py1).astype(f)
    ag = ((gx2 - gx1).astype(f) * (gy2 - gy1).astype(f)).astype(f)
    ap = ((px2 - px1).astype(f) * (py2 - py1).astype(f)).astype(f)
    w = np.maximum(f(0), (np.minimum(gx2[:, None], px2[None]) - np.maximum(gx1[:, None], px1[None])).astype(f))
    h = np.maximum(f(0), (np.minimum(gy2[:, None], py2[None]) - np.maximum(gy1[:, None], py1[None])).astype(f))
    inter = (w * h).astype(f)
    union = ((ag[:, None] + ap[None]).astype(f) - inter).astype(f)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = (inter / union).astype(f)
    iou = np.where(np.isnan(iou), f(0), iou)
    return int((iou > f(iou_thr)).sum()), iou.astype(np.float64).sum()
