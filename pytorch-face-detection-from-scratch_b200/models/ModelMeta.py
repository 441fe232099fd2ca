"""Training-harness mirror of the reference's ``models/ModelMeta.py`` for the hot path it drives.

pytorch_lightning is not a dependency here, so ``ModelMeta`` is a plain nn.Module exposing the same
methods the Lightning loop calls (``forward``, ``configure_optimizers``, ``training_step``,
``validation_step``).  ``step`` keeps the reference's definition (ModelMeta.py:115-227): forward,
SUM over the batch of ``yolo_loss`` (not a mean, :215), decode + NMS of ground truth and prediction,
recall / precision / IoU -- but loss and decode each run as ONE batched kernel instead of a
per-sample Python loop with several host syncs per image.
"""
from __future__ import annotations

from pathlib import Path

import torch
import torch.nn as nn
from torchvision.ops import box_iou

from .. import ops
from ..datasets.utils import ReduceBoundingBoxes
from ..losses.YoloLoss import yolo_loss, yolo_loss_batch  # noqa: F401  (same import surface as the reference)


class ModelMeta(nn.Module):
    def __init__(self, model, lr=1e-4, pretrained=False, log_path=Path("out.log"), *args, **kwargs):
        super().__init__()
        self.model = model
        self.lr = lr
        self.automatic_optimization = True
        self.log_path = log_path
        self.current_epoch = 0
        self.logged = {}

    def log(self, name, value, **kw):
        self.logged[name] = value

    def forward(self, x):
        return self.model(x)

    def configure_optimizers(self):
        """ModelMeta.py:104-112.  The reference's SAMSGD perturbs and un-perturbs the weights without
        recomputing gradients, i.e. it is numerically plain Adam(lr); we return that."""
        optimizer = torch.optim.Adam(self.parameters(), lr=self.lr)
        self.opt = optimizer
        scheduler = torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=[40], gamma=0.1)
        return [optimizer], [scheduler]

    def step(self, batch, batch_idx, validation=False):
        x, y, gt_bbxs = batch
        y_hat = self.forward(x)
        loss = yolo_loss_batch(y_hat, y)                                   # ModelMeta.py:173-176
        rb = self.model.reduce_bounding_boxes
        if isinstance(rb, ReduceBoundingBoxes) and y_hat.is_cuda:
            # ModelMeta.py:184-214 for the whole batch: two decode+NMS launches and one IoU-metrics launch, one D2H
            # copy of [B,4] (hits, sum IoU, n_gt, n_pred) instead of ~10 host syncs per image.
            with torch.no_grad():
                gt_b, gt_n = rb.batch_forward(y)
                pr_b, pr_n = rb.batch_forward(y_hat.detach())
                m = torch.empty((y.shape[0], 4), dtype=torch.float32, device=y.device)
                ops.box_metrics(gt_b, gt_n, pr_b, pr_n, 0.5, m)
            hits, iou_sum, n_gt, n_pred = m.double().unbind(1)
            has_pred = n_pred > 0
            recall = torch.where(n_gt > 0, hits / n_gt.clamp(min=1), torch.zeros_like(hits))   # gt empty, pred not: 0
            total_recall = float((recall * has_pred).sum())
            total_precision = float((hits / n_pred.clamp(min=1) * has_pred).sum())
            total_iou = (iou_sum * has_pred).sum().float()
        else:
            total_iou, total_recall, total_precision = self._metrics_per_image(y, y_hat)
        n = len(y)
        step_outputs = {"loss": loss, "total_iou": total_iou / n, "total_recall": total_recall / n,
                        "total_precision": total_precision / n}
        self.log("step_loss", loss, prog_bar=True, logger=True, on_step=True)
        return step_outputs

    def _metrics_per_image(self, y, y_hat):
        """The reference's per-image loop, used when ``reduce_bounding_boxes`` was replaced by a user callable."""
        with torch.no_grad():
            gt_all = self.model.non_max_suppression(y)
            pred_all = self.model.non_max_suppression(y_hat.detach())
        total_iou = 0.0
        total_recall = 0.0
        total_precision = 0.0
        for gt_bbx, pred_bbx in zip(gt_all, pred_all):
            gt_bbx = gt_bbx[:, 1:].to(y.device)
            if pred_bbx.shape[0] > 0:
                pred_bbx = pred_bbx[:, 1:].to(y.device)
                gt_bbx[:, 2] = gt_bbx[:, 2] + gt_bbx[:, 0]
                gt_bbx[:, 3] = gt_bbx[:, 3] + gt_bbx[:, 1]
                pred_bbx[:, 2] = pred_bbx[:, 2] + pred_bbx[:, 0]
                pred_bbx[:, 3] = pred_bbx[:, 3] + pred_bbx[:, 1]
                iou = torch.nan_to_num(box_iou(gt_bbx, pred_bbx), 0)
                hits = torch.where(iou > 0.5)[0].shape[0]
                if gt_bbx.shape[0] == 0:
                    recall = 1.0 if pred_bbx.shape[0] == 0 else 0.0
                else:
                    recall = hits / gt_bbx.shape[0]
                total_recall += recall
                total_precision += hits / pred_bbx.shape[0]
                total_iou += torch.sum(iou)
        return total_iou, total_recall, total_precision

    def to_torchscript(self, file_path=None, method="script", example_inputs=None, **kwargs):
        """train_model.py:61 ``model_setup.to_torchscript(model_save_path)`` (LightningModule API): exports the wrapped
        detector."""
        return self.model.to_torchscript(file_path)

    def training_step(self, batch, batch_idx):
        return self.step(batch, batch_idx)

    def validation_step(self, batch, batch_idx):
        return self.step(batch, batch_idx, validation=True)

    def test_step(self, batch, batch_idx):
        return self.step(batch, batch_idx, validation=True)
