"""TorchScript export of the detectors: ``train_model.py:61`` (``model_setup.to_torchscript(path)``),
``demo_scripts/convert_checkpoint_to_scripted_model.py:51-54`` (``torch.jit.script(model)``) and the consumer
``demo_model.py:11-21`` (``torch.jit.load(path)(uint8[2,3,480,480], predict=torch.tensor(1))``).

The hot path is hand-written CUDA behind a C ABI, which the TorchScript compiler cannot see into, so the entry points
are registered with the PyTorch dispatcher as custom operators (library ``fd_b200``):

    fd_b200::detector_forward(Tensor x, Tensor flat_state, str family, int[] cfg, float p_thr, float iou_thr, bool predict) -> Tensor

``ScriptedDetector`` is a small scriptable ``nn.Module`` with the reference's ``forward(x, predict=torch.tensor(0))``
signature that carries the weights (one flat fp32 tensor in ``state_dict()`` order) and calls that operator;
``to_torchscript(model, path)`` scripts and saves it.  Loading needs this package imported first (the import registers
the operators) and a CUDA device: the archive executes the same kernels as the eager model, there is no CPU path.
Host tensors are accepted (copied to the GPU, result copied back), so ``extract_face`` of demo_model.py works as written
-- except that the reference demo hides the GPU (``CUDA_VISIBLE_DEVICES=""``, demo_model.py:8), which has to go.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn as nn

_LIB = torch.library.Library("fd_b200", "DEF")
_LIB.define("detector_forward(Tensor x, Tensor flat_state, str family, int[] cfg, float p_thr, float iou_thr, "
            "bool predict) -> Tensor")

_CACHE: Dict[Tuple, Tuple[nn.Module, int, int]] = {}


def _build(family: str, cfg: List[int], p_thr: float, iou_thr: float):
    from . import models
    if family == "PoolResnet":
        f, c, h, w, S, nb = cfg
        return models.PoolResnet.PoolResnet(f, (c, h, w), S, num_of_residual_blocks=nb, probability_threshold=p_thr,
                                            iou_threshold=iou_thr)
    if family == "Resnet":
        f, c, h, w, S, nb = cfg
        return models.Resnet.Resnet(f, (c, h, w), S, num_of_residual_blocks=nb, probability_threshold=p_thr,
                                    iou_threshold=iou_thr)
    if family == "SeparableCNN":
        f, c, h, w, S, nb = cfg
        return models.SeparableCNN.SeparableCNN(f, (c, h, w), num_of_residual_blocks=nb, probability_threshold=p_thr,
                                                iou_threshold=iou_thr)
    if family == "MobilenetV3Backbone":
        f, c, h, w, S, nb = cfg
        return models.MobilenetV3Backbone.MobilenetV3Backbone(f, (c, h, w), S, probability_threshold=p_thr,
                                                              iou_threshold=iou_thr)
    if family == "SSD":
        f, c, h, w, S, nb = cfg
        return models.SSD.SSD(f, (c, h, w), probability_threshold=p_thr, iou_threshold=iou_thr)
    raise ValueError(f"unknown detector family {family!r}")


def family_and_cfg(model) -> Tuple[str, List[int]]:
    family = type(model).__name__
    c, h, w = model.input_shape
    nb = len(model.residual_blocks) if hasattr(model, "residual_blocks") else 0
    f = getattr(model, "min_filters", None) or (model.conv1.out_channels if hasattr(model, "conv1") else 0)
    return family, [int(f), int(c), int(h), int(w), int(getattr(model, "num_of_patches", 0) or 0), int(nb)]


def flatten_state(model) -> torch.Tensor:
    """Parameters and buffers in ``state_dict()`` order as one flat fp32 tensor."""
    return torch.cat([v.detach().reshape(-1).float() for v in model.state_dict().values()])


def _load_flat(model, flat: torch.Tensor):
    sd, off = {}, 0
    for k, v in model.state_dict().items():
        n = v.numel()
        sd[k] = flat[off:off + n].view(v.shape).to(v.dtype)
        off += n
    if off != flat.numel():
        raise RuntimeError(f"flat state has {flat.numel()} elements, the {type(model).__name__} needs {off}")
    model.load_state_dict(sd, strict=True)


def _detector_forward(x, flat_state, family, cfg, p_thr, iou_thr, predict):
    if not torch.cuda.is_available():
        raise RuntimeError("fd_b200::detector_forward needs a CUDA device (the archive runs the sm_100a kernels; "
                           "there is no CPU path)")
    dev = flat_state.device if flat_state.is_cuda else torch.device("cuda", torch.cuda.current_device())
    key = (family, tuple(int(v) for v in cfg), float(p_thr), float(iou_thr), dev)
    ent = _CACHE.get(key)
    if ent is None or ent[1] != flat_state.data_ptr() or ent[2] != flat_state._version:
        model = ent[0] if ent is not None else _build(family, list(cfg), p_thr, iou_thr)
        _load_flat(model, flat_state.detach())
        model = model.to(dev).eval()
        _CACHE[key] = (model, flat_state.data_ptr(), flat_state._version)
    model = _CACHE[key][0]
    on_host = not x.is_cuda
    with torch.no_grad():
        out = model(x.to(dev) if on_host else x, predict=torch.tensor(1 if predict else 0))
    if isinstance(out, tuple):           # SSD predict: ragged tuple -> image 0, like the YOLO models
        out = out[0]
    return out.cpu() if on_host else out


_LIB.impl("detector_forward", _detector_forward, "CompositeExplicitAutograd")


class ScriptedDetector(nn.Module):
    """Scriptable carrier of a detector's weights with the reference's ``forward(x, predict=torch.tensor(0))``."""

    def __init__(self, model):
        super().__init__()
        self.family, self.cfg = family_and_cfg(model)
        self.p_thr = float(model.reduce_bounding_boxes.probability_threshold)
        self.iou_thr = float(model.reduce_bounding_boxes.iou_threshold)
        self.flat_state = nn.Parameter(flatten_state(model), requires_grad=False)
        self.state_names: List[str] = list(model.state_dict().keys())

    def forward(self, x: torch.Tensor, predict: torch.Tensor = torch.tensor(0)) -> torch.Tensor:
        return torch.ops.fd_b200.detector_forward(x, self.flat_state, self.family, self.cfg, self.p_thr, self.iou_thr,
                                                  bool(torch.eq(predict, 1)))


def to_torchscript(model, file_path=None):
    """``torch.jit.script`` of the detector (``LightningModule.to_torchscript`` semantics: eval mode, saved when
    ``file_path`` is given).  Returns the ScriptModule."""
    scripted = torch.jit.script(ScriptedDetector(model).eval())
    if file_path is not None:
        torch.jit.save(scripted, str(file_path))
    return scripted
