"""B200-native detection hot path with the reference's Python interface.

The directory name contains hyphens (it follows the reference repo's name), so import it with

    import importlib; fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")

or call ``install_dropin()`` once and use the reference's own import paths unchanged:

    from models.PoolResnet import PoolResnet
    from models import BaseModel, ModelMeta
    from losses.YoloLoss import yolo_loss
    from datasets.utils import ReduceBoundingBoxes
"""
import importlib
import sys

from . import native, ops  # noqa: F401

_DROPIN = ["models", "models.BaseModel", "models.PoolResnet", "models.Resnet", "models.SeparableCNN", "models.ModelMeta", "losses",
           "losses.YoloLoss", "losses.SSDLoss", "datasets", "datasets.utils", "datasets.WIDERFace",
           "datasets.WIDERFace.dataset", "datasets.WIDERFace.dataset_ssd"]


def install_dropin():
    """Alias the reference's top-level module names (models, losses, datasets) to this package, so
    train_model.py / demo_model.py style imports resolve to the B200 implementation."""
    for name in _DROPIN:
        sys.modules[name] = importlib.import_module(f"{__name__}.{name}")


for _m in _DROPIN + ["engine", "parallel", "optim"]:
    importlib.import_module(f"{__name__}.{_m}")
