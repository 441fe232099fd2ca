"""B200-native detection hot path with the reference's Python interface.

The directory name contains hyphens (it follows the reference repo's name), so import it with

    import importlib; fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")

or call ``install_dropin()`` once and use the reference's own import paths unchanged:

    from models.PoolResnet import PoolResnet
    from models import BaseModel, ModelMeta
    from losses.YoloLoss import yolo_loss
    from datasets.utils import ReduceBoundingBoxes

``install_dropin`` is an OVERLAY, not a replacement: the hot-path modules resolve to this package, every other name
of the reference tree (``datasets.WIDERFace.WIDERFaceDataModule``, ``datasets.utils.draw_bbx``,
``losses.SSDLoss.ssd_loss2``, ``models.ModelMetaSSD`` ...) falls through to the reference's own files when a
reference checkout is on ``sys.path`` (or passed as ``reference_root``), so ``train_model.py`` /
``train_model_ssd.py`` import unchanged.
"""
import importlib
import importlib.util
import os
import sys
import types

from . import native, ops  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))

# reference module name -> implemented here (hot path).  Everything else falls through to the reference tree.
_DROPIN = ["models", "models.BaseModel", "models.BaseSSDModel", "models.PoolResnet", "models.Resnet",
           "models.SeparableCNN", "models.MobilenetV3Backbone", "models.SSD", "models.ModelMeta", "losses",
           "losses.YoloLoss", "losses.SSDLoss", "datasets", "datasets.utils", "datasets.WIDERFace",
           "datasets.WIDERFace.dataset", "datasets.WIDERFace.dataset_ssd"]
_PACKAGES = ["models", "losses", "datasets", "datasets.WIDERFace"]
_reference_root = None


def find_reference_root(extra=None):
    """A directory that holds the reference checkout (models/BaseModel.py + datasets/utils.py + losses/YoloLoss.py) and
    is not this package: ``extra``, $FD_REFERENCE_ROOT, the current directory, then every ``sys.path`` entry."""
    cands = [extra, os.environ.get("FD_REFERENCE_ROOT"), os.getcwd()] + list(sys.path)
    for c in cands:
        if not c:
            continue
        c = os.path.abspath(c)
        if c == _HERE:
            continue
        if all(os.path.isfile(os.path.join(c, rel)) for rel in
               ("models/BaseModel.py", "datasets/utils.py", "losses/YoloLoss.py")):
            return c
    return None


def _reference_module(name):
    """The reference's OWN module ``name`` (e.g. ``datasets.utils``), loaded from the reference tree under the private
    name ``_fd_reference.<name>`` -- used for the names the hot-path mirror does not re-implement."""
    if _reference_root is None:
        return None
    priv = "_fd_reference." + name
    if priv in sys.modules:
        return sys.modules[priv]
    base = os.path.join(_reference_root, *name.split("."))
    path = base + ".py" if os.path.isfile(base + ".py") else os.path.join(base, "__init__.py")
    if not os.path.isfile(path):
        return None
    spec = importlib.util.spec_from_file_location(priv, path)
    mod = importlib.util.module_from_spec(spec)
    # relative imports inside a reference package __init__ (``from .datamodule import WIDERFaceDataModule``) must
    # resolve through the PUBLIC overlay names, so that hot-path submodules are ours and the rest the reference's
    mod.__package__ = name if path.endswith("__init__.py") else name.rpartition(".")[0]
    sys.modules[priv] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(priv, None)
        raise
    return mod


def _fallthrough(public_name, module):
    """Module-level ``__getattr__``: names missing from the mirror are looked up (a) as overlay submodules, (b) in the
    reference's module of the same name."""
    def __getattr__(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        if hasattr(module, "__path__"):
            try:
                return importlib.import_module(public_name + "." + attr)
            except ImportError:
                pass
        ref = _reference_module(public_name)
        if ref is not None and hasattr(ref, attr):
            return getattr(ref, attr)
        where = f"reference tree {_reference_root}" if _reference_root else "no reference checkout on sys.path"
        raise AttributeError(f"module {public_name!r} (fd_b200 drop-in) has no attribute {attr!r} ({where})")
    return __getattr__


def install_dropin(reference_root=None):
    """Alias the reference's top-level module names (models, losses, datasets) to this package's hot-path mirror and
    let everything else fall through to the reference checkout (see the module docstring).  Idempotent."""
    global _reference_root
    _reference_root = find_reference_root(reference_root)
    for name in _DROPIN:
        mod = importlib.import_module(f"{__name__}.{name}")
        sys.modules[name] = mod
        if name in _PACKAGES:
            ours = os.path.join(_HERE, *name.split("."))
            path = [ours]
            if _reference_root is not None:
                path.append(os.path.join(_reference_root, *name.split(".")))
            mod.__path__ = path          # submodules we do not implement are found in the reference directory
        mod.__getattr__ = _fallthrough(name, mod)
    # parent attributes (``import datasets.WIDERFace`` then ``datasets.WIDERFace``)
    for name in _DROPIN:
        parent, _, child = name.rpartition(".")
        if parent and not hasattr(sys.modules[parent], child):      # `models.ModelMeta` must stay the CLASS
            setattr(sys.modules[parent], child, sys.modules[name])
    return _reference_root


def uninstall_dropin():
    """Remove the aliases (tests)."""
    for name in list(sys.modules):
        if name in _DROPIN or name.startswith("_fd_reference.") or any(
                name.startswith(p + ".") for p in ("models", "losses", "datasets")):
            mod = sys.modules.get(name)
            f = getattr(mod, "__file__", "") or ""
            if name.startswith("_fd_reference.") or f.startswith(_HERE) or (
                    _reference_root and f.startswith(_reference_root)):
                sys.modules.pop(name, None)


for _m in _DROPIN + ["engine", "parallel", "optim", "export"]:
    importlib.import_module(f"{__name__}.{_m}")
