import importlib

import pytest
import torch


def fd():
    return importlib.import_module("pytorch-face-detection-from-scratch_b200")


def require_cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def rel_err(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
