"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the module mirrors keep the reference's interface, and there is NO CPU fallback."""
import importlib
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")


def test_library_exports_every_header_symbol():
    import __graft_entry__ as ge
    ge.build()
    hdr = open(os.path.join(ROOT, "include", "fd_b200.h")).read()
    declared = set(re.findall(r"FD_API\s+[\w\s\*]+?\b(fd_\w+)\s*\(", hdr))
    assert len(declared) >= 16
    lib = fd.native.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in fd_b200.h but not exported"
    assert declared == set(fd.native.SIGNATURES), "ctypes table out of sync with the header"
    assert lib.fd_version() >= 100
    assert lib.fd_error_string(-2).decode() == "unsupported shape"


def test_sass_uses_tcgen05_and_tma():
    """The conv kernels must be real Blackwell kernels: UTCHMMA (tcgen05.mma), UTMALDG (TMA), LDTM."""
    import shutil, subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", fd.native.LIB_PATH], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnem in sass, mnem
    # the wide kernels are CTA-pair kernels: tcgen05.mma.cta_group::2, multicast commit, cta_group::2 TMA loads
    for mnem in ("UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST", "UTMALDG.4D.2CTA"):
        assert mnem in sass, mnem


def test_reference_interface_and_state_dict_keys():
    fd.install_dropin()
    from models.PoolResnet import PoolResnet          # reference import paths
    from models.Resnet import Resnet
    from models import BaseModel, ModelMeta
    from losses.YoloLoss import yolo_loss
    from datasets.utils import ReduceBoundingBoxes
    m = PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=10)
    assert isinstance(m, BaseModel) and isinstance(m, torch.nn.Module)
    assert sum(p.numel() for p in m.parameters()) == 769349
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "official_medium.npz"))
    ref_keys = [k[3:] for k in g.files if k.startswith("sd.")]
    assert list(m.state_dict().keys()) == ref_keys
    m.load_state_dict({k: torch.from_numpy(g["sd." + k]) for k in ref_keys}, strict=True)
    assert m.engine.pools == [True, True] + [False] * 8 and (m.engine.So_h, m.engine.So_w) == (10, 10)
    r = Resnet(filters=64, input_shape=(3, 480, 480), num_of_patches=15)
    assert sum(p.numel() for p in r.parameters()) == 743237
    assert r.engine.pools == [True] * 4 + [False] * 6 and r.engine.So_h == 15
    assert isinstance(m.reduce_bounding_boxes, ReduceBoundingBoxes)
    assert callable(yolo_loss) and ModelMeta(m).model is m
    with pytest.raises(AssertionError):
        PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=7)   # BaseModel.py:23-26


def test_no_cpu_fallback():
    from importlib import import_module
    PoolResnet = import_module("pytorch-face-detection-from-scratch_b200.models.PoolResnet").PoolResnet
    m = PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=10)
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 3, 480, 480))
    L = import_module("pytorch-face-detection-from-scratch_b200.losses.YoloLoss")
    with pytest.raises(RuntimeError):
        L.yolo_loss(torch.rand(5, 10, 10), torch.rand(5, 10, 10))
    RB = import_module("pytorch-face-detection-from-scratch_b200.datasets.utils").ReduceBoundingBoxes
    with pytest.raises(RuntimeError):
        RB(0.5, 0.5, (3, 480, 480), 10)(torch.rand(5, 10, 10))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pytorch-face-detection-from-scratch_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                for line in open(os.path.join(dp, f)):
                    if re.match(r"\s*(import|from)\s", line):
                        assert "oracle" not in line, (os.path.join(dp, f), line)


def test_flat_parameter_layout():
    eng = fd.engine.BackboneEngine(64, 3, 480, 480, 10, 10, 8, 2, 6, 0, lambda h: h > 20)
    total = sum(n for _, n, _ in eng.offsets.values())
    assert total == 769349
    for name, (off, n, shape) in eng.offsets.items():
        assert off % 4 == 0
    assert len(eng.param_names()) == 44


def test_div255_fma_sequence_is_the_ieee_quotient():
    """csrc/stem_tc.cu Px4<uint8_t>::div255 replaces `x / 255.0f` (models/PoolResnet.py:95) by q = x * (1/255),
    q' = fma(fma(-q, 255, x), 1/255, q).  Emulated here in exact arithmetic (fp64 holds the fp32 products exactly):
    bit-identical to the fp32 division for all 256 uint8 inputs, while the plain reciprocal multiply is not."""
    import numpy as np
    f = np.float32
    v = np.arange(256, dtype=np.float32)
    want = (v / f(255.0)).astype(np.float32)
    r = f(1.0) / f(255.0)
    q = (v * r).astype(np.float32)

    def fma(a, b, c):
        return np.float32(np.float64(a) * np.float64(b) + np.float64(c))

    e = np.array([fma(-q[i], f(255.0), v[i]) for i in range(256)], dtype=np.float32)
    q2 = np.array([fma(e[i], r, q[i]) for i in range(256)], dtype=np.float32)
    assert (q2 == want).all()
    assert (q != want).any()


def test_torchscript_export_roundtrip_without_gpu(tmp_path):
    """torch.jit.script / to_torchscript of the mirror models produce an archive whose forward calls the
    registered fd_b200 operator; it loads in a fresh module namespace, keeps the reference's forward(x, predict) signature
    and refuses to run without CUDA (no CPU path)."""
    PoolResnet = fd.models.PoolResnet.PoolResnet
    torch.manual_seed(0)
    m = PoolResnet(filters=64, input_shape=(3, 480, 480), num_of_patches=10, probability_threshold=0.7, iou_threshold=0.01)
    path = tmp_path / "poolresnet.pth"
    ts = fd.models.ModelMeta(model=m).to_torchscript(path)          # train_model.py:61
    assert "fd_b200.detector_forward" in ts.code
    back = torch.jit.load(str(path))
    assert back.family == "PoolResnet" and list(back.cfg) == [64, 3, 480, 480, 10, 10]
    assert abs(back.p_thr - 0.7) < 1e-12 and back.flat_state.numel() == 769349
    assert torch.equal(back.flat_state, fd.export.flatten_state(m))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            back(torch.zeros(2, 3, 480, 480, dtype=torch.uint8), predict=torch.tensor(1))
    for cls, args in ((fd.models.SSD.SSD, (16, (3, 480, 480))), (fd.models.MobilenetV3Backbone.MobilenetV3Backbone, (576, (3, 480, 480), 15))):
        s2 = cls(*args).to_torchscript()
        assert "fd_b200.detector_forward" in s2.code


def test_new_entry_points_reject_bad_arguments_before_touching_the_device():
    """fd_conv3x3_wide_chain / fd_stem_fwd_cached / fd_stem_wgrad_pair: null operands are FD_EINVAL (-1) -- checked before
    any CUDA call, so this runs without a GPU; the ctypes record of a chain layer has the C layout (4 ints + 15 pointers)."""
    import ctypes
    lib = fd.native.lib()
    assert ctypes.sizeof(fd.native.WideChainLayer) == 4 * 4 + 15 * 8
    assert lib.fd_conv3x3_wide_chain(None, 1, None, 1, 1, 15, 15, 0.2, None, 1, None) == -1
    assert lib.fd_stem_fwd_cached(None, None, None, 1, 3, 480, 480, 10, 8, 2, None, None) == -1
    assert lib.fd_stem_wgrad_pair(None, None, None, 1, 3, 480, 480, 10, 8, 2, None, None, None) == -1
    assert lib.fd_error_string(-1).decode() != ""
    L = fd.ops.WideChainLayer()
    L.in_index, L.w_index = 3, 5
    assert (L.in_index, L.w_index, L.flags) == (3, 5, 0) and L.out[0] is None
