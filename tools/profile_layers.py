"""Per-call device times of a model's forward (or train step) through the ops layer: every ops.* wrapper is timed with
CUDA events (synchronising after each call; use it for SHARES, the graph-replay time is the truth).

    python tools/profile_layers.py mbv3|ssd [B]
"""
import importlib
import json
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
ops = fd.ops


def instrument(records):
    names = [n for n in dir(ops) if callable(getattr(ops, n)) and not n.startswith("_") and n not in (
        "check", "cur_stream", "dptr", "lib", "pw_packed_elems", "pw_padded_n", "dwconv_se_blocks", "stem_cache_elems",
        "resblock_chain_ok")]
    orig = {}
    for n in names:
        f = getattr(ops, n)
        if not hasattr(f, "__code__"):
            continue
        orig[n] = f

        def make(n, f):
            def timed(*a, **k):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                r = f(*a, **k)
                e.record()
                torch.cuda.synchronize()
                shp = "x".join(str(d) for d in a[0].shape) if len(a) and hasattr(a[0], "shape") else ""
                if isinstance(a[0], (list, tuple)) and len(a[0]) and hasattr(a[0][0], "shape"):
                    shp = f"{len(a[0])}p:" + "x".join(str(d) for d in a[0][0].shape)
                extra = ""
                if n == "pw_conv":
                    extra = f"->{a[3]}"
                if n == "dwconv":
                    extra = f" k{a[3]}s{a[4]}"
                records.append((n + extra, shp, s.elapsed_time(e) * 1e3))
                return r
            return timed
        setattr(ops, n, make(n, f))
    return orig


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "mbv3"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (256 if which == "mbv3" else 16)
    dev = torch.device("cuda")
    records = []
    if which == "mbv3":
        torch.manual_seed(9)
        m = fd.models.MobilenetV3Backbone.MobilenetV3Backbone(576, (3, 480, 480), 15).to(dev).eval()
        x = torch.rand(B, 3, 480, 480, device=dev)
        with torch.no_grad():
            m(x[:2]); m(x)
        instrument(records)
        with torch.no_grad():
            m.engine.forward(x)
    else:
        torch.manual_seed(2)
        m = fd.models.SSD.SSD(filters=16, input_shape=(3, 480, 480)).to(dev).train()
        x = torch.rand(B, 3, 480, 480, device=dev)
        gen = torch.Generator().manual_seed(1)
        sys.path.insert(0, ROOT)
        from bench import synth_boxes
        boxes = [synth_boxes(gen, 1, 119) for _ in range(B)]
        y = fd.datasets.WIDERFace.dataset_ssd.convert_bbx_to_feature_maps_batch(boxes, (480, 480), device=dev)
        m.train_step(x, y); m.train_step(x, y)
        instrument(records)
        m.train_step(x, y)
    agg = OrderedDict()
    for n, shp, us in records:
        k = (n, shp)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"{which} B={B}: {len(records)} calls, {tot:.0f} us (sum of per-call times)")
    for (n, shp), (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{us:9.1f} us {100 * us / tot:5.1f}%  x{c:<3d} {n:28s} {shp}")
    byop = {}
    for (n, shp), (c, us) in agg.items():
        byop[n.split(" ")[0].split("->")[0]] = byop.get(n.split(" ")[0].split("->")[0], 0) + us
    print(json.dumps({k: round(v, 1) for k, v in sorted(byop.items(), key=lambda kv: -kv[1])}))


if __name__ == "__main__":
    main()
