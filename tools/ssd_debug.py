"""SSD engine vs oracle: per-tensor gradient errors and intermediate gradients (debug)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import backbone_oracle as bo
from bench import synth_boxes
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200")
torch.manual_seed(7)
m = fd.models.SSD.SSD(filters=16, input_shape=(3, 480, 480)).cuda().eval()
p = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
B = 2
gen = torch.Generator().manual_seed(8)
x = torch.rand(B, 3, 480, 480, generator=gen)
boxes = [synth_boxes(gen, 5, 80) for _ in range(B)]
y = fd.datasets.WIDERFace.dataset_ssd.convert_bbx_to_feature_maps_batch(boxes, (480, 480))
loss = m.train_step(x.cuda(), y)
y_ref, loss_ref, g_ref = bo.ssd_train_step(x, y.cpu(), p, 10)
print("loss", loss.item(), loss_ref.item())
for k, prm in m.named_parameters():
    g, r = prm.grad.cpu().double(), g_ref[k].double()
    e = ((g - r).norm() / (r.norm() + 1e-30)).item()
    print(f"{k:55s} rel {e:9.4f}  |g| {g.norm().item():10.4g} |ref| {r.norm().item():10.4g}")
eng = m.engine
pl = eng.plan(B, True)
# gradient w.r.t. the stem output from the oracle
leaves = {k: v.clone().requires_grad_(True) for k, v in p.items()}
import torch.nn.functional as F
s = F.conv2d(x, leaves["input_normalizer.weight"], leaves["input_normalizer.bias"], stride=2, padding=1)
s.retain_grad()
# re-run oracle forward from s
def fwd_from(s):
    yv = s; k = 0
    for b in range(9):
        yv = bo.ssd_block(yv, leaves, f"feature_extractor.{b}.", pool=b < 2)
    sc, bx = [], []
    for i in range(4):
        yv = bo.ssd_block(yv, leaves, f"continue_layers.{i}.0.", pool=i != 0)
        z = F.linear(yv.permute(0, 2, 3, 1).contiguous(), leaves[f"extracting_layers.{i}.0.weight"], leaves[f"extracting_layers.{i}.0.bias"]).reshape(B, -1, 5)
        sc.append(z[..., :1]); bx.append(z[..., 1:5])
    out = torch.cat([torch.sigmoid(torch.cat(sc, 1)), torch.cat(bx, 1)], 2)
    mult, pri = bo.ssd_priors_torch()
    o = out.clone()
    o[..., 1:2] = o[..., 1:2] * mult; o[..., 2:3] = o[..., 2:3] * mult; o[..., 1:5] = o[..., 1:5] + pri
    return o
yh = fwd_from(s)
l = bo.ssd_loss_torch(yh[:, :, 0], yh[:, :, 1:], y.cpu()[:, :, 0], y.cpu()[:, :, 1:], 10)
l.backward()
gs_ref = s.grad.permute(0, 2, 3, 1)          # [B,240,240,16]
gs = pl["g_stem"][0].float().cpu()[..., :16]
print("g_stem rel", ((gs - gs_ref).norm() / gs_ref.norm()).item(), gs.norm().item(), gs_ref.norm().item())
print("g_stem padded channels max", pl["g_stem"][0].float()[..., 16:].abs().max().item())
gsw = eng.gpad[eng.pad_stem_w_off:eng.pad_stem_b_off].view(64, 3, 3, 3)
print("stem dw padded-buffer norms: real", gsw[:16].norm().item(), "pad", gsw[16:].norm().item(), "ref", g_ref["input_normalizer.weight"].norm().item())
