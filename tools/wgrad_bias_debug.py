import importlib, os, sys, torch
sys.path.insert(0, "/root/repo")
fd = importlib.import_module("pytorch-face-detection-from-scratch_b200"); ops = fd.ops
torch.manual_seed(0)
for (nprob, B, H) in ((2, 64, 60), (2, 16, 120), (1, 4, 60), (2, 64, 30), (16, 64, 15)):
    x = torch.randn(nprob, B, H, H, 64, device="cuda").bfloat16()
    g = (torch.randn(nprob, B, H, H, 64, device="cuda") * 0.1).bfloat16()
    dw = torch.zeros(nprob, 9 * 64 * 64, device="cuda"); db = torch.zeros(nprob, 64, device="cuda")
    ops.conv3x3_wgrad_multi(x, g, dw.view(-1), 9 * 64 * 64, db.view(-1), 64)
    torch.cuda.synchronize()
    ref = g.float().sum(dim=(1, 2, 3))
    print(nprob, B, H, "dbias nan:", torch.isnan(db).sum().item(), "max err", (db - ref).abs().max().item(), "dw nan:", torch.isnan(dw).sum().item(), "ref max", ref.abs().max().item())
    if torch.isnan(db).any():
        print("  nan problems/channels:", torch.isnan(db).nonzero()[:8].tolist())
