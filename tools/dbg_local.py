#!/usr/bin/env python3
"""Debug: dense forward on 'local attention' inputs (large logit range) vs an fp32 torch reference, error per Q tile."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import _native as N
b, s, h = 1, int(os.environ.get("S", 2600)), 2
g = torch.Generator().manual_seed(3)
t = torch.arange(s).float()
w = torch.randn(128, generator=g) * 0.02
ph = torch.rand(128, generator=g) * 2 * math.pi
e = math.sqrt(2 / 128) * torch.cos(t[:, None] * w[None] + ph[None])
base = float(os.environ.get("AMP", 16.0)) * e[None, :, None, :].expand(b, s, h, 128)
q = (base + 0.5 * torch.randn(b, s, h, 128, generator=g)).to(torch.bfloat16).cuda()
k = (base + 0.5 * torch.randn(b, s, h, 128, generator=g)).to(torch.bfloat16).cuda()
v = torch.randn(b, s, h, 128, generator=g).to(torch.bfloat16).cuda()
out = torch.empty_like(q); lse = torch.empty(b, h, s, device="cuda")
N.fwd(q, k, v, out, lse, 128 ** -0.5); torch.cuda.synchronize()
qf, kf, vf = (x.float().permute(0, 2, 1, 3) for x in (q, k, v))
sc = (qf @ kf.transpose(-1, -2)) * 128 ** -0.5
ref = (torch.softmax(sc, -1) @ vf).permute(0, 2, 1, 3)
lref = torch.logsumexp(sc, -1)
err = (out.float() - ref).abs()
print("max O err", err.max().item(), "max LSE err", (lse - lref).abs().max().item(), "nan", torch.isnan(out.float()).sum().item())
for m in range((s + 127) // 128):
    e_ = err[:, m * 128:(m + 1) * 128]
    print(f"  qtile {m:3d}: max err {e_.max().item():.4f}  rows>0.01: {(e_.amax(dim=(0, 2, 3)) > 0.01).sum().item()}  "
          f"cols half0 {e_[..., :64].max().item():.4f} half1 {e_[..., 64:].max().item():.4f}")
bad = (err.amax(dim=(0, 3)) > 0.01).nonzero()   # (row, head)
print("bad (row, head):", bad[:80].tolist())
for r, hh in bad[:12].tolist():
    o, rf = out[0, r, hh].float(), ref[0, r, hh]
    ratio = (o / rf)[rf.abs() > 0.05]
    # which key tiles carry the mass for this row
    p = torch.softmax(sc[0, hh, r], -1)
    mass = [round(p[t * 176:(t + 1) * 176].sum().item(), 3) for t in range((s + 175) // 176)]
    rowmax = [round(sc[0, hh, r, t * 176:(t + 1) * 176].max().item() * 1.4427, 1) for t in range((s + 175) // 176)]
    print(f"row {r} head {hh}: ratio med {ratio.median().item():.3f} min {ratio.min().item():.3f} max {ratio.max().item():.3f}\n   mass/tile {mass}\n   log2 rowmax/tile {rowmax}")
