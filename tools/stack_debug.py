"""Stacked (3-d) FC ops vs the same ops run per stack (GPU)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nafwebsod_b200 import ops
torch.manual_seed(0)
def stacked(buf, S):
    R = buf.shape[0]
    return buf.view(R, S, buf.shape[1] // S).permute(1, 0, 2)
for dt in (torch.float32, torch.bfloat16):
    for (R, H, N) in ((96, 128, 128), (96, 128, 12), (300, 256, 40)):
        S = 2
        X = torch.randn(R, S * H, device="cuda").to(dt)
        W = (torch.randn(S, N, H, device="cuda") * 0.1).to(dt)
        b = torch.randn(S, 16 if N < 16 else N, device="cuda")[:, :N]
        Y = torch.empty(R, S * N if N % 8 == 0 else S * 16, device="cuda", dtype=dt)
        Ys = Y.view(R, S, -1).permute(1, 0, 2)[:, :, :N]
        ops.FC(stacked(X, S), W, b, relu=True, out=Ys)
        Yr = torch.stack([ops.FC(X[:, s * H:(s + 1) * H], W[s], b[s].contiguous(), relu=True) for s in range(S)])
        print(dt, (R, H, N), "fwd max diff", (Ys.float() - Yr.float()).abs().max().item())
        dY = torch.randn(S, R, 16 if N < 16 else N, device="cuda").to(dt)[:, :, :N]
        dX = torch.empty(R, S * H, device="cuda", dtype=dt)
        ops.FCGradientX(dY, W, act_below=stacked(X, S), out=stacked(dX, S))
        dXr = torch.cat([ops.FCGradientX(dY[s], W[s], act_below=X[:, s * H:(s + 1) * H]) for s in range(S)], dim=1)
        print(dt, (R, H, N), "bwd_x max diff", (dX.float() - dXr.float()).abs().max().item())
        dW = torch.zeros(S, N, H, device="cuda"); db = torch.zeros(S, 16 if N < 16 else N, device="cuda")[:, :N]
        ops.FCGradientW(dY, stacked(X, S), dW=dW, db=db)
        refs = [ops.FCGradientW(dY[s], X[:, s * H:(s + 1) * H]) for s in range(S)]
        print(dt, (R, H, N), "bwd_w max diff", max((dW[s] - refs[s][0]).abs().max().item() for s in range(S)),
              "db", max((db[s] - refs[s][1]).abs().max().item() for s in range(S)), flush=True)
