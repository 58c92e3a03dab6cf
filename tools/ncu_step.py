"""A short instrumented run for ncu: a few head steps at BASELINE config 2 (no timing claims)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from nafwebsod_b200.heads import WeblyHeadModel
from nafwebsod_b200.dp import DataParallelHead
dev = torch.device("cuda", 0)
m = WeblyHeadModel(21, 512, 7, 4096, dtype=torch.bfloat16)
g = torch.Generator(device=dev).manual_seed(2)
m.flat_param[:m.n_weights].normal_(0.0, 0.01, generator=g)
m.sync_shadow(); m.UpdateWorkspaceLr(1e-3)
X, rois, obn, L, offs = bench.synth_inputs(2, 2000, 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
m.FeedBlobs(t(X), t(rois), t(obn), t(L), t(offs), x_layout="NCHW")
dp = DataParallelHead(m)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for i in range(steps):
    dp.step(dropout_seed=i + 1)
torch.cuda.synchronize()
print("done", m.blobs["loss"].cpu().tolist())
