import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nafwebsod_b200 as pkg
from nafwebsod_b200 import ops
import bench
for kv in filter(None, os.environ.get("NAWSOD_TUNING", "").split(",")):      # e.g. NAWSOD_TUNING=pool_rows2=1
    k, v = kv.split("=")
    pkg.set_tuning(k, int(v))
X, rois, obn, L, offs = bench.synth_inputs(2, 2000, 0)        # BASELINE config 2: 2 images x 2000 RoIs
Xd = torch.from_numpy(X).cuda(); r = torch.from_numpy(rois).cuda(); b = torch.from_numpy(obn).cuda()
for dt, train in ((torch.float32, True), (torch.bfloat16, False)):
    Xcl = ops.to_channels_last(Xd, dt)
    for _ in range(2):
        ops.RoIPoolF(Xcl, r, boost=b, is_test=not train, x_layout="NHWC", y_layout="NHWC")
torch.cuda.synchronize()
