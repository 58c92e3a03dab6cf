import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import test_gpu_head as T
from oracle import nawsod_oracle as O
for dtype in (torch.float32, torch.bfloat16):
    prob = T._problem(1, 32, 14, 18, 96, 6, 128, seed=3, wscale=4.0)
    X, rois, obn, L, params, masks, offs = prob
    m, bl = T._run(dtype, prob, use_masks=True)
    Xo = torch.from_numpy(X).to(torch.bfloat16).float().numpy() if dtype == torch.bfloat16 else X
    mk = {k: v.astype(np.float32) for k, v in masks.items()}
    Y, A = O.roi_pool_f(Xo, rois, 1/16)
    feat = O.roi_feature_boost(Y, obn).reshape(96, -1)
    H = 128
    for s, pre in ((0, ""), (1, "noisy_")):
        acts = O.fc_stack_forward(feat, params[pre+"fc6_w"], params[pre+"fc6_b"], params[pre+"fc7_w"], params[pre+"fc7_b"], mk[pre+"drop6"], mk[pre+"drop7"])
        pre6 = O.fc(feat, params[pre+"fc6_w"], params[pre+"fc6_b"])
        d6g = bl["drop6_cat"][:, s*H:(s+1)*H].float().cpu().numpy()
        d7g = bl["drop7_cat"][:, s*H:(s+1)*H].float().cpu().numpy()
        print(dtype, "stack", s, "drop6 rel", T.rel_l2(d6g, acts["drop6"]), "drop7 rel", T.rel_l2(d7g, acts["drop7"]))
        mm6 = ((d6g > 0) != (acts["drop6"] > 0)); mm7 = ((d7g > 0) != (acts["drop7"] > 0))
        print("   relu-mask mismatches drop6:", mm6.sum(), "of", mm6.size, " drop7:", mm7.sum(), " |pre6| at mismatches:", np.abs(pre6[mm6])[:5], "pre6 std", pre6.std())
        # oracle's d_fc6 using ITS OWN logits grads
    ref = T._oracle(prob, image=0, dtype=dtype, use_masks=True)
    # recompute oracle d_fc6 for both stacks
    C = 5
    for s, pre in ((0, ""), (1, "noisy_")):
        acts = O.fc_stack_forward(feat, params[pre+"fc6_w"], params[pre+"fc6_b"], params[pre+"fc7_w"], params[pre+"fc7_b"], mk[pre+"drop6"], mk[pre+"drop7"])
        dc = ref["d_fc8c"] if s == 0 else ref["d_nfc8c"]; dd = ref["d_fc8d"] if s == 0 else ref["d_nfc8d"]
        _, _, dxc = O.fc_grad(acts["drop7"], params[pre+"fc8c_w"], dc); _, _, dxd = O.fc_grad(acts["drop7"], params[pre+"fc8d_w"], dd)
        d_drop7 = dxc + dxd
        d_fc7 = O.relu_grad(acts["fc7"], O.dropout_grad(d_drop7, mk[pre+"drop7"]))
        _, _, d_drop6 = O.fc_grad(acts["drop6"], params[pre+"fc7_w"], d_fc7)
        d_fc6 = O.relu_grad(acts["fc6"], O.dropout_grad(d_drop6, mk[pre+"drop6"]))
        g6 = m._buf["d_fc6"][:, s*H:(s+1)*H].float().cpu().numpy()
        print(dtype, "stack", s, "d_fc6 rel", T.rel_l2(g6, d_fc6), " nonzero-pattern mismatches", ((g6 != 0) != (d_fc6 != 0)).sum(), "max|d_fc6|", np.abs(d_fc6).max())
        if s == 1:
            g7 = m._buf["d_fc7"].float().cpu().numpy()
            print("   d_fc7(noisy, last written) rel", T.rel_l2(g7, d_fc7))
        bad = np.argwhere((g6 != 0) != (d_fc6 != 0))[:5]
        for r, j in bad:
            print("      mismatch at", r, j, "gpu", g6[r, j], "ref", d_fc6[r, j], "fc6 pre", O.fc(feat, params[pre+"fc6_w"], params[pre+"fc6_b"])[r, j], "mask", mk[pre+"drop6"][r, j])
