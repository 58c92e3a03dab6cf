import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import test_gpu_head as T
from oracle import nawsod_oracle as O
for dtype in (torch.float32, torch.bfloat16):
    for use_masks in (False, True):
        prob = T._problem(1, 32, 14, 18, 96, 6, 128, seed=3, wscale=4.0)
        m, bl = T._run(dtype, prob, use_masks=use_masks)
        ref = T._oracle(prob, image=0, dtype=dtype, use_masks=use_masks)
        print("==== dtype", dtype, "masks", use_masks)
        feat = bl["roi_feat"].float().cpu().numpy().reshape(96, 49, 32).transpose(0, 2, 1).reshape(96, -1)
        print("roi_feat", T.rel_l2(feat, ref["roi_feat"]))
        C = 5
        lg = bl["fc8_logits"].cpu().numpy()
        print("drop7", T.rel_l2(bl["drop7_cat"][:, :128].float().cpu().numpy(), ref["drop7"]))
        print("fc8c", T.rel_l2(lg[0][:, :C], ref["fc8c"]), "fc8d", T.rel_l2(lg[0][:, C:2*C], ref["fc8d"]), "nfc8c", T.rel_l2(lg[1][:, :C], ref["nfc8c"]))
        print("abs logit err max", np.abs(lg[0][:, :C]-ref["fc8c"]).max(), "logit scale", np.abs(ref["fc8c"]).max())
        for k in ("rois_pred", "rois_pred_noise"):
            print(k, T.rel_l2(bl[k].cpu().numpy(), ref[k]))
        print("cls_prob", T.rel_l2(bl["cls_prob"][0].cpu().numpy(), ref["cls_prob"][0]), "w_noise", T.rel_l2(bl["class_weight_noise"][0].cpu().numpy(), ref["class_weight_noise"][0]))
        print("loss", bl["loss_cls"][0].item(), ref["loss_cls"], bl["loss_cls_noise"][0].item(), ref["loss_cls_noise"])
        print("d_fc8c", T.rel_l2(bl["d_fc8c"].cpu().numpy(), ref["d_fc8c"]), "d_fc8d", T.rel_l2(bl["d_fc8d"].cpu().numpy(), ref["d_fc8d"]))
        g = m.export_reference_grads()
        for k, ko in (("fc6_w","fc6_w"),("fc6_b","fc6_b"),("fc7_w","fc7_w"),("fc7_b","fc7_b"),("fc8c_w","fc8c_w"),("fc8d_w","fc8d_w"),("fc8c_b","fc8c_b"),("_[noisy]_fc6_w","noisy_fc6_w"),("_[noisy]_fc7_w","noisy_fc7_w"),("noisy_fc8c_w","noisy_fc8c_w"),("noisy_fc8d_b","noisy_fc8d_b")):
            print("  grad", k, T.rel_l2(g[k].float().cpu().numpy(), ref["grads"][ko]))
