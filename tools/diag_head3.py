import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import test_gpu_head as T
from oracle import nawsod_oracle as O
prob = T._problem(1, 512, 38, 50, 2000, 21, 4096, seed=1)
m, bl = T._run(torch.bfloat16, prob)
pat = T._patterns(m, bl, slice(0, 2000), True)
ref = T._oracle(prob, image=0, dtype=torch.bfloat16, relu_patterns=pat)
H = 4096; C = 20
feat = bl["roi_feat"].float().cpu().numpy().reshape(2000, 49, 512).transpose(0, 2, 1).reshape(2000, -1)
print("roi_feat", T.rel_l2(feat, ref["roi_feat"]))
print("drop6", T.rel_l2(bl["drop6_cat"][:, :H].float().cpu().numpy(), ref["acts"]["drop6"]), "drop7", T.rel_l2(bl["drop7_cat"][:, :H].float().cpu().numpy(), ref["acts"]["drop7"]))
lg = bl["fc8_logits"].cpu().numpy()
for nm, a, b in (("fc8c", lg[0][:, :C], ref["fc8c"]), ("fc8d", lg[0][:, C:2*C], ref["fc8d"]), ("nfc8c", lg[1][:, :C], ref["nfc8c"])):
    print(nm, "rel", T.rel_l2(a, b), "abs rms", np.sqrt(((a-b)**2).mean()), "std", b.std(), "max", np.abs(b).max())
# what if fc8 were computed exactly from the GPU's bf16 drop7?
d7 = bl["drop7_cat"][:, :H].float().cpu().numpy()
ex = d7 @ prob[4]["fc8c_w"].T
print("fc8c from gpu drop7 with fp32 W (exact fc8): abs rms vs ref", np.sqrt(((ex-ref["fc8c"])**2).mean()))
for k in ("rois_pred", "rois_pred_noise"):
    print(k, T.rel_l2(bl[k].cpu().numpy(), ref[k]))
print("cls_prob", T.rel_l2(bl["cls_prob"][0].cpu().numpy(), ref["cls_prob"][0]), "cls_prob_noise", T.rel_l2(bl["cls_prob_noise"][0].cpu().numpy(), ref["cls_prob_noise"][0]))
print("w_noise", T.rel_l2(bl["class_weight_noise"][0].cpu().numpy(), ref["class_weight_noise"][0]))
print("loss", bl["loss_cls"][0].item(), ref["loss_cls"], bl["loss_cls_noise"][0].item(), ref["loss_cls_noise"])
for k in ("d_fc8c", "d_fc8d", "d_nfc8c", "d_nfc8d"):
    print(k, T.rel_l2(bl[k].cpu().numpy(), ref[k]))
g = m.export_reference_grads()
for k, ko in (("fc6_w","fc6_w"),("fc6_b","fc6_b"),("fc7_w","fc7_w"),("fc8c_w","fc8c_w"),("fc8d_w","fc8d_w"),("_[noisy]_fc6_w","noisy_fc6_w"),("_[noisy]_fc7_w","noisy_fc7_w"),("noisy_fc8c_w","noisy_fc8c_w")):
    print("  grad", k, T.rel_l2(g[k].float().cpu().numpy(), ref["grads"][ko]))
