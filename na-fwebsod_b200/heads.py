"""Host-side mirror of the reference's WSL / webly head builders for the per-proposal path.

The reference builds a Caffe2 graph with ``add_VGG16_roi_2fc_noise_head`` ->
``add_webly_outputs`` -> ``add_webly_losses`` (detectron/modeling/webly_heads.py:463-502, 32-74,
123-216; single-stack WSDDN: detectron/modeling/wsl_heads.py:654-681, 23-56) and runs it with
``workspace.RunNet``.  Here the same-named functions execute eagerly on a :class:`WeblyHeadModel`
(the ``model`` argument of the reference builders), each one a short run of libnawsod C-ABI
calls; blobs keep the reference's names (``rois``, ``obn_scores``, ``labels_oh``, ``roi_feat``,
``drop7``, ``fc8c`` ... ``cls_prob``, ``loss_cls``, ``loss_cls_noise``) and parameters are
imported / exported under the reference's names and layouts (``fc6_w`` [4096, 25088] with
K-index c*49+ph*7+pw, ``_[noisy]_fc6_w``, ``noisy_fc8c_w`` ...).

Device-side layout:
  * conv5 map channels-last, pooled features [R, 49*C] in (ph, pw, c) order -> fc6 weights are
    stored K-permuted once at load time;
  * the two stacks' fc6 are ONE GEMM (W6 = [fc6_w; _[noisy]_fc6_w], N = 8192): the pooled
    features are read once; fc8c|fc8d of a stack are one GEMM with N = 2C;
  * all parameters / gradients / momenta live in three flat float32 buffers (weights first,
    biases last) so the data-parallel all-reduce is a handful of large NCCL calls and the SGD
    update is two fused kernel launches; bf16 runs keep a flat bf16 shadow of the parameters.
PyTorch is used for memory, streams and (in dp.py) NCCL plumbing only.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops

_ALIGN = 64  # elements; keeps every carved view 256-byte aligned


def _round_up(v, a):
    return (v + a - 1) // a * a


def lr_change_correction(cur_lr, new_lr, scale_momentum=True, threshold=1.1):
    """Factor ``_SetNewLr`` scales every ``<param>_momentum`` blob by when the learning rate goes from ``cur_lr`` to
    ``new_lr`` (detectron/modeling/detector.py:528-560, 581-586; both float32 like the reference's blob and schedule):
    ``new_lr / cur_lr`` if SOLVER.SCALE_MOMENTUM, ``cur_lr > 1e-7`` and the change ratio exceeds the threshold, else 1."""
    cur_lr, new_lr = np.float32(cur_lr), np.float32(new_lr)
    if cur_lr == new_lr:
        return np.float32(1.0)
    eps = 1e-10
    ratio = max(new_lr / max(cur_lr, eps), cur_lr / max(new_lr, eps))
    if scale_momentum and cur_lr > 1e-7 and ratio > threshold:
        return np.float32(new_lr / cur_lr)
    return np.float32(1.0)


class WeblyHeadModel:
    """State + eager execution of the NA-fWebSOD head (``noise=True``) or plain WSDDN head.

    num_classes follows the reference: it INCLUDES background (cfg.MODEL.NUM_CLASSES), the
    heads predict ``num_classes - 1`` classes (detectron/modeling/wsl_heads.py:33).
    """

    def __init__(self, num_classes=21, dim_in=512, roi_size=7, hidden_dim=4096, spatial_scale=1.0 / 16,
                 noise=True, entropy=True, mean_loss=True, dtype=torch.bfloat16, device="cuda", train=True,
                 freeze_conv_body=True, precision=None):
        """``dtype`` is the storage type of the GEMM operands.  ``precision`` (float32 storage only): ``"tf32"`` (default) = one
        tensor-core pass per product on operands rounded to the nearest TF32 (~1e-3 end to end); ``"fp32"`` = the reference's
        precision (Caffe2 FC = sgemm, detectron/modeling/wsl_heads.py:674-679): every operand is kept as a TF32 (high, low)
        pair and every product is summed from three passes (ops.FC ``X_lo`` / ``W_lo``), ~1e-5 end to end at a third of the
        TF32 path's GEMM throughput -- the parity path, single-GPU schedules."""
        if dtype not in (torch.bfloat16, torch.float32):
            raise RuntimeError("dtype must be bfloat16 or float32 (TF32 tensor path)")
        if precision is None:
            precision = "bf16" if dtype == torch.bfloat16 else "tf32"
        if precision not in ("bf16", "tf32", "fp32") or (precision == "bf16") != (dtype == torch.bfloat16):
            raise RuntimeError("precision must be 'tf32' or 'fp32' with float32 storage, 'bf16' with bfloat16")
        self.precision = precision
        self.num_classes = num_classes
        self.C = num_classes - 1
        self.dim_in, self.roi_size, self.H = dim_in, roi_size, hidden_dim
        self.D = dim_in * roi_size * roi_size
        self.spatial_scale = spatial_scale
        self.noise, self.entropy, self.mean_loss = noise, entropy, mean_loss
        self.S = 2 if noise else 1
        self.dtype, self.device, self.train = dtype, torch.device(device), train
        self.freeze_conv_body = freeze_conv_body
        self.tf32 = dtype == torch.float32
        self.x3 = precision == "fp32"       # split operands, three passes per product
        self.blobs = {}
        self._buf = {}
        self.profile = None        # dict name -> [(start_event, end_event)] when bench.py instruments a run
        self.iter_count = 0
        self._lr_host = np.float32(0.0)     # value of the `lr` blob (UpdateWorkspaceLr keeps it; starts at 0 like the reference's)
        # callables that join work still in flight on side streams (dp.DataParallelHead registers its flush(): the
        # previous step's exchange / SGD pipeline reads `lr` and read-modify-writes the momenta and parameters there)
        self.pre_mutation_hooks = []
        # set by dp.DataParallelHead for one step at a time: the ops.FC ``gate`` of the stacked fc6 forward (its weight rows
        # arrive panel by panel from the previous step's exchange), or None
        self.fc6_gate = None
        self._alloc_params()

    # ------------------------------------------------------------------ parameters
    def _alloc_params(self):
        S, H, D, C2 = self.S, self.H, self.D, 2 * self.C
        C2p = _round_up(C2, 8)     # per-stack pitch of the fc8 biases (keeps every stack 16-byte aligned)
        # the stacks' fc7 / fc8 parameters are [S, ., .] blocks: one stacked GEMM launch serves all stacks
        shapes_w = [("W6", (S * H, D)), ("W7", (S, H, H)), ("W8", (S, C2, H))]
        shapes_b = [("b6", (S * H,)), ("b7", (S, H)), ("b8", (S, C2p))]
        off, self._slices = 0, {}
        for name, shp in shapes_w:
            n = int(np.prod(shp))
            self._slices[name] = (off, n, shp)
            off = _round_up(off + n, _ALIGN)
        self.n_weights = off
        for name, shp in shapes_b:
            n = int(np.prod(shp))
            self._slices[name] = (off, n, shp)
            off = _round_up(off + n, _ALIGN)
        self.n_total = off
        z = lambda dt: torch.zeros(self.n_total, dtype=dt, device=self.device)
        self.flat_param, self.flat_grad, self.flat_mom = z(torch.float32), z(torch.float32), z(torch.float32)
        # GEMM-operand copy of the parameters: bf16, or float rounded to the nearest TF32
        self.flat_lp = z(self.dtype)
        self.flat_lo = z(torch.float32) if self.x3 else None      # fp32 path: low parts of the parameters (flat_lp = high parts)
        self.lr = torch.zeros(1, dtype=torch.float32, device=self.device)   # the reference's `lr` blob
        view = lambda flat, k: flat[self._slices[k][0]: self._slices[k][0] + self._slices[k][1]].view(self._slices[k][2])
        def views(flat):
            d = {k: view(flat, k) for k in self._slices}
            d["b8"] = d["b8"][:, :C2]
            for s in range(S):                                             # per-stack aliases
                d["W7_%d" % s], d["W8_%d" % s] = d["W7"][s], d["W8"][s]
                d["b7_%d" % s], d["b8_%d" % s] = d["b7"][s], d["b8"][s]
            return d
        self.p = views(self.flat_param)       # fp32 masters
        self.g = views(self.flat_grad)
        self.w = views(self.flat_lp)          # GEMM operands (bf16 shadow or the master)
        self.wl = views(self.flat_lo) if self.x3 else {}

    def _join_side_streams(self):
        """Before the host touches `lr`, the momenta or the parameters on the compute stream (or reads them back):
        wait for every update that a data-parallel step still has in flight on its exchange stream."""
        for hook in self.pre_mutation_hooks:
            hook()

    def _ref_names(self, s):
        """Reference blob-name prefixes of stack s (detectron/modeling/webly_heads.py:490-498, 36-55)."""
        return ("", "") if s == 0 else ("_[noisy]_", "noisy_")

    def _k_permute(self, W_ref):
        """[O, c*49+bin] (reference, NCHW flatten) -> [O, bin*Cin+c] (pooled-NHWC order)."""
        O = W_ref.shape[0]
        bins = self.roi_size * self.roi_size
        return W_ref.reshape(O, self.dim_in, bins).permute(0, 2, 1).reshape(O, self.D)

    def _k_unpermute(self, W_fast):
        O = W_fast.shape[0]
        bins = self.roi_size * self.roi_size
        return W_fast.reshape(O, bins, self.dim_in).permute(0, 2, 1).reshape(O, self.D)

    def load_reference_params(self, params):
        """params: dict of reference-named arrays (numpy or torch).  Stack-2 names may be given
        as the reference writes them (``_[noisy]_fc6_w``, ``noisy_fc8c_w``) or with a plain
        ``noisy_`` prefix (the oracle's convention)."""
        def get(*names):
            for n in names:
                if n in params:
                    v = params[n]
                    return (torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v).to(
                        self.device, torch.float32)
            raise KeyError(names[0])
        H, C = self.H, self.C
        for s in range(self.S):
            a, b = self._ref_names(s)
            alt = "noisy_" if s else ""
            self.p["W6"][s * H:(s + 1) * H].copy_(self._k_permute(get(a + "fc6_w", alt + "fc6_w")))
            self.p["b6"][s * H:(s + 1) * H].copy_(get(a + "fc6_b", alt + "fc6_b"))
            self.p["W7_%d" % s].copy_(get(a + "fc7_w", alt + "fc7_w"))
            self.p["b7_%d" % s].copy_(get(a + "fc7_b", alt + "fc7_b"))
            self.p["W8_%d" % s][:C].copy_(get(b + "fc8c_w"))
            self.p["W8_%d" % s][C:].copy_(get(b + "fc8d_w"))
            self.p["b8_%d" % s][:C].copy_(get(b + "fc8c_b"))
            self.p["b8_%d" % s][C:].copy_(get(b + "fc8d_b"))
        self.sync_shadow()

    def _param_targets(self):
        """Reference blob name -> (parameter view, momentum view, transform applied to a loaded array)."""
        H, C, out = self.H, self.C, {}
        mom = {k: self.flat_mom[self._slices[k][0]: self._slices[k][0] + self._slices[k][1]].view(self._slices[k][2])
               for k in self._slices}
        mom["b8"] = mom["b8"][:, :2 * C]
        ident = lambda v: v
        for s in range(self.S):
            a, b = self._ref_names(s)
            rows = slice(s * H, (s + 1) * H)
            out[a + "fc6_w"] = (self.p["W6"][rows], mom["W6"][rows], self._k_permute)
            out[a + "fc6_b"] = (self.p["b6"][rows], mom["b6"][rows], ident)
            out[a + "fc7_w"] = (self.p["W7"][s], mom["W7"][s], ident)
            out[a + "fc7_b"] = (self.p["b7"][s], mom["b7"][s], ident)
            out[b + "fc8c_w"] = (self.p["W8"][s][:C], mom["W8"][s][:C], ident)
            out[b + "fc8d_w"] = (self.p["W8"][s][C:], mom["W8"][s][C:], ident)
            out[b + "fc8c_b"] = (self.p["b8"][s][:C], mom["b8"][s][:C], ident)
            out[b + "fc8d_b"] = (self.p["b8"][s][C:], mom["b8"][s][C:], ident)
        return out

    def initialize_from_weights(self, src_blobs):
        """``initialize_gpu_from_weights_file`` (detectron/utils/net_wsl.py:53-137) for the head's parameters, given the
        unpickled blob dictionary (the ``'blobs'`` sub-dictionary when present, :67-70).  Per parameter blob, in model
        order: a blob named ``_[xyz]_foo`` that is NOT in the file is initialised from ``foo`` (:79-88 -- how the noisy
        fc6 / fc7 start from the ImageNet VGG16 weights of the clean stack); a source blob that is missing altogether
        leaves the current initialisation (:89-91); ``<name>_momentum`` is loaded along when present (:93, 115-119); a
        shape mismatch is an error (:105-111).  Returns the list of (destination, source, with_momentum) loaded."""
        if "blobs" in src_blobs:
            src_blobs = src_blobs["blobs"]
        self._join_side_streams()
        loaded = []
        for name, (pv, mv, xf) in self._param_targets().items():
            src_name = name[name.find("]_") + 2:] if (name.find("]_") >= 0 and name not in src_blobs) else name
            if src_name not in src_blobs:
                continue
            def as_tensor(v):
                v = torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v
                return v.to(self.device, torch.float32)
            ref_shape = tuple(self._k_unpermute(pv).shape) if xf == self._k_permute else tuple(pv.shape)
            w = as_tensor(src_blobs[src_name])
            if tuple(w.shape) != ref_shape:
                raise RuntimeError("blob %s with shape %s does not match weights file shape %s of %s" % (
                    name, ref_shape, tuple(w.shape), src_name))
            pv.copy_(xf(w))
            has_momentum = src_name + "_momentum" in src_blobs
            if has_momentum:
                mv.copy_(xf(as_tensor(src_blobs[src_name + "_momentum"])))
            loaded.append((name, src_name, has_momentum))
        if self.flat_param.is_cuda:
            self.sync_shadow()
        return loaded

    def weights_file_blobs(self):
        """The ``blobs`` dictionary ``save_model_to_weights_file`` pickles for the head (detectron/utils/net_wsl.py:140-181):
        every parameter under its unscoped blob name and ``<name>_momentum`` for every trainable parameter, in the
        reference's layouts (fc6 K-order ``c*49 + bin``).  Feeding the result to ``initialize_from_weights`` restores the
        state exactly.  (In data-parallel runs call ``dp.gather_master_state()`` first.)"""
        self._join_side_streams()
        blobs = {}
        for name, (pv, mv, xf) in self._param_targets().items():
            back = self._k_unpermute if xf == self._k_permute else (lambda v: v)
            blobs[name] = back(pv).detach().cpu().numpy().copy()
            blobs[name + "_momentum"] = back(mv).detach().cpu().numpy().copy()
        return blobs

    def sync_shadow(self):
        if self.x3:
            ops.split_tf32(self.flat_param.view(1, -1), hi=self.flat_lp.view(1, -1), lo=self.flat_lo.view(1, -1))
        elif self.dtype == torch.bfloat16:
            ops.to_bf16(self.flat_param.view(1, -1), out=self.flat_lp.view(1, -1))
        else:
            ops.round_to_tf32(self.flat_param.view(1, -1), out=self.flat_lp.view(1, -1))

    def _export(self, src):
        out, H, C = {}, self.H, self.C
        for s in range(self.S):
            a, b = self._ref_names(s)
            out[a + "fc6_w"] = self._k_unpermute(src["W6"][s * H:(s + 1) * H])
            out[a + "fc6_b"] = src["b6"][s * H:(s + 1) * H]
            out[a + "fc7_w"], out[a + "fc7_b"] = src["W7_%d" % s], src["b7_%d" % s]
            out[b + "fc8c_w"], out[b + "fc8d_w"] = src["W8_%d" % s][:C], src["W8_%d" % s][C:]
            out[b + "fc8c_b"], out[b + "fc8d_b"] = src["b8_%d" % s][:C], src["b8_%d" % s][C:]
        return out

    def export_reference_params(self):
        """Parameters under the reference's blob names and layouts (checkpoint contract,
        detectron/utils/net_wsl.py:140-181)."""
        self._join_side_streams()
        return self._export(self.p)

    def export_reference_grads(self):
        self._join_side_streams()
        return self._export(self.g)

    def _timed(self, name, fn):
        """Run fn(); when instrumented, bracket it with CUDA events on the launching stream."""
        if self.profile is None:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        self.profile.setdefault(name, []).append((a, b))
        return r

    # ------------------------------------------------------------------ scratch
    def _scratch(self, name, shape, dtype):
        t = self._buf.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._buf[name] = t
        return t

    def _lo(self, name, buf):
        """fp32 path: split ``buf`` [rows, cols] in place into its TF32 high part and a low part kept in scratch; None otherwise."""
        if not self.x3:
            return None
        lo = self._scratch(name + "_lo", tuple(buf.shape), torch.float32)
        ops.split_tf32(buf, hi=buf, lo=lo)
        return lo

    # ------------------------------------------------------------------ inputs
    def FeedBlobs(self, data_conv5, rois, obn_scores, labels_oh=None, roi_offsets=None, x_layout="NHWC"):
        """Feed the head's input blobs (the contract of detectron/roi_data/wsl.py:20-58 downstream of
        the frozen conv body): conv5 map, ``rois`` [R,5], ``obn_scores`` [R,1] (already +1,
        roi_data/wsl.py:101-103), ``labels_oh`` [B,C].  ``roi_offsets`` [B+1] int32 marks the
        contiguous per-image row ranges (one image when omitted)."""
        if x_layout == "NCHW":
            data_conv5 = ops.to_channels_last(data_conv5, self.dtype)
        elif data_conv5.dtype != self.dtype:
            raise RuntimeError("channels-last conv5 map must already be %s" % self.dtype)
        R = rois.shape[0]
        if roi_offsets is None:
            roi_offsets = torch.tensor([0, R], dtype=torch.int32, device=self.device)
        self.blobs.update(conv5=data_conv5, rois=rois, obn_scores=obn_scores.reshape(-1), roi_offsets=roi_offsets)
        if labels_oh is not None:
            self.blobs["labels_oh"] = labels_oh

    # ------------------------------------------------------------------ forward pieces
    def _fc_stack(self, dropout_masks=None, dropout_seed=None, stacks=None, on_before_params=None, dropout=True,
                  on_before_fc7=None):
        """RoIFeatureTransform -> RoIFeatureBoost -> (fc6 -> Relu -> Dropout -> fc7 -> Relu -> Dropout) per stack.

        Dropout follows ``DropoutIfTraining`` (detectron/modeling/wsl_heads.py:1259-1267): ALWAYS on (ratio 0.5) when the
        model trains -- from the injected masks (parity runs), else from ``dropout_seed``, else from a seed derived from
        the iteration count; only an explicit ``dropout=False`` turns it off (the oracle comparisons without masks)."""
        bl, H = self.blobs, self.H
        stacks = list(range(self.S)) if stacks is None else stacks
        need_argmax = self.train and not self.freeze_conv_body
        roi_feat, argmax = self._timed("roi_pool_f", lambda: ops.RoIPoolF(
            bl["conv5"], bl["rois"], pooled_h=self.roi_size, pooled_w=self.roi_size, spatial_scale=self.spatial_scale,
            is_test=not need_argmax, boost=bl["obn_scores"], x_layout="NHWC", y_layout="NHWC", out_dtype=self.dtype))
        R = roi_feat.shape[0]
        feat = roi_feat.view(R, self.D)
        feat_lo = None
        if self.x3:
            feat_lo = self._lo("roi_feat", feat)
        elif self.tf32:
            ops.round_to_tf32(feat, out=feat)     # GEMM operand: nearest-TF32 (the stand-alone RoIPoolF op stays exact)
        bl["roi_feat"], bl["_argmax_roi_feat"], bl["_roi_feat_lo"] = feat, argmax, feat_lo
        if on_before_params is not None:
            on_before_params()            # first parameter read of the step follows (fc6)
        if self.x3:
            # the update kernels refresh the high parts only: join every update and re-split the masters
            if on_before_fc7 is not None:
                on_before_fc7()
                on_before_fc7 = None
            if self.train:
                self.sync_shadow()
        rt = self.tf32 and not self.x3        # single-pass TF32: epilogues store operands rounded to the nearest TF32
        nS = len(stacks)
        drop6 = self._scratch("drop6", (R, nS * H), self.dtype)
        drop7 = self._scratch("drop7", (R, nS * H), self.dtype)
        use_drop = bool(self.train and dropout)
        if use_drop and dropout_masks is None:
            if dropout_seed is None:
                dropout_seed = self.iter_count + 1
            if int(dropout_seed) <= 0:
                raise RuntimeError("dropout_seed must be positive (pass dropout=False to train without Dropout)")
        dropout_seed = int(dropout_seed or 0)
        self._dropped = use_drop
        m6 = m7 = None
        if use_drop and dropout_masks is not None:
            names = [("drop6", "drop7"), ("noisy_drop6", "noisy_drop7")]
            m6 = torch.cat([dropout_masks[names[s][0]] for s in stacks], dim=1).contiguous()
            m7 = [dropout_masks[names[s][1]] for s in stacks]
        s0, s1 = stacks[0], stacks[-1] + 1
        x3 = self.x3
        self._timed("fc6_fwd", lambda: ops.FC(
            feat, self.w["W6"][s0 * H:s1 * H], self.p["b6"][s0 * H:s1 * H], relu=True, dropout=use_drop,
            dropout_mask=m6, dropout_seed=(dropout_seed * 4 + 1) if (use_drop and m6 is None) else 0, out=drop6,
            round_tf32=rt, gate=self.fc6_gate if (nS == self.S and not x3) else None,
            X_lo=feat_lo, W_lo=self.wl["W6"][s0 * H:s1 * H] if x3 else None))
        drop6_lo = self._lo("drop6", drop6)
        if on_before_fc7 is not None:
            on_before_fc7()               # first read of the fc7 / fc8 parameters follows
        if nS == self.S:      # all stacks: one launch ([S, R, H] views of the column blocks; nothing is copied)
            ops.FC(self._stacked(drop6, nS), self.w["W7"], self.p["b7"], relu=True, dropout=use_drop,
                   dropout_mask=None if m7 is None else torch.stack(m7),
                   dropout_seed=(dropout_seed * 4 + 2) if (use_drop and m7 is None) else 0,
                   out=self._stacked(drop7, nS), round_tf32=rt,
                   X_lo=self._stacked(drop6_lo, nS) if x3 else None, W_lo=self.wl["W7"] if x3 else None)
        else:
            for i, s in enumerate(stacks):
                ops.FC(drop6[:, i * H:(i + 1) * H], self.w["W7_%d" % s], self.p["b7_%d" % s], relu=True, dropout=use_drop,
                       dropout_mask=None if m7 is None else m7[i],
                       dropout_seed=(dropout_seed * 4 + 2 + s) if (use_drop and m7 is None) else 0,
                       out=drop7[:, i * H:(i + 1) * H], round_tf32=rt,
                       X_lo=drop6_lo[:, i * H:(i + 1) * H] if x3 else None, W_lo=self.wl["W7_%d" % s] if x3 else None)
        drop7_lo = self._lo("drop7", drop7)
        bl["drop6_cat"], bl["drop7_cat"], bl["_drop6_lo"], bl["_drop7_lo"] = drop6, drop7, drop6_lo, drop7_lo
        return drop6, drop7

    @staticmethod
    def _stacked(buf, nS):
        """[R, nS*H] activation buffer -> [nS, R, H] strided view of its per-stack column blocks."""
        R = buf.shape[0]
        return buf.view(R, nS, buf.shape[1] // nS).permute(1, 0, 2)

    def _fc8(self, drop7, stacks=None):
        """fc8c | fc8d of every stack: one [R, 2C] GEMM per stack (fp32 logits)."""
        H, C2 = self.H, 2 * self.C
        stacks = list(range(self.S)) if stacks is None else stacks
        R = drop7.shape[0]
        ld = _round_up(C2, 8)
        logits = self._scratch("logits", (len(stacks), R, ld), torch.float32)
        lo7 = self.blobs.get("_drop7_lo") if self.x3 else None
        if len(stacks) == self.S:
            ops.FC(self._stacked(drop7, self.S), self.w["W8"], self.p["b8"], out=logits[:, :, :C2],
                   X_lo=self._stacked(lo7, self.S) if self.x3 else None, W_lo=self.wl["W8"] if self.x3 else None)
        else:
            for i, s in enumerate(stacks):
                ops.FC(drop7[:, i * H:(i + 1) * H], self.w["W8_%d" % s], self.p["b8_%d" % s], out=logits[i][:, :C2],
                       X_lo=lo7[:, i * H:(i + 1) * H] if self.x3 else None, W_lo=self.wl["W8_%d" % s] if self.x3 else None)
        self.blobs["fc8_logits"] = logits
        return logits

    # ------------------------------------------------------------------ the reference's builder names
    def RunTrainStep(self, dropout_masks=None, dropout_seed=None, need_dX=False, fc6_panels=1, on_small_grads=None,
                     on_fc6_panel=None, on_before_params=None, dropout=True, on_before_fc7=None):
        """One fwd+bwd pass of the head on the fed blobs (the slice of ``workspace.RunNet(net)``,
        detectron/utils/train_wsl.py:59, that lies between conv5 and the parameter gradients).
        Gradients land in ``self.g`` / ``self.flat_grad``; returns the blob dict.

        Data-parallel hooks (dp.py): the dominant fc6 weight gradient is computed right after the
        activation-gradient chain in ``fc6_panels`` row panels and ``on_fc6_panel(r0, r1)`` fires
        after each one (its exchange then overlaps the next panel's GEMM and the fc7 / fc8
        weight-gradient GEMMs); ``on_small_grads()`` fires once those are enqueued.  The bias
        gradients of fc6 are complete with the last panel, all others with ``on_small_grads``.
        ``on_before_params()`` fires after RoI pooling, right before the first parameter read (fc6):
        the place to join a parameter update that is still in flight from the previous step; ``on_before_fc7()`` fires
        before the first read of the fc7 / fc8 parameters (their update may land while fc6 runs)."""
        if not self.train:
            raise RuntimeError("RunTrainStep on a test-mode model")
        bl, H, C, C2 = self.blobs, self.H, self.C, 2 * self.C
        drop6, drop7 = self._fc_stack(dropout_masks, dropout_seed, on_before_params=on_before_params, dropout=dropout,
                                      on_before_fc7=on_before_fc7)
        logits = self._fc8(drop7)
        R = drop7.shape[0]
        ld = logits.shape[2]
        dlog = self._scratch("dlogits", (self.S, R, ld), torch.float32)
        gout = {"d_fc8c": dlog[0][:, :C], "d_fc8d": dlog[0][:, C:C2]}
        if self.noise:
            gout.update(d_nfc8c=dlog[1][:, :C], d_nfc8d=dlog[1][:, C:C2])
        out = self._timed("mil_head", lambda: ops.mil_head(
            logits[0][:, :C], logits[0][:, C:C2], bl["rois"], bl["roi_offsets"], bl["labels_oh"],
            logits[1][:, :C] if self.noise else None, logits[1][:, C:C2] if self.noise else None,
            entropy=self.entropy, is_mean=self.mean_loss, backward=True, grads_out=gout))
        bl.update(out)
        bl["loss_cls"] = out["loss"][:, 0]
        if self.noise:
            bl["loss_cls_noise"] = out["loss"][:, 1]
        # ---- backward (Caffe2 AddGradientOperators order: fc8 -> fc7 -> fc6) ----
        if self.dtype == torch.bfloat16:
            dl = self._scratch("dlogits_lp", (self.S, R, ld), torch.bfloat16)
            ops.to_bf16(dlog.view(self.S * R, ld), out=dl.view(self.S * R, ld))
        else:
            dl = self._scratch("dlogits_lp", (self.S, R, ld), torch.float32)
            if self.x3:
                dl_lo = self._scratch("dlogits_lo", (self.S, R, ld), torch.float32)
                ops.split_tf32(dlog.view(self.S * R, ld), hi=dl.view(self.S * R, ld), lo=dl_lo.view(self.S * R, ld))
            else:
                ops.round_to_tf32(dlog.view(self.S * R, ld), out=dl.view(self.S * R, ld))
        d6 = self._scratch("d_fc6", (R, self.S * H), self.dtype)
        d7 = self._scratch("d_fc7", (R, self.S * H), self.dtype)
        # activation-gradient chain first (fc8 dX -> fc7 dX): it is the critical path to the fc6 weight
        # gradient, which carries 86 % of the gradient bytes and so must start its exchange earliest
        S = self.S
        dl3, a7, a6 = dl[:, :, :C2], self._stacked(drop7, S), self._stacked(drop6, S)
        d73, d63 = self._stacked(d7, S), self._stacked(d6, S)
        x3, rt = self.x3, self.tf32 and not self.x3
        dl3_lo = dl_lo[:, :, :C2] if x3 else None
        ops.FCGradientX(dl3, self.w["W8"], act_below=a7, dropout=self._dropped, out=d73, round_tf32=rt,
                        dY_lo=dl3_lo, W_lo=self.wl["W8"] if x3 else None)
        d7_lo = self._lo("d_fc7", d7)
        d73_lo = self._stacked(d7_lo, S) if x3 else None
        ops.FCGradientX(d73, self.w["W7"], act_below=a6, dropout=self._dropped, out=d63, round_tf32=rt,
                        dY_lo=d73_lo, W_lo=self.wl["W7"] if x3 else None)
        d6_lo = self._lo("d_fc6", d6)
        a7_lo = self._stacked(bl["_drop7_lo"], S) if x3 else None
        a6_lo = self._stacked(bl["_drop6_lo"], S) if x3 else None
        if need_dX:
            # before the fc6 panels: a data-parallel exchange may refresh the W6 operands right behind them
            if bl["_argmax_roi_feat"] is None:
                raise RuntimeError("need_dX requires freeze_conv_body=False (argmax is not kept otherwise)")
            d_feat = ops.FCGradientX(d6, self.w["W6"], out_dtype=self.dtype, dY_lo=d6_lo, W_lo=self.wl["W6"] if x3 else None)
            N, Hh, Ww, Cc = bl["conv5"].shape
            bl["d_conv5"] = ops.RoIPoolFGradient(bl["conv5"], bl["rois"], bl["_argmax_roi_feat"],
                                                 d_feat.view(R, self.roi_size, self.roi_size, Cc),
                                                 boost=bl["obn_scores"], layout="NHWC")
        rows = self.S * H
        step = _round_up((rows + fc6_panels - 1) // fc6_panels, 256)
        for r0 in range(0, rows, step):
            r1 = min(rows, r0 + step)
            self._timed("fc6_bwd_w", lambda: ops.FCGradientW(d6[:, r0:r1], bl["roi_feat"], dW=self.g["W6"][r0:r1],
                                                             db=self.g["b6"][r0:r1],
                                                             dY_lo=d6_lo[:, r0:r1] if x3 else None, X_lo=bl["_roi_feat_lo"]))
            if on_fc6_panel is not None:
                on_fc6_panel(r0, r1)
        ops.FCGradientW(dl3, a7, dW=self.g["W8"], db=self.g["b8"], dY_lo=dl3_lo, X_lo=a7_lo)
        ops.FCGradientW(d73, a6, dW=self.g["W7"], db=self.g["b7"], dY_lo=d73_lo, X_lo=a6_lo)
        if on_small_grads is not None:
            on_small_grads()
        return bl

    def RunTestNet(self, want_cls_prob=True):
        """Test-time forward (detectron/core/test_wsl.py:142 ``RunNet``): clean stack only, no dropout;
        ``cls_prob`` [R, C+1] with column 0 duplicated (detectron/modeling/wsl_heads.py:57-67).
        ``want_cls_prob=False`` stops at ``rois_pred`` (test_time.im_detect_bbox builds cls_prob while it
        maps the scores back to the original boxes)."""
        bl, C, C2 = self.blobs, self.C, 2 * self.C
        self._join_side_streams()
        was_train, self.train = self.train, False
        try:
            _, drop7 = self._fc_stack(stacks=[0])
            logits = self._fc8(drop7, stacks=[0])
        finally:
            self.train = was_train
        R = drop7.shape[0]
        B = bl["roi_offsets"].numel() - 1
        labels = bl.get("labels_oh")
        if labels is None or labels.shape[0] != B:          # labels do not enter the test-time outputs
            labels = torch.zeros((B, C), dtype=torch.float32, device=self.device)
        out = ops.mil_head(logits[0][:, :C], logits[0][:, C:C2], bl["rois"], bl["roi_offsets"], labels,
                           entropy=False, is_mean=self.mean_loss, backward=False)
        rp = out["rois_pred"]
        bl["rois_pred"] = rp
        if not want_cls_prob:
            return rp
        bl["cls_prob"] = ops.scatter_scores(rp)                 # Split/Concat of the reference: a copy, no arithmetic
        return bl["cls_prob"]

    # ------------------------------------------------------------------ optimizer (single GPU; dp.py adds the all-reduce)
    def UpdateWorkspaceLr(self, lr, scale_momentum=True, scale_momentum_threshold=1.1):
        """``DetectionModelHelper.UpdateWorkspaceLr`` (detectron/modeling/detector.py:509-560): write the ``lr`` blob and,
        when the rate changes by more than SOLVER.SCALE_MOMENTUM_THRESHOLD (1.1) from a rate above 1e-7, scale the update
        history of every trainable parameter by ``new_lr / cur_lr`` (``_CorrectMomentum``: V := mu*V + lr*grad is not
        independent of lr) -- at the flickr schedule's 1e-3 -> 1e-4 step all momenta are multiplied by 0.1.  The current
        rate is tracked on the host (the blob starts at 0, optimizer_wsl.py:81-85, so the first call only writes it).
        Returns the factor applied to the momenta (1.0 when they were left alone)."""
        new_lr, cur_lr = np.float32(lr), self._lr_host
        factor = lr_change_correction(cur_lr, new_lr, scale_momentum, scale_momentum_threshold)
        if cur_lr != new_lr or factor != 1.0:
            # the reference calls this before every RunNet; the previous step's update pipeline (dp.py) may still be
            # reading `lr` / rewriting the momenta on its side stream -- join it before either changes
            self._join_side_streams()
        if cur_lr != new_lr:
            self.lr.fill_(float(new_lr))
            self._lr_host = new_lr
        if factor != 1.0:
            s = torch.full((1,), float(factor), dtype=torch.float32, device=self.device)
            ops.RoIFeatureBoost(self.flat_mom.view(1, -1), s, out=self.flat_mom.view(1, -1))   # Scale([m] -> [m]) in place
        return float(factor)

    def param_update(self, momentum=0.9, weight_decay=5e-4, gpu_num=1, iter_size=1):
        """``ACMWeightDecayMomentumSGDUpdate`` for every parameter (detectron/modeling/optimizer_wsl.py:96-137):
        weights wd=5e-4, lr_mult=1; biases wd=0, lr_mult=2.  Two fused launches over the flat buffers."""
        if iter_size != 1:
            raise RuntimeError("iter_size > 1 needs an accumulator; use ops.ACMWeightDecayMomentumSGDUpdate directly")
        nw, nt = self.n_weights, self.n_total
        lp = self.flat_lp
        ops.ACMWeightDecayMomentumSGDUpdate(self.flat_grad[:nw], self.flat_mom[:nw], self.lr, self.flat_param[:nw], None,
                                            momentum=momentum, gpu_num=gpu_num, lr_mult=1.0, weight_decay=weight_decay,
                                            iter_count=self.iter_count, p_shadow=lp[:nw])
        ops.ACMWeightDecayMomentumSGDUpdate(self.flat_grad[nw:nt], self.flat_mom[nw:nt], self.lr, self.flat_param[nw:nt],
                                            None, momentum=momentum, gpu_num=gpu_num, lr_mult=2.0, weight_decay=0.0,
                                            iter_count=self.iter_count, p_shadow=lp[nw:nt])
        self.iter_count += 1


# ---------------------------------------------------------------------------------------------
# Reference builder names (eager): the functions a caller of detectron/modeling/*_heads.py knows
# ---------------------------------------------------------------------------------------------
def add_VGG16_roi_2fc_head(model, blob_in=None, dim_in=None, spatial_scale=None, prefix=""):
    """detectron/modeling/wsl_heads.py:654-681 (clean stack only).  Returns (drop7, 4096)."""
    _, drop7 = model._fc_stack(stacks=[0])
    return drop7[:, :model.H], model.H


def add_VGG16_roi_2fc_noise_head(model, blob_in=None, dim_in=None, spatial_scale=None, prefix="", dropout_masks=None,
                                 dropout_seed=None, dropout=True):
    """detectron/modeling/webly_heads.py:463-502.  Returns ([drop7, noisy drop7], [4096, 4096])."""
    _, drop7 = model._fc_stack(dropout_masks, dropout_seed, dropout=dropout)
    H = model.H
    return [drop7[:, s * H:(s + 1) * H] for s in range(model.S)], [H] * model.S


def add_webly_outputs(model, blob_in=None, dim=None, prefix=""):
    """detectron/modeling/webly_heads.py:32-74 (and wsl_heads.add_wsl_outputs for the clean stack):
    the fc8 logits; the softmaxes / product are evaluated by ``add_webly_losses``' fused kernel."""
    return model._fc8(model.blobs["drop7_cat"])
