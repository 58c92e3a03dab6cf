"""Golden vectors for the frozen VGG16 conv body.  Run in the BUILD container (needs /root/reference):

    python tests/golden/make_golden_vgg16_body.py        -> tests/golden/vgg16_body.npz

Imports detectron/modeling/VGG16.py UNMODIFIED and runs ``add_VGG16_conv5_body_origin`` -- the MODEL.CONV_BODY of the
shipped flickr_voc config, merged here, so WSL.DILATION is the shipped 2 -- and once more with WSL.DILATION = 1, on a model
helper that (a) records every operator the builder emits (type, input, output, arguments) and (b) executes it at once on a
float32 NumPy workspace (oracle.conv_body_oracle.run_op: torch CPU float32 conv2d / relu / max_pool2d stand in for the
Caffe2 built-ins, whose sources are not in the tree).  Stored: the two operator traces, the builder's return values, and for
one small image the input and the conv5_3 / pool4 / conv3_3 blobs.  The 14.7 M weights are regenerated from the seed
(oracle.conv_body_oracle.synth_params; checksums stored).
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import conv_body_oracle as CB     # noqa: E402


class TracingModel:
    """The builder-facing surface of DetectionModelHelper that VGG16.py uses: Conv / Relu / MaxPool / StopGradient."""

    def __init__(self, ws, params):
        self.ws, self.params, self.trace = ws, params, []

    def _rec(self, kind, src, dst, **args):
        self.trace.append("%s(%s)->(%s)%s" % (kind, src, dst, "".join(" %s=%s" % kv for kv in sorted(args.items()))))

    def Conv(self, blob_in, blob_out, dim_in, dim_out, kernel, **kw):
        args = dict(dim_in=dim_in, dim_out=dim_out, kernel=kernel, **kw)
        self._rec("Conv", blob_in, blob_out, **args)
        w, b = self.params[blob_out + "_w"], self.params[blob_out + "_b"]
        assert w.shape == (dim_out, dim_in, kernel, kernel)
        self.ws[blob_out] = CB.run_op("Conv", self.ws[blob_in], args, w, b)
        return blob_out

    def Relu(self, blob_in, blob_out):
        self._rec("Relu", blob_in, blob_out)
        self.ws[blob_out] = CB.run_op("Relu", self.ws[blob_in], {})
        return blob_out

    def MaxPool(self, blob_in, blob_out, **kw):
        self._rec("MaxPool", blob_in, blob_out, **kw)
        self.ws[blob_out] = CB.run_op("MaxPool", self.ws[blob_in], kw)
        return blob_out

    def StopGradient(self, blob_in, blob_out):
        self._rec("StopGradient", blob_in, blob_out)
        return blob_out


def main():
    spec = importlib.util.spec_from_file_location("make_golden_roi_data", os.path.join(HERE, "make_golden_roi_data.py"))
    maker = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(maker)
    sys.meta_path.insert(0, maker._Absent())
    import future.utils
    future.utils.iteritems = lambda d: iter(d.items())
    sys.path.insert(0, "/root/reference")
    from detectron.core.config import cfg, merge_cfg_from_file
    import yaml
    import detectron.utils.env as envu
    envu.yaml_load = lambda f: yaml.load(f, Loader=yaml.SafeLoader)
    merge_cfg_from_file("/root/reference/configs/flickr_voc/na_wsddn_V-16-C5_1x.yaml")
    import detectron.modeling.VGG16 as VGG16
    assert cfg.MODEL.CONV_BODY == "VGG16.add_VGG16_conv5_body_origin" and cfg.WSL.DILATION == 2 and cfg.TRAIN.FREEZE_CONV_BODY

    seed = 77
    params = CB.synth_params(seed)
    rng = np.random.default_rng(5)
    data = (rng.standard_normal((1, 3, 40, 56)) * 50.0).astype(np.float32)      # mean-subtracted pixels: O(50)
    out = {"seed": np.int32(seed), "param_checksum": CB.param_checksum(params), "data": data}
    for tag, dil in (("d2", 2), ("d1", 1)):
        cfg.immutable(False)
        cfg.WSL.DILATION = dil
        m = TracingModel({"data": data}, params)
        blob, dim, scale = VGG16.add_VGG16_conv5_body_origin(m)
        assert blob == "conv5_3"
        out[tag + "_trace"] = np.array(m.trace)
        out[tag + "_dim_out"] = np.int32(dim)
        out[tag + "_spatial_scale"] = np.float64(scale)
        for k in ("conv3_3", "pool4", "conv5_3"):
            out[tag + "_" + k] = m.ws[k]
        print(tag, len(m.trace), "operators; conv5_3", m.ws["conv5_3"].shape, "scale", scale,
              "mean", float(m.ws["conv5_3"].mean()), "zeros", float((m.ws["conv5_3"] == 0).mean()))
    np.savez_compressed(os.path.join(HERE, "vgg16_body.npz"), **out)
    print("wrote vgg16_body.npz")


if __name__ == "__main__":
    main()
