#!/usr/bin/env python
"""Build oracle/_ref/libnawsod_ref_kernels.so (TEST INFRASTRUCTURE): the reference's own CUDA kernels for RoIIoU,
the in-tree RoI max-pooling clone (RoILoopPool) and MinEntropyLoss, run on the HOST.

Those operators exist only as CUDA code in the reference (no CPU implementation), and their files cannot be compiled
as a whole without the Caffe2 runtime.  Each file keeps its device code in one anonymous namespace; this script cuts
that block out of the file where it lies under /root/reference -- byte for byte, nothing is edited -- into
oracle/_ref/gen/*.inc (git-ignored, like everything under oracle/_ref/) and compiles oracle/ref_kernels_driver.cc,
which includes the blocks behind oracle/cuda_host_shim.h.  No reference source is copied into the repository.

    python oracle/build_ref_kernels.py        (a no-op where /root/reference does not exist: the GPU box uses the prebuilt .so)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NAWSOD_REFERENCE", "/root/reference")
OPS = os.path.join(REF, "detectron", "ops")
FILES = {"roi_iou_op.cu": "roi_iou_kernels.inc", "roi_loop_pool_op.cu": "roi_loop_pool_kernels.inc",
         "min_entropy_loss_op.cu": "min_entropy_loss_kernels.inc"}


def device_block(path):
    """The lines strictly between the first `namespace {` and the `} // namespace` that closes it."""
    lines = open(path).read().split("\n")
    start = next(i for i, ln in enumerate(lines) if ln.strip() == "namespace {")
    end = next(i for i in range(start + 1, len(lines)) if lines[i].replace(" ", "") == "}//namespace")
    block = lines[start + 1:end]
    if not any("__global__" in ln for ln in block):
        raise RuntimeError("no kernel found in the anonymous namespace of " + path)
    return "\n".join(block) + "\n"


def main():
    if not os.path.isdir(OPS):
        sys.stderr.write("build_ref_kernels: %s not present (GPU box): keeping prebuilt oracle/_ref\n" % OPS)
        return 0
    gen = os.path.join(HERE, "_ref", "gen")
    os.makedirs(gen, exist_ok=True)
    for src, dst in FILES.items():
        with open(os.path.join(gen, dst), "w") as f:
            f.write("// generated from %s by oracle/build_ref_kernels.py -- do not commit\n" % os.path.join(OPS, src))
            f.write(device_block(os.path.join(OPS, src)))
    out = os.path.join(HERE, "_ref", "libnawsod_ref_kernels.so")
    # -ffp-contract=off: the host must not fuse what the kernels write as separate operations
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-I" + HERE,
                           "-I" + os.path.join(HERE, "_ref"), os.path.join(HERE, "ref_kernels_driver.cc"), "-o", out])
    print("built " + out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
