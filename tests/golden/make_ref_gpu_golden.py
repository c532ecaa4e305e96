"""Generate tests/golden/ref_gpu_*.npz by RUNNING THE REFERENCE'S OWN CUDA KERNELS on a B200.

    make -C oracle ref                       # build container: compiles the unmodified reference
                                             # .cu files for sm_100a into oracle/_ref/
    gpurun -- python tests/golden/make_ref_gpu_golden.py      # GPU box; writes gpurun_out/golden/
    cp gpurun_out/golden/ref_gpu_*.npz tests/golden/

Each fixture stores the inputs next to the reference outputs: nms_cuda_compute keep lists,
ROIAlignForward/Backward, ROIPoolForward/Backward, and the RoIAlignAvg / RoIAlignMax module
outputs (reference kernel followed by torch's avg_pool2d / max_pool2d, modules/roi_align.py:26-42).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_gpu  # noqa: E402
import _cases  # noqa: E402


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    nms = {}
    for name, (dets, thr) in _cases.nms_cases().items():
        keep = ref_gpu.nms(torch.from_numpy(dets).to(dev), thr)
        nms[name + "__dets"] = dets
        nms[name + "__thresh"] = np.float32(thr)
        nms[name + "__keep"] = keep.cpu().numpy().astype(np.int32)
        print("nms %-10s n=%5d kept=%5d" % (name, len(dets), len(keep)))
    np.savez_compressed(os.path.join(out_dir, "ref_gpu_nms.npz"), **nms)

    for name, c in _cases.roi_cases().items():
        feat = torch.from_numpy(c["features"]).to(dev)
        rois = torch.from_numpy(c["rois"]).to(dev)
        ah, aw, s = c["ah"], c["aw"], c["scale"]
        out = dict(features=c["features"], rois=c["rois"], scale=np.float32(s), ah=ah, aw=aw)
        y = ref_gpu.roi_align_forward(feat, rois, ah, aw, s)
        out["align_fwd"] = y.cpu().numpy()
        td = torch.from_numpy(_cases.top_diff_for(c, ah, aw)).to(dev)
        out["align_top_diff"] = td.cpu().numpy()
        out["align_bwd"] = ref_gpu.roi_align_backward(td, rois, feat.shape, s).cpu().numpy()
        # module level: RoIAlignAvg / RoIAlignMax with aligned size (ah-1, aw-1)
        if ah > 1 and aw > 1:
            out["align_avg_fwd"] = F.avg_pool2d(y, kernel_size=2, stride=1).cpu().numpy()
            out["align_max_fwd"] = F.max_pool2d(y, kernel_size=2, stride=1).cpu().numpy()
            td2 = torch.from_numpy(_cases.top_diff_for(c, ah - 1, aw - 1, seed=43)).to(dev)
            out["pooled_top_diff"] = td2.cpu().numpy()
            for mode, fn in (("avg", F.avg_pool2d), ("max", F.max_pool2d)):
                yy = y.clone().requires_grad_(True)
                fn(yy, kernel_size=2, stride=1).backward(td2)
                out["align_%s_bwd" % mode] = ref_gpu.roi_align_backward(
                    yy.grad.contiguous(), rois, feat.shape, s).cpu().numpy()
        ph, pw = max(ah - 1, 1), max(aw - 1, 1)
        p, am = ref_gpu.roi_pool_forward(feat, rois, ph, pw, s)
        out["pool_fwd"] = p.cpu().numpy()
        out["pool_argmax"] = am.cpu().numpy()
        td3 = torch.from_numpy(_cases.top_diff_for(c, ph, pw, seed=47)).to(dev)
        out["pool_top_diff"] = td3.cpu().numpy()
        out["pool_bwd"] = ref_gpu.roi_pool_backward(td3, am, rois, feat.shape, s).cpu().numpy()
        torch.cuda.synchronize()
        path = os.path.join(out_dir, "ref_gpu_roi_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("roi %-18s %6.1f KB" % (name, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
