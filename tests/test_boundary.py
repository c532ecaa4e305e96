"""CPU-side checks of the drop-in boundary (no GPU needed): the C-ABI library loads and exports
every symbol include/nafae_b200.h declares, the product never touches oracle/, the Python surface
mirrors the reference's module paths / names / signatures."""
import ctypes
import inspect
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nafae_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|size_t|char\s*\*|const char\s*\*)\s*\*?\s*(\w+)\s*\(",
                       src, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_reference_entry_points():
    names = _declared_symbols()
    # the symbols the reference's cffi glue binds (nms_cuda_kernel.h:5-6, roi_align_kernel.h:13-27,
    # roi_pooling_kernel.h:8-18) must be exported under the same names
    for ref in ("nms_cuda_compute", "ROIAlignForwardLaucher", "ROIAlignBackwardLaucher",
                "ROIPoolForwardLaucher", "ROIPoolBackwardLaucher"):
        assert ref in names
    assert len(names) >= 17


def test_library_exports_every_declared_symbol():
    from nafae_b200 import _C
    lib = ctypes.CDLL(_C.LIB_PATH)
    missing = [n for n in _declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    assert not _C.MISSING
    assert set(_declared_symbols()) == set(_C.SIGNATURES), "ctypes table out of sync with the header"
    assert lib.nafae_abi_version() == 3 == _C.ABI_VERSION


def test_python_constants_equal_the_header_macros():
    """Flags and pooling modes are passed as plain integers through ctypes: they must be the header's."""
    import re
    from nafae_b200 import _C
    hdr = open(os.path.join(ROOT, "include", "nafae_b200.h")).read()
    macros = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+NAFAE_(\w+)\s+(\d+)u?\b", hdr)}
    for name in ("FLAG_EXACT", "FLAG_NO_GATE", "FLAG_OUT_BF16", "FLAG_OVERWRITE", "FLAG_DETERMINISTIC",
                 "POOL_NONE", "POOL_AVG", "POOL_MAX"):
        assert macros[name] == getattr(_C, name), name
    flags = [v for k, v in macros.items() if k.startswith("FLAG_")]
    assert len(set(flags)) == len(flags) and all(v & (v - 1) == 0 for v in flags)  # distinct single bits


def test_invalid_arguments_return_zero_without_a_gpu():
    """Argument validation happens before any CUDA call: status 0 + message, never exit()."""
    from nafae_b200 import _C
    st = _C.lib.nafae_proposal_tail(None, None, 2, 10, 6000, 0, 0.7, None, None, None, None)
    assert st == 0 and "post_nms_topn" in _C.last_error()
    st = _C.lib.nafae_roi_align_forward(None, 0.0625, 1, 4, 14, 14, 8, 7, 7, 9, None, None, 0, None,
                                        0, None)
    assert st == 0 and "pool_mode" in _C.last_error()
    st = _C.lib.nafae_ground_forward(None, None, None, 0, 5, 20, 13, 512, 10.0, 4.13, 1, None, None,
                                     None, None, 0, None)
    assert st == 0
    assert _C.lib.nafae_nms_workspace_bytes(40, 2352) == 40 * 2352 * 37 * 8
    assert _C.lib.nafae_ground_workspace_bytes(8, 5, 20, 13, 512) > 0


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under nafae_b200/ may import, load or name it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "nafae_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "libnafae_oracle" in txt \
                        or "libnafae_ref" in txt:
                    bad.append(os.path.join(dirpath, fn))
    assert not bad, bad


def test_no_cpu_fallback_in_the_operator_surface():
    import torch
    from nafae_b200.model.nms.nms_wrapper import nms
    from nafae_b200.model.roi_align.modules.roi_align import RoIAlignAvg
    from nafae_b200.model.roi_pooling.modules.roi_pool import _RoIPooling
    from nafae_b200.grounding import ground
    with pytest.raises(NotImplementedError):
        nms(torch.zeros(4, 5), 0.7)
    with pytest.raises(NotImplementedError):
        RoIAlignAvg(7, 7, 1 / 16.)(torch.zeros(1, 8, 14, 14), torch.zeros(2, 5))
    with pytest.raises(NotImplementedError):
        _RoIPooling(7, 7, 1 / 16.)(torch.zeros(1, 8, 14, 14), torch.zeros(2, 5))
    with pytest.raises(NotImplementedError):
        ground(torch.zeros(20, 16), torch.zeros(4, 16), [1], 1, 20, 4, 10.0, 1.0, True)
    assert nms(torch.zeros(0, 5), 0.7) == []  # nms_wrapper.py:13-14


def test_python_surface_mirrors_reference_names():
    from nafae_b200.model.nms import nms_wrapper, nms_gpu
    from nafae_b200.model.roi_align.modules import roi_align as ra
    from nafae_b200.model.roi_align.functions import roi_align as raf
    from nafae_b200.model.roi_pooling.modules import roi_pool as rp
    from nafae_b200.model.roi_pooling.functions import roi_pool as rpf
    from nafae_b200 import grounding
    assert list(inspect.signature(nms_wrapper.nms).parameters) == ["dets", "thresh", "force_cpu"]
    assert list(inspect.signature(nms_gpu.nms_gpu).parameters) == ["dets", "thresh"]
    for cls in (ra.RoIAlign, ra.RoIAlignAvg, ra.RoIAlignMax):
        ps = list(inspect.signature(cls.__init__).parameters)
        assert ps[:4] == ["self", "aligned_height", "aligned_width", "spatial_scale"]
        assert list(inspect.signature(cls.forward).parameters) == ["self", "features", "rois"]
    assert hasattr(raf, "RoIAlignFunction") and hasattr(rpf, "RoIPoolFunction")
    ps = list(inspect.signature(rp._RoIPooling.__init__).parameters)
    assert ps == ["self", "pooled_height", "pooled_width", "spatial_scale"]
    assert list(inspect.signature(grounding.DVSA.__init__).parameters) == ["self", "args", "cfg"]
    assert list(inspect.signature(grounding.DVSA.forward).parameters) == [
        "self", "vis_feats", "word_feats", "entities_length"]
    assert list(inspect.signature(grounding.postprocess).parameters) == ["D", "D_sim", "Na", "Ns", "Nb",
                                                                         "Ne"]
    for m in ("init_train", "init_eval"):
        assert hasattr(grounding.DVSA, m)


def test_reference_style_imports_work_with_package_on_syspath(tmp_path):
    """`from model.nms.nms_wrapper import nms` as the reference writes it."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "import nafae_b200;"
            "from nafae_b200.model.nms.nms_wrapper import nms;"
            "from nafae_b200.model.roi_align.modules.roi_align import RoIAlignAvg;"
            "from nafae_b200.model.roi_pooling.modules.roi_pool import _RoIPooling;"
            "print('ok')") % (ROOT, os.path.join(ROOT, "nafae_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_allreduce_abi_argument_checks_without_a_gpu():
    from nafae_b200 import _C
    n = 2201600
    assert _C.lib.nafae_ar_buffer_bytes(n, 8) >= n * 4 + _C.lib.nafae_ar_data_offset()
    assert _C.lib.nafae_ar_buffer_bytes(5, 8) == _C.lib.nafae_ar_data_offset() + 32 * 4  # padded to 4*world
    ptrs = (ctypes.c_void_p * 2)()
    assert _C.lib.nafae_allreduce_avg(ptrs, 0, 9, 64, 8, 128, 0, None) == 0        # world > 8
    assert _C.lib.nafae_allreduce_avg(ptrs, 0, 2, 10, 8, 128, 0, None) == 0        # count not padded
    assert _C.lib.nafae_allreduce_avg(ptrs, 0, 1, 64, 8, 256, 0, None) == 1        # world 1: nothing to do
    assert _C.lib.nafae_allreduce_avg(ptrs, 0, 2, 64, 8, 7, 0, None) == 0          # bad cta_threads


def test_multicast_and_optimizer_abi_argument_checks_without_a_gpu():
    from nafae_b200 import _C
    assert _C.lib.nafae_mc_supported() == 0  # no driver / no NVSwitch here: must answer, not crash
    assert _C.lib.nafae_mc_buffer_bytes(2201600, 8) == _C.lib.nafae_ar_buffer_bytes(2201600, 8)
    h, fd = ctypes.c_void_p(), ctypes.c_int(-1)
    assert _C.lib.nafae_mc_create(8, 1 << 20, ctypes.byref(h), ctypes.byref(fd)) <= 0 and _C.last_error()
    assert _C.lib.nafae_allreduce_mc(None, None, 0, 2, 64, 8, 512, None) == 0      # NULL buffers
    assert _C.lib.nafae_mc_free(None) == 1
    assert _C.lib.nafae_clip_adam_workspace_bytes() >= 64
    assert _C.lib.nafae_clip_adam_step(None, None, None, None, 16, 1e-3, 0.9, 0.999, 1e-8, 1e-5, 100.0,
                                       None, 0, None) == 0
    assert _C.lib.nafae_gate_sync(None, None) == 0


def test_pure_host_modules_import_without_loading_the_cuda_library():
    """synth / evaluate / checkpoint / bridge / the sharding helpers must not dlopen
    libnafae_b200.so (bench.py --impl reference maps only oracle/)."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r);"
            "import nafae_b200, nafae_b200.synth, nafae_b200.evaluate, nafae_b200.checkpoint,"
            " nafae_b200.bridge, nafae_b200.parallel;"
            "assert nafae_b200.parallel.shard_segments(10, 1, 4) == (3, 6);"
            "assert 'nafae_b200._C' not in sys.modules;"
            "assert 'libnafae_b200' not in open('/proc/self/maps').read();"
            "print('ok')") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
