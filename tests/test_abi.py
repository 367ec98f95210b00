"""The C-ABI library loads without a GPU and exports every symbol include/mcb200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mcb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mcb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(mcb):
    lib = ctypes.CDLL(os.path.join(ROOT, "open-solution-mapping-challenge_b200", "libmcb200.so"))
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.mcb_version.restype = ctypes.c_int
    assert lib.mcb_version() >= 100


def test_compute_entry_points_fail_loudly_without_gpu(mcb):
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    from mcb200 import _lib as L
    a = L.ConvFwdArgs()
    rc = L.lib.mcb_conv_fwd(ctypes.byref(a), None)
    assert rc != 0 and len(L.lib.mcb_last_error()) > 0
    from mcb200.unet_models import UNetResNet
    net = UNetResNet(34, 2, 32, 0.0, False, True)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 64, 64))
    from mcb200 import postprocessing as pp
    with pytest.raises(RuntimeError):
        pp.label_multilayer_image(__import__("numpy").zeros((2, 8, 8), bool))


def test_postprocessor_honours_the_step_transformer_contract(tmp_path):
    """src/steps/base.py:254-269: a Step calls fit_transform / save / load on its transformer and may pickle it"""
    import pickle
    import mcb200  # noqa: F401
    from mcb200.postprocessing import MaskPostprocessor
    pp = MaskPostprocessor((300, 300), "crop", 2, 2)
    assert pp.fit() is pp and pp.load(str(tmp_path / "x")) is pp
    pp.save(str(tmp_path / "pp.pkl"))
    assert (tmp_path / "pp.pkl").exists()
    pp.__dict__["_graphs"] = {"k": object()}          # stands in for captured graphs
    clone = pickle.loads(pickle.dumps(pp))
    assert clone.target_size == (300, 300) and clone.mode == "crop" and (clone.erode, clone.dilate) == (2, 2)
    assert "_graphs" not in clone.__dict__
