"""Drop-in boundary, build container only: the reference's UNCHANGED src/pipelines.py builds its `unet`,
`unet_weighted`, `unet_padded` and `unet_tta` pipelines with the mcb200 classes bound to the names it imports
(PyTorchUNet*, `post`, the TTA transformers); every Step constructs, the model Step's transformer honours the
transformer contract of src/steps/base.py (save -> transformer_is_cached -> load), and the reference's own callbacks
(src/models.py:60, 295-307) are what the constructor builds.  No GPU: nothing is computed."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")


@pytest.fixture()
def patched(tmp_path, monkeypatch):
    ref_shim.install()
    import mcb200  # noqa: F401
    from mcb200 import loaders as mlo, models as mm, postprocessing as mpp
    import src.loaders as rlo
    import src.pipeline_config as pc
    import src.pipelines as pl
    for name in ("PyTorchUNet", "PyTorchUNetStream", "PyTorchUNetWeighted", "PyTorchUNetWeightedStream"):
        monkeypatch.setattr(pl, name, getattr(mm, name))
    monkeypatch.setattr(pl, "post", mpp)
    monkeypatch.setattr(rlo, "TestTimeAugmentationGenerator", mlo.TestTimeAugmentationGenerator)
    monkeypatch.setattr(rlo, "TestTimeAugmentationAggregator", mlo.TestTimeAugmentationAggregator)

    def plain(d):   # nested attribute dicts built up front (the shim's AttrDict wraps plain children in COPIES on access)
        return ref_shim._AttrDict({k: plain(v) if isinstance(v, dict) else v for k, v in d.items()})

    cfg = plain(pc.SOLUTION_CONFIG)
    cfg["env"]["cache_dirpath"] = str(tmp_path / "cache")
    cfg["unet"]["architecture_config"]["model_params"]["encoder"] = "ResNet34"
    cfg["unet"]["callbacks_config"]["model_checkpoint"]["filepath"] = str(tmp_path / "ckpt" / "unet" / "best.torch")
    return pl, cfg, mm, mpp, mlo


@pytest.mark.parametrize("stream", [False, True])
def test_reference_pipelines_build_with_mcb200_classes(patched, stream):
    pl, cfg, mm, mpp, mlo = patched
    cfg["execution"]["stream_mode"] = stream
    for loader_mode in ("resize", "crop_and_pad"):
        cfg["execution"]["loader_mode"] = loader_mode
        out = pl.unet(cfg, train_mode=True)
        step = out.get_step("unet")
        assert type(step.transformer) is (mm.PyTorchUNetStream if stream else mm.PyTorchUNet)
        assert step.is_trainable and [s.name for s in step.input_steps] == ["loader"]
        names = set(out.all_steps)
        assert {"mask_resize", "category_mapper", "mask_erosion", "labeler", "mask_dilation", "score_builder"} <= names
        # the post-processing Steps wrap OUR functions (same names as src/postprocessing.py)
        assert pl.post is mpp and callable(pl.post.label_multilayer_image)
        w = pl.unet_weighted(cfg, train_mode=True)
        assert type(w.get_step("unet").transformer) is (mm.PyTorchUNetWeightedStream if stream else mm.PyTorchUNetWeighted)
    cfg["execution"]["loader_mode"] = "crop_and_pad"
    padded = pl.unet_padded(cfg)
    assert "prediction_crop" in padded.all_steps
    if not stream:
        tta = pl.unet_tta(cfg)
        assert type(tta.get_step("tta_aggregator").transformer) is mlo.TestTimeAugmentationAggregator
        assert type(tta.get_step("tta_generator").transformer) is mlo.TestTimeAugmentationGenerator


def test_constructor_builds_the_reference_callbacks_and_step_roundtrips_the_transformer(patched):
    pl, cfg, mm, mpp, mlo = patched
    cfg["execution"]["stream_mode"] = False
    cfg["execution"]["loader_mode"] = "resize"
    out = pl.unet_weighted(cfg, train_mode=True)
    step = out.get_step("unet")
    tr = step.transformer
    import src.steps.pytorch.callbacks as rcb
    assert isinstance(tr.callbacks, rcb.CallbackList)
    kinds = [type(c).__name__ for c in tr.callbacks.callbacks]
    assert kinds == ["ExperimentTiming", "TrainingMonitor", "ValidationMonitorSegmentation", "ModelCheckpoint",
                     "ExponentialLRScheduler", "EarlyStopping", "NeptuneMonitorSegmentation"]
    # the callbacks read these attributes in set_params (src/steps/pytorch/callbacks.py:26-32)
    tr.callbacks.set_params(tr, validation_datagen=None, meta_valid=None)
    sched = [c for c in tr.callbacks.callbacks if type(c).__name__ == "ExponentialLRScheduler"][0]
    assert sched.optimizer is tr.optimizer and tr.output_names == ["multichannel_map"]
    lr0 = tr.optimizer.param_groups[0]["lr"]
    sched.on_train_begin()
    tr.optimizer.step = lambda *a, **k: None      # (scheduler warns when stepped before the optimizer; irrelevant here)
    sched.on_epoch_end()
    assert tr.optimizer.param_groups[0]["lr"] == pytest.approx(lr0 * sched.gamma)   # _fit_loop reads param_groups each step
    # Step contract: not cached -> save -> cached -> load restores the weights
    assert not step.transformer_is_cached
    sd0 = {k: v.clone() for k, v in tr.model.state_dict().items()}
    tr.save(step.cache_filepath_step_transformer)
    assert step.transformer_is_cached
    saved = torch.load(step.cache_filepath_step_transformer)
    assert all(k.startswith("module.") for k in saved)          # the reference's DataParallel checkpoints
    with torch.no_grad():
        for p in tr.model.parameters():
            p.add_(1.0)
    tr.load(step.cache_filepath_step_transformer)
    for k, v in tr.model.state_dict().items():
        assert torch.equal(v, sd0[k]), k
    # a checkpoint written by the reference's own network loads too (same keys / shapes)
    import src.unet_models as rum
    ref_net = rum.UNetResNet(34, 2, 32, 0.0, False, True)
    missing = tr.model.load_state_dict(ref_net.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys


def test_fit_loop_metrics_shape_suits_the_reference_callbacks():
    """TrainingMonitor.on_batch_end does loss.data.cpu().numpy()[0] (src/steps/pytorch/callbacks.py:134): the step's
    metric must be a 1-element 1-D tensor (a 0-d tensor raises IndexError there)"""
    import inspect
    import mcb200  # noqa: F401
    from mcb200 import models as mm
    src = inspect.getsource(mm.FusedTrainStep.step)
    assert "reshape(1)" in src
    assert torch.zeros(()).reshape(1).clone().data.cpu().numpy()[0] == 0.0


def test_arena_survives_device_noops_and_rebuilds_on_moves():
    """`.cuda()` on a model that is already there (save_model's round trip ends with it, _to_device runs every
    batch) must keep arenas / plans; a real move re-packs them and bumps the generation the fused step checks"""
    import mcb200  # noqa: F401
    from mcb200.unet_models import UNetResNet
    net = UNetResNet(34, 2, 32, 0.0, False, True)
    gen, p32 = net._generation, net._p32
    net.float()                        # nothing moves
    net.to(torch.device("cpu"))
    assert net._generation == gen and net._p32 is p32 and net._params_alias_arena()
    w = net.final.weight
    assert w.data_ptr() == p32.data_ptr() + 4 * net._slots[id(w)].off
    with pytest.raises(AssertionError):    # a dtype move tears the views off; fp32 masters are required, said loudly
        net.double()
