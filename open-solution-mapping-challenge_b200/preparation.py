"""Input side of the path on the GPU: mirror of the deterministic parts of
/root/reference/src/loaders.py:141-171,311-317 (image / target tensors of the padded loaders),
/root/reference/src/augmentation.py:40-86 (PadFixed) and /root/reference/src/preparation.py:151-195 (two-nearest
building distances, component-size map).

Out of scope: JPEG / PNG decoding, COCO polygon rasterisation (pycocotools), the random imgaug augmenters, and the
PIL-resampled `resize` loader mode (the bench's crop_and_pad mode pads instead).  Everything here is batched device
work in libmcb200.so (csrc/input.cu); no CPU fallback."""
import numpy as np
import torch

from . import _lib as L
from .postprocessing import MEAN, STD, _dev, _to_dev, label_batch

PAD_MODES = {"replicate": 0, "reflect": 1}


def image_transform_batch(images, pad=(0, 0), pad_method="replicate", mean=MEAN, std=STD):
    """images (N, H, W, 3) uint8 (numpy or cuda tensor) -> (N, 3, H + 2 pad_h, W + 2 pad_w) float32 cuda:
    PadFixed(pad, pad_method) -> transforms.ToTensor() -> transforms.Normalize(mean, std), bit-exact"""
    x = _to_dev(images, torch.uint8)
    if x.dim() != 4 or x.shape[3] != 3:
        raise ValueError("expected images (N, H, W, 3) uint8, got %s" % (tuple(x.shape),))
    n, h, w, _ = x.shape
    ph, pw = int(pad[0]), int(pad[1])
    out = torch.empty((n, 3, h + 2 * ph, w + 2 * pw), dtype=torch.float32, device=x.device)
    m = (L.C.c_float * 3)(*[float(np.float32(v)) for v in mean])
    s = (L.C.c_float * 3)(*[float(np.float32(v)) for v in std])
    L.fcall("mcb_image_pad_normalize", x.data_ptr(), out.data_ptr(), n, h, w, ph, pw, PAD_MODES[pad_method], m, s)
    return out


def two_nearest_distances(instance_masks):
    """update_distances + clean_distances (src/preparation.py:151-168) for ONE image: instance_masks (K, H, W) uint8|bool
    (one non-empty plane per building) -> (distances float16 (H, W) numpy, second_nearest float64 (H, W) numpy)"""
    m = np.asarray(instance_masks)
    if m.ndim != 3:
        raise ValueError("expected (K, H, W) instance masks")
    k, h, w = m.shape
    dev = _dev()
    md = _to_dev((m != 0).astype(np.uint8), torch.uint8) if k else None
    ws = torch.empty(max(k, 1) * h * w, dtype=torch.int32, device=dev)
    dsum = torch.empty((h, w), dtype=torch.float16, device=dev)
    second = torch.empty((h, w), dtype=torch.float64, device=dev)
    L.fcall("mcb_edt_two_nearest", None if md is None else md.data_ptr(), k, h, w, ws.data_ptr(), dsum.data_ptr(),
            second.data_ptr())
    return dsum.cpu().numpy(), second.cpu().numpy()


def get_size_matrix(mask):
    """src/preparation.py:189-195: pixel count of each pixel's 4-connected component of `mask`, 1 on background"""
    m = np.asarray(mask)
    md = _to_dev((m != 0).astype(np.uint8), torch.uint8)[None].contiguous()
    labels, counts = label_batch(md, return_counts=True)
    k = int(counts.item())
    if k == 0:
        return np.ones_like(m)           # the reference returns the untouched np.ones_like(mask) in that case
    area = torch.bincount(labels.reshape(-1), minlength=k + 1)[1:].to(torch.int32).contiguous()
    out = torch.empty(m.shape, dtype=torch.int64, device=md.device)
    L.fcall("mcb_size_matrix", labels.data_ptr(), area.data_ptr(), out.data_ptr(), m.shape[0], m.shape[1])
    return out.cpu().numpy()


def target_batch(masks, distances, sizes, pad=(0, 0), pad_method="replicate"):
    """the (N, 3, H', W') float32 target of MetadataImageSegmentationDatasetDistances (src/loaders.py:141-171) from the
    prepared per-image arrays: masks (N, H, W) uint8 {0,1}, distances (N, H, W) float16, sizes (N, H, W) int64"""
    md = _to_dev(np.asarray(masks).astype(np.uint8), torch.uint8)
    dd = _to_dev(np.asarray(distances).astype(np.float16), torch.float16)
    sd = _to_dev(np.asarray(sizes).astype(np.int64), torch.int64)
    n, h, w = md.shape
    ph, pw = int(pad[0]), int(pad[1])
    out = torch.empty((n, 3, h + 2 * ph, w + 2 * pw), dtype=torch.float32, device=md.device)
    L.fcall("mcb_target_channels", md.data_ptr(), dd.data_ptr(), sd.data_ptr(), out.data_ptr(), n, h, w, ph, pw,
            PAD_MODES[pad_method])
    return out
