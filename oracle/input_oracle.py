"""TEST INFRASTRUCTURE — CPU restatement of the input side; only tests/ import it.
Pinned in tests/test_input_cpu.py against the real libraries the reference calls (cv2.copyMakeBorder, torchvision
ToTensor / Normalize, scipy.ndimage.distance_transform_edt) and, in the build container, against the reference's own
src/preparation.py functions through oracle/ref_shim.py."""
import numpy as np
from scipy import ndimage as ndi

MEAN = [0.485, 0.456, 0.406]
STD = [0.229, 0.224, 0.225]


def pad_image(img, pad, pad_method="replicate"):
    """PadFixed._pad (src/augmentation.py:68-80) without cv2: np.pad 'edge' / 'reflect' == BORDER_REPLICATE / REFLECT_101"""
    ph, pw = pad
    mode = {"replicate": "edge", "reflect": "reflect"}[pad_method]
    width = ((ph, ph), (pw, pw)) + ((0, 0),) * (img.ndim - 2)
    return np.pad(img, width, mode=mode)


def image_transform(img, pad=(0, 0), pad_method="replicate"):
    """(H, W, 3) uint8 -> (3, H', W') float32: pad, ToTensor (float32 / 255), Normalize (fp32 sub, fp32 div)"""
    x = pad_image(img, pad, pad_method).transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    m = np.asarray(MEAN, np.float32)[:, None, None]
    s = np.asarray(STD, np.float32)[:, None, None]
    return (x - m) / s


def update_distances(dist, mask):
    """src/preparation.py:151-156"""
    if dist.sum() == 0:
        return ndi.distance_transform_edt(1 - mask)
    return np.dstack([dist, ndi.distance_transform_edt(1 - mask)])


def clean_distances(distances):
    """src/preparation.py:159-168"""
    if len(distances.shape) < 3:
        distances = np.dstack([distances, distances])
    else:
        distances.sort(axis=2)
        distances = distances[:, :, :2]
    second = distances[:, :, 1]
    return np.sum(distances, axis=2).astype(np.float16), second


def two_nearest_distances(instance_masks):
    d = np.zeros(instance_masks.shape[1:])
    for m in instance_masks:
        d = update_distances(d, (m != 0).astype(np.uint8))
    return clean_distances(d)


def get_size_matrix(mask):
    """src/preparation.py:189-195"""
    sizes = np.ones_like(mask)
    labeled, _ = ndi.label(mask)
    for l in range(1, labeled.max() + 1):
        sizes = np.where(labeled == l, (labeled == l).sum(), sizes)
    return sizes


def target(mask, distances, sizes, pad=(0, 0), pad_method="replicate"):
    """src/loaders.py:141-171 (deterministic part): -> (3, H', W') float32"""
    d = distances.astype(np.uint16)
    s = np.sqrt(sizes.astype(np.uint16)).astype(np.uint16)
    chans = [pad_image(c.astype(np.uint8), pad, pad_method).astype(np.float32) for c in (mask, d, s)]
    return np.stack(chans)


def pil_resize(img, size):
    """transforms.Resize(size) on a PIL image: the REAL Pillow resampler (Pillow is installed wherever the tests run)"""
    from PIL import Image
    return np.array(Image.fromarray(img).resize((int(size[1]), int(size[0])), Image.BILINEAR))


def image_transform_resize(img, size):
    return image_transform(pil_resize(img, size))
