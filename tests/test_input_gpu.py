"""GPU parity of the input side (csrc/input.cu through mcb200.preparation) against oracle/input_oracle.py: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import input_oracle as IO
from oracle import synthetic

pytestmark = pytest.mark.gpu


def test_image_pad_normalize_bit_exact(mcb, cuda):
    from mcb200 import preparation as prep
    rs = np.random.RandomState(0)
    imgs = rs.randint(0, 256, (3, 60, 52, 3)).astype(np.uint8)
    imgs[0, :2, :2] = 255
    imgs[1] = 0
    for pad, method in (((10, 10), "replicate"), ((0, 0), "replicate"), ((7, 3), "reflect")):
        got = prep.image_transform_batch(imgs, pad, method).cpu().numpy()
        for i in range(3):
            assert np.array_equal(got[i], IO.image_transform(imgs[i], pad, method)), (pad, method, i)
    big = rs.randint(0, 256, (2, 300, 300, 3)).astype(np.uint8)
    got = prep.image_transform_batch(big, (10, 10)).cpu().numpy()
    assert got.shape == (2, 3, 320, 320) and np.array_equal(got[1], IO.image_transform(big[1], (10, 10)))


@pytest.mark.parametrize("h,w,k", [(40, 52, 5), (64, 33, 1), (30, 30, 0), (97, 130, 12)])
def test_two_nearest_distances_bit_exact(mcb, cuda, h, w, k):
    from mcb200 import preparation as prep
    rs = np.random.RandomState(h + k)
    masks = np.zeros((k, h, w), np.uint8)
    for i in range(k):
        m, _ = synthetic.rectangles_mask(rs, h, w, n_rect=1, lo=3, hi=12)
        if i == 0:
            m[0, 0] = 1                       # a corner pixel: distances across the whole image
        masks[i] = m
    got_sum, got_second = prep.two_nearest_distances(masks)
    want_sum, want_second = IO.two_nearest_distances(masks) if k else IO.clean_distances(np.zeros((h, w)))
    assert got_sum.dtype == np.float16 and got_second.dtype == np.float64
    assert np.array_equal(got_second, want_second)
    assert np.array_equal(got_sum, want_sum)


def test_size_matrix_and_target_tensor_bit_exact(mcb, cuda):
    from mcb200 import preparation as prep
    rs = np.random.RandomState(3)
    mask, _ = synthetic.rectangles_mask(rs, 90, 120, n_rect=14, lo=4, hi=20)
    sizes = prep.get_size_matrix(mask)
    want = IO.get_size_matrix(mask)
    assert sizes.dtype == want.dtype and np.array_equal(sizes, want)
    assert np.array_equal(prep.get_size_matrix(np.zeros((9, 9), np.uint8)), np.ones((9, 9), np.uint8))
    inst = np.stack([(mask > 0) & (rs.rand(*mask.shape) > 0.0)]).astype(np.uint8)
    dsum, _ = prep.two_nearest_distances(inst)
    dsum = (dsum.astype(np.float32) * 3.7).astype(np.float16)          # spread beyond 255: the uint8 wrap matters
    big = sizes.copy()
    big[mask > 0] += 70000                                             # beyond uint16: the wrap matters
    for pad, method in (((10, 10), "replicate"), ((0, 0), "replicate")):
        got = prep.target_batch(mask[None], dsum[None], big[None], pad, method).cpu().numpy()[0]
        assert np.array_equal(got, IO.target(mask, dsum, big, pad, method)), pad


@pytest.mark.parametrize("h,w,oh,ow", [(300, 300, 256, 256), (37, 53, 20, 31), (64, 64, 64, 64), (50, 40, 80, 90)])
def test_pil_bilinear_resize_bit_exact(mcb, cuda, h, w, oh, ow):
    """the `resize` loader mode (neptune.yaml default): transforms.Resize on a PIL image, against Pillow itself"""
    from mcb200 import preparation as prep
    rs = np.random.RandomState(h + ow)
    imgs = rs.randint(0, 256, (2, h, w, 3)).astype(np.uint8)
    got = prep.pil_resize_batch(imgs, (oh, ow)).cpu().numpy()
    for i in range(2):
        assert np.array_equal(got[i], IO.pil_resize(imgs[i], (oh, ow)))
    full = prep.image_transform_resize_batch(imgs, (oh, ow)).cpu().numpy()
    assert np.array_equal(full[0], IO.image_transform_resize(imgs[0], (oh, ow)))
    mask = (rs.rand(1, h, w, 1) > 0.6).astype(np.uint8)                       # single-band targets resize the same way
    got1 = prep.pil_resize_batch(mask, (oh, ow)).cpu().numpy()[0, :, :, 0]
    from PIL import Image
    assert np.array_equal(got1, np.array(Image.fromarray(mask[0, :, :, 0]).resize((ow, oh), Image.BILINEAR)))
