"""Host-side pieces of the product that need no GPU: the COCO string / bbox formatters of mcb200.utils against the
maskApi.c restatement (and its known-answer vectors), and Pillow's fixed-point bilinear filter tables of
mcb200.preparation against Pillow itself (a numpy resample with those tables must equal Image.resize bit for bit --
the CUDA kernel applies the same tables, tests/test_input_gpu.py)."""
import numpy as np

import mcb200  # noqa: F401
from mcb200 import preparation as prep
from mcb200 import utils as U
from oracle import instances_oracle as I


def test_coco_string_and_bbox_formatters():
    assert U.rle_counts_to_string([1, 3]) == b"13" and U.rle_counts_to_string([0, 1, 3]) == b"013"
    assert U.rle_counts_to_string([40, 2, 5, 30]) == b"X125l0" and U.rle_counts_to_string([5, 10, 7, 3]) == b"5:7I"
    assert U.rle_to_bbox([1, 3], 2, 2) == [0.0, 0.0, 2.0, 2.0] and U.rle_to_bbox([6], 3, 2) == [0.0] * 4
    rs = np.random.RandomState(0)
    for _ in range(200):
        c = [int(v) for v in rs.randint(0, 200000, rs.randint(1, 40))]
        s = I.rle_to_string(c)
        assert U.rle_counts_to_string(c) == s and U.rle_string_to_counts(s) == c
        h = int(rs.randint(2, 50))
        assert U.rle_to_bbox(c, h, 10 ** 6) == I.rle_to_bbox(c, h, 10 ** 6)
    assert U.bounding_box_from_rle({"size": [2, 2], "counts": b"13"}) == [0.0, 0.0, 2.0, 2.0]


def test_pillow_filter_tables_reproduce_pillow():
    from PIL import Image
    rs = np.random.RandomState(1)
    for h, w, oh, ow in [(300, 300, 256, 256), (37, 53, 20, 31), (50, 40, 80, 90), (64, 64, 64, 64)]:
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        ch, bh = prep.pil_bilinear_coeffs(w, ow)
        cv, bv = prep.pil_bilinear_coeffs(h, oh)
        tmp = np.zeros((h, ow, 3), np.uint8)
        for xx in range(ow):
            x0, t = bh[xx]
            acc = (1 << 21) + (img[:, x0:x0 + t].astype(np.int64) * ch[xx, :t][None, :, None]).sum(1)
            tmp[:, xx] = np.clip(acc >> 22, 0, 255)
        out = np.zeros((oh, ow, 3), np.uint8)
        for yy in range(oh):
            y0, t = bv[yy]
            acc = (1 << 21) + (tmp[y0:y0 + t].astype(np.int64) * cv[yy, :t][:, None, None]).sum(0)
            out[yy] = np.clip(acc >> 22, 0, 255)
        assert np.array_equal(out, np.array(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))), (h, w, oh, ow)


def test_tta_specs_and_codes():
    from mcb200 import loaders as lo
    specs = lo.tta_specs()
    assert specs == I.tta_specs() and len(specs) == 16 and specs[0]["rotation"] == 0
    # `if ud ... elif lr`: both flips set -> up-down only (src/loaders.py:471-474)
    assert lo.spec_code({"ud_flip": True, "lr_flip": True, "rotation": 90, "color_shift": False}) == (1 | (1 << 2))
    assert lo.spec_code({"ud_flip": False, "lr_flip": True, "rotation": 270, "color_shift": False}) == (3 | (2 << 2))
    gen = lo.TestTimeAugmentationGenerator(flip_ud=True, flip_lr=False, rotation=False, color_shift_runs=False)
    out = gen.transform([["a"], ["b"], ["c"]])
    assert out["img_ids"] == [0, 0, 1, 1, 2, 2] and len(out["tta_params"]) == 6
