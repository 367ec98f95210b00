"""Parity of the CUDA post-processing (csrc/postproc.cu through mcb200.postprocessing) against the CPU oracle and the
golden fixtures produced by the unmodified reference.  Bit-exact for bool / uint8 / int32 outputs and for the float64
resize; scores (float64 sums in a different order) to 1e-9 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import post_oracle as P
from oracle import synthetic

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def G(mcb, cuda):
    from mcb200 import postprocessing
    return postprocessing


def test_golden_chain(G):
    g = np.load(os.path.join(GOLD, "postproc.npz"))
    for i, p in enumerate(g["probs"]):
        r = G.resize_image(p, (75, 75))
        assert r.dtype == np.float64 and np.array_equal(r, g["resize_%d" % i])
        c = G.categorize_multilayer_image(r)
        assert c.dtype == bool and np.array_equal(c, g["cat_%d" % i])
        l = G.label_multilayer_image(c)
        assert l.dtype == np.int32 and np.array_equal(l, g["label_%d" % i])
        assert np.array_equal(G.dilate_image(l, 2), g["dilate2_%d" % i])
        assert np.array_equal(G.dilate_image(l, 3), g["dilate3_%d" % i])
        e = G.erode_image(c[1], 2)
        assert e.dtype == np.uint8 and np.array_equal(e, g["erode2_%d" % i])
        assert np.array_equal(G.erode_image(c[1], 3), g["erode3_%d" % i])
        _, s = G.build_score(g["dilate2_%d" % i], r)
        assert np.allclose(np.array([float(v) for v in s[0]]), g["score0_%d" % i], rtol=1e-9, atol=0)
        assert np.allclose(np.array([float(v) for v in s[1]]), g["score1_%d" % i], rtol=1e-9, atol=0)
        assert np.array_equal(G.crop_image_center_per_class(p, 56, 56), g["crop_%d" % i])
    sm = G.softmax(g["softmax_in"], axis=1)
    assert np.allclose(sm, g["softmax_out"], rtol=0, atol=2e-7)
    assert np.array_equal(G.label_multiclass_image(g["docstring_mask"]), g["docstring_labels"])


@pytest.mark.parametrize("hw_in,hw_out", [((256, 256), (300, 300)), ((64, 96), (75, 100)), ((320, 320), (300, 300)),
                                           ((33, 47), (51, 60))])
def test_resize_bit_exact(G, cuda, hw_in, hw_out):
    rs = np.random.RandomState(hw_in[0] * 7 + hw_out[1])
    x = rs.rand(3, 2, *hw_in).astype(np.float32)
    x[0, 0] = x[0, 0] * 2 - 1  # negative values exercise the clip bounds
    got = G.resize_batch(torch.from_numpy(x).to(cuda), hw_out).cpu().numpy()
    ref = np.stack([P.resize_image(im, hw_out) for im in x])
    assert np.array_equal(got, ref)


def _random_masks(rs, planes, h, w, density):
    return (rs.rand(planes, h, w) < density).astype(np.uint8)


@pytest.mark.parametrize("h,w,density", [(300, 300, 0.5), (300, 300, 0.62), (64, 64, 0.3), (1, 37, 0.5), (45, 1, 0.6),
                                          (17, 300, 0.9), (128, 128, 0.0), (128, 128, 1.0)])
def test_ccl_matches_scipy_numbering(G, cuda, h, w, density):
    rs = np.random.RandomState(h * 31 + w)
    m = _random_masks(rs, 5, h, w, density)
    lab, cnt = G.label_batch(torch.from_numpy(m).to(cuda), return_counts=True)
    ref = np.stack([P.label(x) for x in m])
    assert lab.dtype == torch.int32
    assert np.array_equal(lab.cpu().numpy(), ref)
    assert np.array_equal(cnt.cpu().numpy(), ref.reshape(5, -1).max(1))


def test_ccl_structured_shapes(G, cuda):
    """spirals / U-shapes whose arms merge late stress the first-pixel numbering rule"""
    h = w = 96
    m = np.zeros((3, h, w), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    m[0] = ((yy // 4 + xx // 4) % 2 == 0)
    m[1, 10:80, 10] = 1; m[1, 10:80, 70] = 1; m[1, 79, 10:71] = 1; m[1, 5:9, 30:40] = 1
    r = np.sqrt((yy - 48.0) ** 2 + (xx - 48.0) ** 2); th = np.arctan2(yy - 48.0, xx - 48.0)
    m[2] = (np.abs(((r - 3 * th) % 12) - 6) < 1.5)
    lab = G.label_batch(torch.from_numpy(m).to(cuda)).cpu().numpy()
    assert np.array_equal(lab, np.stack([P.label(x) for x in m]))


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7])
def test_morphology_and_dropped_objects(G, cuda, k):
    rs = np.random.RandomState(k)
    probs = synthetic.probability_maps(3, 96, seed=k, n_rect=14)
    m = (probs > 0.5).astype(np.uint8).reshape(6, 96, 96)
    lab = np.stack([P.label(x) for x in m])
    d = G.morph_batch(torch.from_numpy(lab).to(cuda), k, True).cpu().numpy()
    assert np.array_equal(d, np.stack([P.dilate_image(x, k) for x in lab]))
    e = G.erode_batch(torch.from_numpy(m).to(cuda), k).cpu().numpy()
    ref_e = np.stack([P.erode_image(x != 0, k) for x in m])
    assert e.dtype == np.uint8 and np.array_equal(e, ref_e)
    noise = _random_masks(rs, 4, 50, 70, 0.55)
    e2 = G.erode_batch(torch.from_numpy(noise).to(cuda), k).cpu().numpy()
    assert np.array_equal(e2, np.stack([P.erode_image(x != 0, k) for x in noise]))


def test_batched_transformer_matches_per_image_oracle(G, cuda):
    probs = synthetic.probability_maps(6, 128, seed=11, n_rect=20)
    pp = G.MaskPostprocessor((150, 150), "resize", erode_selem_size=0, dilate_selem_size=2)
    out = pp.transform(probs)["y_pred"]
    assert len(out) == 6
    for p, (labels, scores) in zip(probs, out):
        r = P.resize_image(p, (150, 150))
        l = P.dilate_image(P.label_multilayer_image(P.categorize_multilayer_image(r)), 2)
        _, s = P.build_score(l, r)
        assert labels.dtype == np.int32 and np.array_equal(labels, l)
        for a, b in zip(scores, s):
            assert len(a) == len(b)
            for u, v in zip(a, b):
                assert (u is np.ma.masked and v is np.ma.masked) or abs(float(u) - float(v)) <= 1e-9 * abs(float(v))
    # crop mode (unet_padded pipeline): 128 -> 120 centre crop, float32 probabilities
    pc = G.MaskPostprocessor((120, 120), "crop", 0, 0).transform(probs)["y_pred"]
    for p, (labels, scores) in zip(probs, pc):
        c = P.crop_image_center_per_class(p, 120, 120)
        l = P.label_multilayer_image(P.categorize_multilayer_image(c))
        assert np.array_equal(labels, l)


def test_full_size_properties(G, cuda):
    """batch 64 @ 256 -> 300 (BASELINE config 4): size-independent properties instead of a slow CPU run"""
    probs = torch.from_numpy(synthetic.probability_maps(64, 256, seed=2)).to(cuda)
    pp = G.MaskPostprocessor((300, 300), "resize", 0, 2)
    labels, scores, counts, pr = pp.run_device(probs)
    cnts = counts.cpu().numpy()
    assert labels.shape == (64, 2, 300, 300) and labels.dtype == torch.int32
    # labelling is idempotent on its own binarisation, labels are dense 1..K
    pre = G.label_batch(G.threshold_batch(pr))
    again = G.label_batch((pre > 0).to(torch.uint8))
    assert torch.equal(pre, again)
    k = pre.view(128, -1).max(dim=1).values
    assert torch.equal(k.to(torch.int32), torch.from_numpy(cnts.astype(np.int32)).to(cuda))
    # dilation never removes foreground and never invents labels
    assert bool(((labels > 0) | (pre == 0)).all())
    assert int(labels.max()) == int(pre.max())
    # one image against the oracle
    ref = P.dilate_image(P.label_multilayer_image(P.categorize_multilayer_image(P.resize_image(probs[17].cpu().numpy(), (300, 300)))), 2)
    assert np.array_equal(labels[17].cpu().numpy(), ref)
    assert all(bool(torch.isfinite(scores[i, :cnts[i]]).all()) for i in range(128))
    # the CUDA-graph replay of the same chain returns the same labels / counts, scores equal up to fp64 atomic order
    for _ in range(2):
        gl, gs, gc, _ = pp.run_device_graphed(probs)
    assert torch.equal(gl, labels) and torch.equal(gc, counts)
    for i in (0, 17, 127):
        assert torch.allclose(gs[i, :cnts[i]], scores[i, :cnts[i]], rtol=1e-12, atol=0)
