"""dense CRF and watershed kernels against their CPU oracles.  PARITY UNPINNED for both (pydensecrf is absent from the
reference tree and this image; watershed is not a reference function): the oracle restates / defines the semantics, the
tests check the CUDA path against it — float tolerance 1e-4 for the CRF probabilities (float32 exp/sum order),
bit-exact int32 labels for the watershed."""
import numpy as np
import pytest
import torch

from oracle import post_oracle as P
from oracle import synthetic

pytestmark = pytest.mark.gpu


def test_dense_crf_matches_oracle(mcb, cuda):
    from mcb200 import postprocessing as G
    rs = np.random.RandomState(0)
    probs = synthetic.probability_maps(2, 72, seed=4, n_rect=9)
    imgs = rs.randn(2, 3, 72, 72).astype(np.float32)
    # make the image correlate with the mask so the bilateral term matters
    imgs += (probs[:, 1:2] > 0.5) * 1.5
    got = G.dense_crf_batch(torch.from_numpy(imgs).to(cuda), torch.from_numpy(probs).to(cuda)).cpu().numpy()
    for i in range(2):
        ref = P.dense_crf(imgs[i], probs[i])
        assert got[i].shape == ref.shape == (2, 72, 72)
        assert np.abs(got[i] - ref).max() < 1e-4, np.abs(got[i] - ref).max()
        assert np.allclose(got[i].sum(0), 1.0, atol=1e-5)
    one = G.dense_crf(imgs[0], probs[0], iterations=2)
    assert np.abs(one - P.dense_crf(imgs[0], probs[0], iterations=2)).max() < 1e-4
    # the CRF sharpens: fewer uncertain pixels than the input
    assert (np.abs(got[:, 1] - 0.5) < 0.25).sum() < (np.abs(probs[:, 1] - 0.5) < 0.25).sum()


def test_crf_rgb_conversion_matches_numpy_cast(mcb, cuda):
    from mcb200 import _lib as L
    rs = np.random.RandomState(1)
    img = (rs.randn(2, 3, 17, 19) * 1.2).astype(np.float32)
    rgb = torch.empty((2, 17, 19, 3), dtype=torch.uint8, device=cuda)
    L.fcall("mcb_crf_rgb_from_normalized", torch.from_numpy(img).to(cuda).data_ptr(), rgb.data_ptr(), 2, 17, 19)
    ref = np.stack([P.crf_rgb_image(x) for x in img])
    assert np.array_equal(rgb.cpu().numpy(), ref)


@pytest.mark.parametrize("h,w,seed", [(64, 64, 0), (96, 70, 1), (130, 45, 2)])
def test_watershed_bit_exact_against_oracle(mcb, cuda, h, w, seed):
    from mcb200 import postprocessing as G
    from scipy import ndimage as ndi
    rs = np.random.RandomState(seed)
    planes = []
    for _ in range(3):
        z = ndi.gaussian_filter(rs.randn(h, w), 3.0)
        z = (z - z.min()) / (z.max() - z.min())
        planes.append(z.astype(np.float32))
    prob = np.stack(planes)
    markers = np.stack([P.label(p > 0.75) for p in prob]).astype(np.int32)
    mask = prob > 0.35
    got = G.watershed_batch(torch.from_numpy(prob).to(cuda), torch.from_numpy(markers).to(cuda),
                            torch.from_numpy(mask).to(cuda)).cpu().numpy()
    for i in range(3):
        ref = P.minimax_watershed(prob[i], markers[i], mask[i])
        assert np.array_equal(got[i], ref), (i, (got[i] != ref).sum())
        # properties: markers keep their label, nothing outside the mask, every label is a marker label
        assert np.array_equal(got[i][markers[i] > 0], markers[i][markers[i] > 0])
        assert (got[i][~(mask[i] | (markers[i] > 0))] == 0).all()
    split = G.watershed_split(torch.from_numpy(prob).to(cuda), hi=0.75, lo=0.35).cpu().numpy()
    assert np.array_equal(split, got)


def test_watershed_plateaus_and_empty(mcb, cuda):
    from mcb200 import postprocessing as G
    prob = np.full((2, 40, 40), 0.9, np.float32)  # one flat plateau: ties decided by geodesic distance, then label
    markers = np.zeros((2, 40, 40), np.int32)
    markers[0, 5, 5] = 2
    markers[0, 30, 33] = 1
    mask = np.ones((2, 40, 40), bool)
    got = G.watershed_batch(torch.from_numpy(prob).to(cuda), torch.from_numpy(markers).to(cuda),
                            torch.from_numpy(mask).to(cuda)).cpu().numpy()
    assert np.array_equal(got[0], P.minimax_watershed(prob[0], markers[0], mask[0]))
    assert (got[1] == 0).all()  # no markers -> nothing is labelled
