"""CLIPSeg-ensemble fusion step: oracle vs the reference-generated golden file (CPU), CUDA vs oracle (GPU)."""
import os

import numpy as np
import pytest

from oracle import ensemble_oracle as EO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ensemble_alpha.npz")


def test_oracle_matches_reference_search_best_alpha():
    """tests/golden/ensemble_alpha.npz was produced by executing the reference's own search_best_alpha / ConfusionMatrix source
    (oracle/gen_golden_ensemble.py): same best alpha, per-alpha mIoU within the reference's 4-decimal print, same final masks."""
    g = np.load(GOLD)
    clip, unet, labels = EO.make_case(20240, EO.CASE_SIZES)
    best, best_miou, mious, conf = EO.search_best_alpha(clip, unet, labels)
    assert abs(best - float(g["best_alpha"])) < 1e-12
    assert np.abs(mious - g["mious"]).max() < 6e-5
    assert np.array_equal(conf, g["confusion"])
    for i, (c, u, lab) in enumerate(zip(clip, unet, labels)):
        shape = tuple(int(v) for v in g[f"final{i}_shape"])
        ref = np.unpackbits(g[f"final{i}"])[: shape[0] * shape[1]].reshape(shape)
        mine = EO.nearest_resize(EO.fuse_predict(EO.bilinear_resize(c, u.shape[1], u.shape[2]), u, best), *shape)
        assert (ref != mine).mean() < 1e-3


def test_nearest_index_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    for n_src, n_dst in [(96, 96), (100, 150), (64, 47), (75, 113), (353, 480), (7, 1)]:
        src = np.arange(n_src, dtype=np.uint8)[None, :].repeat(2, 0)
        got = cv2.resize(src, (n_dst, 2), interpolation=cv2.INTER_NEAREST)[0]
        assert np.array_equal(got, EO.nearest_index(n_src, n_dst).astype(np.uint8))


def test_empty_and_ignored_labels_oracle():
    clip, unet, _ = EO.make_case(3, [((16, 20), (16, 20))])
    lab = np.full((16, 20), 255, dtype=np.uint8)             # every pixel ignored -> empty confusion -> mIoU 0 -> alpha stays 0.0
    best, miou, _, conf = EO.search_best_alpha(clip, unet, [lab], search_step=5)
    assert best == 0.0 and miou == 0.0 and conf.sum() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("pre_interpolated", [False, True])
def test_cuda_alpha_sweep_bit_exact_vs_oracle(pre_interpolated):
    from egm_unet_b200 import ensemble as ENS
    clip, unet, labels = EO.make_case(20240, EO.CASE_SIZES)
    if pre_interpolated:     # the reference flow: clip logits already at the UNet size (eval_CLIPseg.py:885-888)
        clip = [EO.bilinear_resize(c, u.shape[1], u.shape[2]) for c, u in zip(clip, unet)]
    alphas = np.linspace(0.1, 10.0, 100)
    conf, miou, best = ENS.alpha_sweep([c[None] for c in clip], [u[None] for u in unet], labels, alphas)
    o_best, o_miou, o_mious, o_conf = EO.search_best_alpha(clip, unet, labels)
    assert np.array_equal(conf.cpu().numpy(), o_conf), "confusion counts must be bit-exact"
    assert np.abs(miou.cpu().numpy().astype(np.float64) - o_mious).max() < 1e-6
    assert float(best[0]) == o_best and abs(float(best[1]) - o_miou) < 1e-6
    g = np.load(GOLD)
    assert ENS.search_best_alpha([c[None] for c in clip], [u[None] for u in unet], labels) == float(g["best_alpha"])


@pytest.mark.gpu
def test_cuda_fuse_predict_and_edge_cases():
    from egm_unet_b200 import ensemble as ENS
    clip, unet, labels = EO.make_case(7, [((40, 56), (61, 83))])
    for alpha, size in [(0.3, (83, 61)), (2.5, (56, 40)), (10.0, (17, 9))]:          # PIL size = (width, height)
        got = ENS.fuse_predict(clip[0][None], unet[0][None], alpha, size)
        want = EO.nearest_resize(EO.fuse_predict(EO.bilinear_resize(clip[0], 40, 56), unet[0], alpha), size[1], size[0])
        assert got.shape == want.shape and np.array_equal(got, want)
    # ignored labels only -> best alpha stays at the reference's initial 0.0
    lab = np.full((61, 83), 255, dtype=np.uint8)
    conf, miou, best = ENS.alpha_sweep([clip[0][None]], [unet[0][None]], [lab], np.linspace(0.1, 10.0, 7))
    assert int(conf.sum()) == 0 and float(best[0]) == 0.0 and float(best[1]) == 0.0
