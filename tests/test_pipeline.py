"""Input pipeline: oracle vs digests of the reference's own transforms.py outputs (CPU), CUDA vs oracle, bit-exact (GPU)."""
import hashlib
import os
import random

import numpy as np
import pytest
import torch

from oracle import pipeline_oracle as PO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pipeline.npz")


def _digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def _case(ci):
    h, w, train = PO.GOLDEN_CASES[ci]
    img, mask = PO.synth_image(h, w, 40 + ci)
    random.seed(900 + ci); torch.manual_seed(900 + ci)
    return img, mask, train, h, w


@pytest.mark.parametrize("ci", range(len(PO.GOLDEN_CASES)))
def test_oracle_matches_reference_transforms(ci):
    """tests/golden/pipeline.npz holds SHA-256 digests of what /root/reference/transforms.py (PIL + torchvision) produced for these
    seeded images and RNG states (oracle/gen_golden_pipeline.py): the restatement must reproduce every byte."""
    g = np.load(GOLD)
    img, mask, train, h, w = _case(ci)
    x, t = PO.transform(img, mask, PO.draw_params(h, w, train))
    assert np.array_equal(_digest(x), g[f"img{ci}"]) and np.array_equal(_digest(t), g[f"tgt{ci}"])
    assert np.array_equal(x[:, ::97, ::89], g[f"spot{ci}"])


def test_oracle_collate_matches_reference():
    g = np.load(GOLD)
    items = [PO.transform(*PO.synth_image(h, w, 70 + ci), dict(size=120, hflip=False, vflip=False, crop=None)) for ci, (h, w) in enumerate([(300, 260), (200, 280)])]
    bi, bt = PO.collate(items)
    assert np.array_equal(_digest(bi), g["collate_img"]) and np.array_equal(_digest(bt), g["collate_tgt"])
    assert (bt == 255).any() and bi.shape == (2, 3, 138, 168)


def test_host_tables_match_oracle_tables():
    """egm_unet_b200.data builds the Pillow coefficient / index tables vectorised; they must equal the scalar restatement."""
    import egm_unet_b200  # noqa: F401
    from egm_unet_b200 import data as D
    for n_in, n_out in [(565, 601), (640, 502), (1024, 646), (260, 303), (400, 387), (584, 584 * 2), (700, 1187), (3, 1), (1, 5)]:
        a = PO.bilinear_coeffs(n_in, n_out)
        b = D._bilinear_tables(n_in, n_out)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), (n_in, n_out)
        assert np.array_equal(PO.nearest_table(n_in, n_out), D._nearest_table(n_in, n_out))
    random.seed(5); torch.manual_seed(5)
    p1 = PO.draw_params(584, 565, True)
    random.seed(5); torch.manual_seed(5)
    assert D.DevicePipeline.__new__(D.DevicePipeline).__class__ is D.DevicePipeline
    pipe = object.__new__(D.DevicePipeline)
    pipe.train, pipe.base_size, pipe.crop_size, pipe.hflip_prob, pipe.vflip_prob = True, 565, 480, 0.5, 0.5
    assert pipe.draw(584, 565) == p1


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(PO.GOLDEN_CASES)))
def test_cuda_pipeline_bit_exact(ci):
    from egm_unet_b200.data import DevicePipeline
    g = np.load(GOLD)
    img, mask, train, h, w = _case(ci)
    pipe = DevicePipeline(train=train)
    x, t = pipe([img], [mask])                      # draws with the same RNG state as the reference run
    x, t = x[0].cpu().numpy(), t[0].cpu().numpy()
    assert np.array_equal(_digest(x), g[f"img{ci}"]), "image bytes differ from the reference pipeline"
    assert np.array_equal(_digest(t), g[f"tgt{ci}"]), "target bytes differ from the reference pipeline"


@pytest.mark.gpu
def test_cuda_pipeline_collate_and_batch():
    from egm_unet_b200.data import DevicePipeline
    g = np.load(GOLD)
    pipe = DevicePipeline(train=False, base_size=120)
    imgs, masks = zip(*[PO.synth_image(h, w, 70 + ci) for ci, (h, w) in enumerate([(300, 260), (200, 280)])])
    bi, bt = pipe(list(imgs), list(masks))
    assert np.array_equal(_digest(bi.cpu().numpy()), g["collate_img"]) and np.array_equal(_digest(bt.cpu().numpy()), g["collate_tgt"])
    # a training batch: every entry equals the oracle on its own draw, and feeds the model input contract (float32 NCHW / int64)
    pipe = DevicePipeline(train=True)
    random.seed(3); torch.manual_seed(3)
    sizes = [(584, 565), (480, 640), (375, 500), (768, 1024)]
    data = [PO.synth_image(h, w, 200 + i) for i, (h, w) in enumerate(sizes)]
    params = [pipe.draw(h, w) for h, w in sizes]
    x, t = pipe([d[0] for d in data], [d[1] for d in data], params)
    assert x.shape == (4, 3, 480, 480) and x.dtype == torch.float32 and t.shape == (4, 480, 480) and t.dtype == torch.int64
    for k, (d, p) in enumerate(zip(data, params)):
        xo, to = PO.transform(d[0], d[1], p)
        assert np.array_equal(x[k].cpu().numpy(), xo) and np.array_equal(t[k].cpu().numpy(), to)
