"""Pin oracle/pipeline_oracle.py against the REFERENCE's transforms.py (executed on PIL images with Pillow + torchvision in this
container) and write tests/golden/pipeline.npz: SHA-256 digests + spot values of the reference's output tensors for seeded
synthetic images.   python oracle/gen_golden_pipeline.py      (authoring container only: needs /root/reference)"""
import hashlib
import importlib.util
import os
import random
import sys

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pipeline_oracle as PO  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_transforms", "/root/reference/transforms.py")
T = importlib.util.module_from_spec(spec)
spec.loader.exec_module(T)


class PresetTrain:      # train.py:17-36
    def __init__(self, base_size, crop_size, hflip_prob=0.5, vflip_prob=0.5, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
        trans = [T.RandomResize(int(0.5 * base_size), int(1.2 * base_size))]
        if hflip_prob > 0:
            trans.append(T.RandomHorizontalFlip(hflip_prob))
        if vflip_prob > 0:
            trans.append(T.RandomVerticalFlip(vflip_prob))
        trans.extend([T.RandomCrop(crop_size), T.ToTensor(), T.Normalize(mean=mean, std=std)])
        self.transforms = T.Compose(trans)

    def __call__(self, img, target):
        return self.transforms(img, target)


class PresetEval:       # train.py:39-48
    def __init__(self, base_size, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
        self.transforms = T.Compose([T.RandomResize(base_size, base_size), T.ToTensor(), T.Normalize(mean=mean, std=std)])

    def __call__(self, img, target):
        return self.transforms(img, target)


def cat_list(images, fill_value=0):     # my_dataset.py:127-133
    max_size = tuple(max(s) for s in zip(*[img.shape for img in images]))
    batch_shape = (len(images),) + max_size
    batched_imgs = images[0].new(*batch_shape).fill_(fill_value)
    for img, pad_img in zip(images, batched_imgs):
        pad_img[..., :img.shape[-2], :img.shape[-1]].copy_(img)
    return batched_imgs


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out = {}
    for ci, (h, w, train) in enumerate(PO.GOLDEN_CASES):
        img, mask = PO.synth_image(h, w, 40 + ci)
        random.seed(900 + ci); torch.manual_seed(900 + ci)
        pil_i = Image.fromarray(img).convert("RGB")
        tgt = np.array(Image.fromarray(mask).convert("L")) / 255            # my_dataset.py:105-108
        pil_m = Image.fromarray(np.clip(tgt, a_min=0, a_max=255))
        xr, tr = (PresetTrain(565, 480) if train else PresetEval(565))(pil_i, pil_m)
        random.seed(900 + ci); torch.manual_seed(900 + ci)
        p = PO.draw_params(h, w, train)
        xo, to = PO.transform(img, mask, p)
        assert np.array_equal(xr.numpy(), xo) and np.array_equal(tr.numpy(), to), (ci, h, w, train, p)
        out[f"img{ci}"] = np.frombuffer(bytes.fromhex(digest(xr.numpy())), dtype=np.uint8)
        out[f"tgt{ci}"] = np.frombuffer(bytes.fromhex(digest(tr.numpy())), dtype=np.uint8)
        out[f"spot{ci}"] = xr.numpy()[:, ::97, ::89].copy()
        print(ci, (h, w), "train" if train else "eval", p, tuple(xr.shape), "ok")
    # collate_fn on an eval batch of different sizes (my_dataset.py:118-133)
    items = []
    for ci, (h, w) in enumerate([(300, 260), (200, 280)]):
        img, mask = PO.synth_image(h, w, 70 + ci)
        pil_i = Image.fromarray(img).convert("RGB")
        pil_m = Image.fromarray(np.clip(np.array(Image.fromarray(mask).convert("L")) / 255, a_min=0, a_max=255))
        items.append(PresetEval(120)(pil_i, pil_m))
    bi, bt = cat_list([i[0] for i in items], 0), cat_list([i[1] for i in items], 255)
    oi, ot = PO.collate([PO.transform(*PO.synth_image(h, w, 70 + ci), dict(size=120, hflip=False, vflip=False, crop=None))
                         for ci, (h, w) in enumerate([(300, 260), (200, 280)])])
    assert np.array_equal(bi.numpy(), oi) and np.array_equal(bt.numpy(), ot)
    out["collate_img"] = np.frombuffer(bytes.fromhex(digest(bi.numpy())), dtype=np.uint8)
    out["collate_tgt"] = np.frombuffer(bytes.fromhex(digest(bt.numpy())), dtype=np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "pipeline.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
