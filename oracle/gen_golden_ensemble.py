"""Pin oracle/ensemble_oracle.py against the REFERENCE's own code and write tests/golden/ensemble_alpha.npz.

Run in the authoring container only (needs /root/reference, cv2, torch):   python oracle/gen_golden_ensemble.py
The reference module eval_CLIPseg.py cannot be imported (it pulls in the CLIPSeg / CLIP checkpoints' packages at import time), so
the LIVE definitions of `search_best_alpha` and `ConfusionMatrix` (eval_CLIPseg.py:656-748) are cut out of the file with `ast`
and executed unmodified; F.interpolate is called exactly as eval_CLIPseg.py:885-888 does.
"""
import ast
import io
import os
import re
import sys
from contextlib import redirect_stdout

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ensemble_oracle as EO  # noqa: E402

REF = "/root/reference/eval_CLIPseg.py"


def reference_namespace():
    src = open(REF, encoding="utf-8").read()
    tree = ast.parse(src)
    wanted = {}
    for node in tree.body:                       # later definitions override earlier ones, as at import time
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in ("search_best_alpha", "ConfusionMatrix"):
            wanted[node.name] = ast.get_source_segment(src, node)
    ns = {"np": np, "torch": torch, "cv2": cv2}
    for name in ("ConfusionMatrix", "search_best_alpha"):
        exec(compile(wanted[name], REF + ":" + name, "exec"), ns)
    return ns


def main():
    ns = reference_namespace()
    clip, unet, labels = EO.make_case(20240, EO.CASE_SIZES)
    clip_t = [torch.from_numpy(c)[None] for c in clip]
    unet_t = [torch.from_numpy(u)[None] for u in unet]
    # eval_CLIPseg.py:885-888
    clip_up = [torch.nn.functional.interpolate(c, size=u.shape[2:], mode="bilinear", align_corners=False) for c, u in zip(clip_t, unet_t)]
    ups_ref = [c[0].numpy() for c in clip_up]
    ups_mine = [EO.bilinear_resize(c, u.shape[1], u.shape[2]) for c, u in zip(clip, unet)]
    max_up = max(float(np.abs(a - b).max()) for a, b in zip(ups_ref, ups_mine))
    print("bilinear: max |oracle - F.interpolate| =", max_up)
    assert max_up < 6e-6          # fp32 rounding (ATen fuses the multiply-adds differently), values are O(5)
    buf = io.StringIO()
    with redirect_stdout(buf):
        best_ref = ns["search_best_alpha"](clip_up, unet_t, labels)
    mious_ref = np.array([float(m) for m in re.findall(r"mIoU=([0-9.]+)", buf.getvalue())][:100])
    best, best_miou, mious, conf = EO.search_best_alpha(clip, unet, labels)
    print("reference best alpha", best_ref, "oracle", best, "mIoU", best_miou)
    assert abs(best - best_ref) < 1e-12, (best, best_ref)
    assert np.abs(mious - mious_ref).max() < 6e-5, np.abs(mious - mious_ref).max()          # the reference prints 4 decimals
    # final masks (eval_CLIPseg.py:901-912): argmax of the fused logits, uint8, cv2 INTER_NEAREST to the "original" size
    finals = []
    for up, u, lab in zip(clip_up, unet_t, labels):
        fused = up + best_ref * u
        pred = torch.argmax(fused, dim=1).squeeze(0).cpu().numpy().astype(np.uint8)
        pred = cv2.resize(pred, (lab.shape[1] + 7, lab.shape[0] + 5), interpolation=cv2.INTER_NEAREST)
        finals.append(pred)
    mine = [EO.nearest_resize(EO.fuse_predict(m, u, best), lab.shape[0] + 5, lab.shape[1] + 7) for m, u, lab in zip(ups_mine, unet, labels)]
    for a, b in zip(finals, mine):
        assert a.shape == b.shape and (a != b).mean() < 1e-3, (a != b).mean()
    out = os.path.join(ROOT, "tests", "golden", "ensemble_alpha.npz")
    np.savez_compressed(out, best_alpha=np.float64(best_ref), mious=mious_ref, confusion=conf,
                        **{f"final{i}": np.packbits(f) for i, f in enumerate(finals)}, **{f"final{i}_shape": np.array(f.shape) for i, f in enumerate(finals)})
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
