"""CPU restatement of the CLIPSeg-ensemble fusion step that consumes the UNet logits.

TEST INFRASTRUCTURE ONLY: imported by tests/ and oracle/gen_golden_ensemble.py, never by the product path.

Follows (reference file:line):
  * eval_CLIPseg.py:885-888 / predict_CLIPseg.py:500-503  -- F.interpolate(clip_logits, size=unet.shape[2:], bilinear, align_corners=False)
  * eval_CLIPseg.py:656-724  search_best_alpha            -- alpha grid np.linspace(0.1, 10, 100); fused = clip + alpha*unet; argmax;
                                                            cv2 INTER_NEAREST resize to the label size; ONE confusion matrix over all images per alpha;
                                                            first alpha with strictly larger mIoU wins (best starts at alpha 0.0 / mIoU 0.0)
  * eval_CLIPseg.py:726-748  ConfusionMatrix              -- mat[label][pred]; IoU = diag / (row + col - diag) with 0-denominators -> 1; mean
  * eval_CLIPseg.py:901-912 / predict_CLIPseg.py:519-526  -- final mask = uint8(argmax(clip + best_alpha*unet)) resized INTER_NEAREST to the image size
Pinned by oracle/gen_golden_ensemble.py, which executes the reference's own search_best_alpha / ConfusionMatrix source on the same inputs.
"""
import numpy as np


def bilinear_resize(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """x [C,h,w] float32 -> [C,out_h,out_w]; ATen upsample_bilinear2d, align_corners=False, evaluated in float32."""
    c, h, w = x.shape
    f32 = np.float32

    def axis(n_in, n_out):
        scale = f32(n_in) / f32(n_out)
        src = np.maximum(scale * (np.arange(n_out, dtype=f32) + f32(0.5)) - f32(0.5), f32(0.0)).astype(f32)
        i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
        i1 = np.minimum(i0 + 1, n_in - 1)
        l1 = (src - i0.astype(f32)).astype(f32)
        return i0, i1, (f32(1.0) - l1).astype(f32), l1

    h0, h1, lh0, lh1 = axis(h, out_h)
    w0, w1, lw0, lw1 = axis(w, out_w)
    top = x[:, h0][:, :, w0] * lw0 + x[:, h0][:, :, w1] * lw1
    bot = x[:, h1][:, :, w0] * lw0 + x[:, h1][:, :, w1] * lw1
    return (top * lh0[None, :, None] + bot * lh1[None, :, None]).astype(f32)


def nearest_index(n_src: int, n_dst: int) -> np.ndarray:
    """cv2.resize(..., INTER_NEAREST) source index per destination index (resizeNN: min(floor(x * (1/fx)), n_src-1), fx = n_dst/n_src in double)."""
    fx = float(n_dst) / float(n_src)
    ifx = 1.0 / fx
    return np.minimum(np.floor(np.arange(n_dst, dtype=np.float64) * ifx).astype(np.int64), n_src - 1)


def nearest_resize(a: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    return a[nearest_index(a.shape[0], out_h)][:, nearest_index(a.shape[1], out_w)]


def fuse_predict(clip_up: np.ndarray, unet: np.ndarray, alpha) -> np.ndarray:
    """argmax_c(clip_up + float32(alpha) * unet) -> uint8 [H,W]   (ties -> lowest class, like torch.argmax)"""
    fused = clip_up + (np.float32(alpha) * unet).astype(np.float32)
    return np.argmax(fused, axis=0).astype(np.uint8)


def miou_from_confusion(mat: np.ndarray) -> float:
    h = mat.astype(np.float32)
    denom = h.sum(1) + h.sum(0) - np.diag(h)
    safe = np.where(denom == 0, np.float32(1.0), denom)
    return float((np.diag(h) / safe).astype(np.float32).mean(dtype=np.float32))


def confusion_per_alpha(clip_list, unet_list, labels, alphas, num_classes=2) -> np.ndarray:
    """[n_alpha, C, C] int64; clip_list[i] [C,hc,wc], unet_list[i] [C,H,W] float32, labels[i] [Hl,Wl] integer."""
    out = np.zeros((len(alphas), num_classes, num_classes), dtype=np.int64)
    ups = [bilinear_resize(c, u.shape[1], u.shape[2]) for c, u in zip(clip_list, unet_list)]
    for ai, alpha in enumerate(alphas):
        for up, u, lab in zip(ups, unet_list, labels):
            pred = fuse_predict(up, u, alpha)
            if pred.shape != lab.shape:
                pred = nearest_resize(pred, lab.shape[0], lab.shape[1])
            a = lab.reshape(-1).astype(np.int64)
            b = pred.reshape(-1).astype(np.int64)
            k = (a >= 0) & (a < num_classes)
            out[ai] += np.bincount(num_classes * a[k] + b[k], minlength=num_classes ** 2).reshape(num_classes, num_classes)
    return out


def search_best_alpha(clip_list, unet_list, labels, search_scale=(0.1, 10.0), search_step=100, num_classes=2):
    alphas = np.linspace(search_scale[0], search_scale[1], search_step)
    conf = confusion_per_alpha(clip_list, unet_list, labels, alphas, num_classes)
    best_alpha, best_miou, mious = 0.0, 0.0, []
    for alpha, m in zip(alphas, conf):
        miou = miou_from_confusion(m)
        mious.append(miou)
        if miou > best_miou:
            best_miou, best_alpha = miou, float(alpha)
    return best_alpha, best_miou, np.asarray(mious, dtype=np.float64), conf


def make_case(seed: int, sizes):
    """Synthetic validation set with an interior optimum of alpha: CLIP logits at 352^2 carry the truth plus smooth (blob) errors,
    UNet logits carry a weaker truth signal plus per-pixel noise; labels at the given ((H,W),(Hl,Wl)) sizes."""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    clip, unet, labels = [], [], []
    for (h, w), (hl, wl) in sizes:
        truth = bilinear_resize(rng.standard_normal((1, 9, 9)).astype(f32), hl, wl)[0] > 0.1
        t352 = nearest_resize(truth.astype(np.uint8), 352, 352).astype(f32) - f32(0.5)
        blob = bilinear_resize(rng.standard_normal((1, 24, 24)).astype(f32), 352, 352)[0]
        dc = t352 * f32(1.0) + blob * f32(0.9)
        c = np.stack([-dc, dc]).astype(f32) * f32(0.5) + rng.standard_normal((2, 352, 352)).astype(f32) * f32(0.05)
        tl = nearest_resize(truth.astype(np.uint8), h, w).astype(f32) - f32(0.5)
        du = tl * f32(0.2) + rng.standard_normal((h, w)).astype(f32) * f32(0.5)
        u = np.stack([-du, du]).astype(f32) * f32(0.5) + rng.standard_normal((2, h, w)).astype(f32) * f32(0.02)
        clip.append(c.astype(f32)); unet.append(u.astype(f32)); labels.append(truth.astype(np.uint8))
    return clip, unet, labels


CASE_SIZES = [((96, 128), (96, 128)), ((100, 75), (150, 113)), ((64, 64), (47, 90))]
