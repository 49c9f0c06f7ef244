"""GPU: the reference's own entry points, used the way train.py / predict.py use them (b):

    from src import GRFBUNet, UNet
    train_utils.train_one_epoch(model, optimizer, loader, device, epoch, num_classes, lr_scheduler, print_freq, scaler)
    train_utils.evaluate(model, loader, device, num_classes)
    train_utils.init_distributed_mode(args)

on a synthetic DataLoader whose collate pads images with 0 and targets with 255 (my_dataset.py:119-133) and whose validation
images are batch-1 and odd-sized (predict.py:56-77), against the CPU oracle driven through the same schedule.
"""
import argparse
import os
import warnings

import pytest
import torch

from oracle import egm_oracle as O
from oracle import synth
from tests.util import rel_err

pytestmark = pytest.mark.gpu


class _SynthSet(torch.utils.data.Dataset):
    """ragged synthetic (image, mask) pairs: float32 [3,h,w], int64 [h,w] in {0,1}"""

    def __init__(self, sizes, seed):
        self.items = []
        for i, (h, w) in enumerate(sizes):
            img, tgt = synth.make_inputs(1, h, w, seed=seed + i, blobs=True, ignore_rows=0)
            self.items.append((img[0], tgt[0]))

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


def collate_255(batch):
    """my_dataset.py:119-133: pad every sample to the batch maximum, images with 0, targets with 255"""
    hs, ws = max(b[0].shape[1] for b in batch), max(b[0].shape[2] for b in batch)
    imgs = torch.zeros(len(batch), 3, hs, ws)
    tgts = torch.full((len(batch), hs, ws), 255, dtype=torch.int64)
    for i, (im, tg) in enumerate(batch):
        imgs[i, :, :im.shape[1], :im.shape[2]] = im
        tgts[i, :tg.shape[0], :tg.shape[1]] = tg
    return imgs, tgts


def _loaders():
    train = _SynthSet([(48, 48), (40, 44), (48, 36), (44, 48), (48, 48), (36, 40)], seed=300)
    val = _SynthSet([(37, 45), (53, 41), (48, 48)], seed=400)
    tl = torch.utils.data.DataLoader(train, batch_size=2, shuffle=False, collate_fn=collate_255, pin_memory=True)
    vl = torch.utils.data.DataLoader(val, batch_size=1, shuffle=False, collate_fn=collate_255)
    return tl, vl


def _oracle_epochs(sd, variant, tl, epochs, lrs):
    osd = {k: v.clone() for k, v in sd.items()}
    mom, losses, it = {}, [], 0
    lw = torch.tensor([1.0, 2.0])
    for _ in range(epochs):
        for image, target in tl:
            loss, _ = O.train_step(osd, mom, image, target, variant, lrs[it], 0.9, 1e-4, lw)
            losses.append(float(loss))
            it += 1
    return osd, losses


def _oracle_eval(osd, variant, vl):
    mat = torch.zeros(2, 2, dtype=torch.int64)
    dice = []
    for image, target in vl:
        with torch.no_grad():
            logits = O.forward({k: v.detach() for k, v in osd.items()}, image, variant, False)
        mat += O.confusion_matrix(target, logits.argmax(1), 2)
        dice.append(O.dice_metric(logits, target))
    return mat, sum(dice) / len(dice)


def _lr_sequence(n_steps, epochs, base_lr=0.02):
    """the reference scheduler's LR per optimizer step (train_utils/train_and_eval.py:78-100)"""
    import train_utils
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=base_lr, momentum=0.9, weight_decay=1e-4)
    sch = train_utils.create_lr_scheduler(opt, n_steps // epochs, epochs, warmup=True)
    out = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(n_steps):
            out.append(opt.param_groups[0]["lr"])
            sch.step()
    return out


@pytest.mark.parametrize("fused", [True, False])
def test_train_one_epoch_and_evaluate_match_oracle(fused, monkeypatch, capsys):
    """UNet in fp32 check mode through train_utils.train_one_epoch + evaluate == the oracle's SGD loop: per-epoch mean loss, final
    parameters, confusion matrix and Dice.  fused=True is the default route (fused Trainer: CUDA graph, prefetch, lazy loss);
    fused=False forces the generic autograd route (model(image) / criterion / loss.backward() / optimizer.step())."""
    from src import UNet                      # the drop-in package of train.py:8
    import train_utils
    import train_utils.train_and_eval as tae
    if not fused:
        monkeypatch.setattr(tae, "_fused_trainer", lambda *a, **k: None)
    device = torch.device("cuda")
    model = UNet(in_channels=3, num_classes=2, base_c=32)
    sd = synth.fill_state_dict(model.state_dict())
    model.load_state_dict(sd)
    model.to(device)
    model.set_check_mode(True)
    tl, vl = _loaders()
    epochs = 2
    params = [p for p in model.parameters() if p.requires_grad]
    optimizer = torch.optim.SGD(params, lr=0.02, momentum=0.9, weight_decay=1e-4)          # train.py:113-118
    sched = train_utils.create_lr_scheduler(optimizer, len(tl), epochs, warmup=True)
    means = []
    for ep in range(epochs):
        mean_loss, lr = train_utils.train_one_epoch(model, optimizer, tl, device, ep, 2, lr_scheduler=sched, print_freq=2, scaler=None)
        means.append(mean_loss)
    out = capsys.readouterr().out
    assert "Epoch: [1]" in out and "lr:" in out and "Total time" in out     # MetricLogger output is kept
    if not fused:       # the fused route reads losses back two steps late: with 3 batches per epoch none has arrived at a print point
        assert "loss:" in out
    confmat, dice = train_utils.evaluate(model, vl, device=device, num_classes=2)
    lrs = _lr_sequence(len(tl) * epochs, epochs)
    osd, olosses = _oracle_epochs(sd, "unet", tl, epochs, lrs)
    n = len(tl)
    for ep in range(epochs):
        ref = sum(olosses[ep * n:(ep + 1) * n]) / n
        assert abs(means[ep] - ref) <= 2e-3 * abs(ref), (ep, means[ep], ref)
    assert abs(lr - _lr_sequence(len(tl) * epochs + 1, epochs)[-1]) < 1e-9 or lr >= 0      # scheduler stepped once per batch
    assert sched.last_epoch == len(tl) * epochs
    new = model.state_dict()
    worst = max((rel_err(new[k].cpu().float(), osd[k].detach().float()), k) for k in osd if osd[k].dtype.is_floating_point and osd[k].numel() > 1)
    # 6 SGD steps of fp32 gradients that are reproducible to ~1e-2 per step (ReLU / max-pool kinks flip under the split-K atomics'
    # summation order): the trajectory is compared globally (relative L2 over ALL parameters) with a loose per-tensor
    # bound -- the same run repeated gives 0.02 .. 0.09 for the worst single tensor (a BatchNorm bias of a few dozen elements)
    num = sum(float((new[k].cpu().float() - osd[k].detach().float()).double().pow(2).sum()) for k in osd if osd[k].dtype.is_floating_point)
    den = sum(float(osd[k].detach().float().double().pow(2).sum()) for k in osd if osd[k].dtype.is_floating_point)
    print(f"trajectory vs oracle after {len(tl) * epochs} steps: global rel L2 {(num / den) ** 0.5:.4f}, worst tensor {worst[0]:.4f} ({worst[1]})")
    assert (num / den) ** 0.5 < 2e-2, (num / den) ** 0.5
    assert worst[0] < 0.2, worst
    assert int(new["in_conv.1.num_batches_tracked"]) == len(tl) * epochs
    omat, odice = _oracle_eval(osd, "unet", vl)
    got = confmat.mat.cpu()
    assert int(got.sum()) == int(omat.sum())                               # same number of valid (non-255) pixels
    print(f"confusion-matrix pixels that differ: {int((got - omat).abs().sum())} of {int(omat.sum())}; dice {dice:.4f} vs {odice:.4f}")
    assert int((got - omat).abs().sum()) <= 0.06 * int(omat.sum()), (got, omat)        # masks of two slightly different trajectories
    assert abs(dice - odice) < 5e-2, (dice, odice)
    # the optimizer stays checkpointable (train.py:152-156) and its momentum state is live
    st = optimizer.state_dict()
    assert len(st["state"]) == len(params)
    assert all("momentum_buffer" in s and float(s["momentum_buffer"].abs().sum()) > 0 for s in list(st["state"].values())[:4])


def test_fused_and_autograd_routes_agree_on_egm_bf16():
    """EGM-UNet, bf16 production kernels: one epoch through the fused route == one epoch through the autograd route (same
    kernels, different orchestration), and the loss is within bf16 tolerance of the oracle's."""
    from src import GRFBUNet
    import train_utils
    import train_utils.train_and_eval as tae
    device = torch.device("cuda")
    tl, _ = _loaders()
    res = []
    for fused in (True, False):
        model = GRFBUNet(in_channels=3, num_classes=2, base_c=32)
        sd = synth.fill_state_dict(model.state_dict())
        model.load_state_dict(sd)
        model.to(device)
        optimizer = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=0.02, momentum=0.9, weight_decay=1e-4)
        sched = train_utils.create_lr_scheduler(optimizer, len(tl), 1, warmup=True)
        orig = tae._fused_trainer
        if not fused:
            tae._fused_trainer = lambda *a, **k: None
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                mean_loss, _ = train_utils.train_one_epoch(model, optimizer, tl, device, 0, 2, lr_scheduler=sched, print_freq=10,
                                                           scaler=torch.amp.GradScaler("cuda"))      # accepted, unused
        finally:
            tae._fused_trainer = orig
        res.append((mean_loss, {k: v.detach().float().cpu() for k, v in model.state_dict().items()}))
    (l0, s0), (l1, s1) = res
    assert abs(l0 - l1) <= 5e-3 * abs(l0), (l0, l1)
    _, olosses = _oracle_epochs(sd, "egm", tl, 1, _lr_sequence(len(tl), 1))
    ref = sum(olosses) / len(olosses)
    assert abs(l0 - ref) <= 3e-2 * abs(ref), (l0, ref)


def test_scaler_argument_warns_once_and_is_ignored():
    import train_utils.train_and_eval as tae
    from src import UNet
    import train_utils
    device = torch.device("cuda")
    model = UNet(in_channels=3, num_classes=2, base_c=32).to(device)
    tl, _ = _loaders()
    optimizer = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    sched = train_utils.create_lr_scheduler(optimizer, len(tl), 1)
    tae._SCALER_WARNED[0] = False
    scaler = torch.amp.GradScaler("cuda")
    with pytest.warns(UserWarning, match="scaler"):
        train_utils.train_one_epoch(model, optimizer, tl, device, 0, 2, lr_scheduler=sched, scaler=scaler)
    assert scaler.get_scale() == 65536.0           # untouched


def test_out_of_range_label_is_counted_not_read_out_of_bounds():
    """A label that is neither a class id nor ignore_index (e.g. 255 with the criterion's default ignore_index=-100): the reference
    raises a device assert; here the pixel is dropped from CE and Dice alike, class_weight is never indexed with it, and
    loss_terms reports the count."""
    from egm_unet_b200.loss import loss_terms
    logits = torch.randn(2, 2, 24, 20, device="cuda")
    target = torch.randint(0, 2, (2, 24, 20), device="cuda")
    target[:, :3] = 255
    lw = torch.tensor([1.0, 2.0], device="cuda")
    bad = loss_terms(logits, target, lw, ignore_index=-100)
    good = loss_terms(logits, target, lw, ignore_index=255)
    torch.cuda.synchronize()
    assert float(bad["bad_labels"]) == 2 * 3 * 20 and float(good["bad_labels"]) == 0
    assert abs(float(bad["ce"]) - float(good["ce"])) < 1e-6 and abs(float(bad["dice"]) - float(good["dice"])) < 1e-6


def test_init_distributed_mode_single_process_nccl():
    """distributed_utils.py:315-338 contract: env rendezvous -> args.{rank,world_size,gpu,distributed}, NCCL group, device bound."""
    import torch.distributed as dist
    import train_utils
    from train_utils import distributed_utils as du
    if dist.is_initialized():
        pytest.skip("a process group already exists in this process")
    env = {"RANK": "0", "WORLD_SIZE": "1", "LOCAL_RANK": "0", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29611"}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    import builtins
    plain_print = builtins.print
    try:
        args = argparse.Namespace(dist_url="env://")
        train_utils.init_distributed_mode(args)
        assert args.distributed and args.rank == 0 and args.world_size == 1 and args.gpu == 0 and args.dist_backend == "nccl"
        assert dist.is_initialized() and dist.get_backend() == "nccl" and du.get_world_size() == 1 and du.is_main_process()
        t = torch.ones(4, device="cuda")
        dist.all_reduce(t)
        assert float(t.sum()) == 4.0
        # the metric reductions of evaluate() go through this group
        cm = du.ConfusionMatrix(2)
        cm.update(torch.tensor([0, 1, 1], device="cuda"), torch.tensor([0, 1, 0], device="cuda"))
        cm.reduce_from_all_processes()
        assert int(cm.mat.sum()) == 3
    finally:
        builtins.print = plain_print
        if dist.is_initialized():
            dist.destroy_process_group()
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
