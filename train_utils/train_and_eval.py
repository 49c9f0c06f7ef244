"""API mirror of the reference's train_utils/train_and_eval.py (criterion :7, evaluate :22, train_one_epoch :43,
create_lr_scheduler :78) running on the libegm_b200 kernels."""
import torch

import train_utils.distributed_utils as utils
from egm_unet_b200.loss import criterion, EvalMetrics  # noqa: F401  (criterion: same signature as the reference)


def evaluate(model, data_loader, device, num_classes):
    model.eval()
    confmat = utils.ConfusionMatrix(num_classes)
    dice = utils.DiceCoefficient(num_classes=num_classes, ignore_index=255)
    metric_logger = utils.MetricLogger(delimiter="  ")
    header = 'Test:'
    fused = None
    with torch.no_grad():
        for image, target in metric_logger.log_every(data_loader, 100, header):
            image, target = image.to(device), target.to(device)
            output = model(image)['out']
            if fused is None:
                fused = EvalMetrics(num_classes, 255, output.device)
            fused.update(output, target)               # one fused launch: argmax + confusion + dice sums
        if fused is not None:
            confmat.mat = fused.confusion().clone()
            fused._drain()
            dice.cumulative_dice = torch.tensor([fused.dice_sum], dtype=torch.float32, device=device)
            dice.count = torch.tensor([float(fused.count)], dtype=torch.float32, device=device)
        confmat.reduce_from_all_processes()
        dice.reduce_from_all_processes()
    return confmat, dice.value.item()


def _fused_trainer(model, optimizer, num_classes):
    """The fused step (egm_unet_b200.trainer.Trainer: forward + criterion + backward + gradient all-reduce + SGD as ONE replayed
    CUDA graph) bound to (model, optimizer) -- possible when the optimizer is the plain SGD train.py:113-118 builds over all of
    the model's parameters.  Returns None otherwise (Adam, several parameter groups, nesterov, frozen parameters ...): the
    caller then runs the generic autograd loop.  The optimizer stays the owner of the hyper-parameters (lr / momentum /
    weight_decay are re-read every step, so LR schedulers work) and its `momentum_buffer` state tensors alias the trainer's flat
    momentum buffer, so `optimizer.state_dict()` checkpoints (train.py:152-156) and `load_state_dict` resumes keep working."""
    from egm_unet_b200.models import _B200Net
    from egm_unet_b200.trainer import Trainer
    if not isinstance(model, _B200Net) or type(optimizer) is not torch.optim.SGD or len(optimizer.param_groups) != 1:
        return None
    g = optimizer.param_groups[0]
    params = list(model.parameters())
    if g.get("nesterov") or g.get("dampening", 0) != 0 or g.get("maximize") or len(g["params"]) != len(params):
        return None
    if any(a is not b for a, b in zip(g["params"], params)) or not all(p.requires_grad and p.is_cuda for p in params):
        return None
    tr = getattr(model, "_egm_trainer", None)
    if tr is None or tr._owner is not optimizer or tr.num_classes != num_classes:
        cw = [1.0, 2.0] if num_classes == 2 else None
        tr = Trainer(model, lr=g["lr"], momentum=g["momentum"], weight_decay=g["weight_decay"], class_weight=cw, ignore_index=255,
                     use_graph=True)
        tr._owner, tr.num_classes = optimizer, num_classes
        object.__setattr__(model, "_egm_trainer", tr)
    if not tr.store.valid():
        tr._rebind_store()
    # momentum: optimizer.state[p]["momentum_buffer"] <-> view of the flat buffer (copy in whatever a resume loaded)
    st = tr.store
    for p, off in zip(st.plist, st.offsets):
        view = tr.mom_buf[off:off + p.numel()].view(p.shape)
        cur = optimizer.state[p].get("momentum_buffer") if p in optimizer.state else None
        if cur is not None and cur.data_ptr() != view.data_ptr():
            view.copy_(cur.to(view.device, torch.float32))
        optimizer.state[p]["momentum_buffer"] = view
    return tr


_SCALER_WARNED = [False]


def train_one_epoch(model, optimizer, data_loader, device, epoch, num_classes, lr_scheduler, print_freq=10, scaler=None):
    """Reference signature and return value (train_utils/train_and_eval.py:43-75).

    `scaler`: the reference uses a GradScaler for fp16 autocast (:57-65).  The B200 path stores activations in bf16 and
    accumulates in fp32 inside its own kernels; bf16 has fp32's exponent range, so no loss scaling is needed.  A scaler is
    accepted for call compatibility and deliberately left untouched (a warning says so once)."""
    if scaler is not None and not _SCALER_WARNED[0]:
        import warnings
        warnings.warn("train_one_epoch: `scaler` is accepted for API compatibility but not used -- the B200 path computes in "
                      "bf16 storage / fp32 accumulate, which needs no loss scaling")
        _SCALER_WARNED[0] = True
    model.train()
    metric_logger = utils.MetricLogger(delimiter="  ")
    metric_logger.add_meter('lr', utils.SmoothedValue(window_size=1, fmt='{value:.6f}'))
    header = 'Epoch: [{}]'.format(epoch)
    tr = _fused_trainer(model, optimizer, num_classes) if torch.device(device).type == "cuda" else None
    if tr is not None:
        # fast path: batches are prefetched to the device on a copy stream, each step is one CUDA-graph replay, and the loss is
        # read back lazily (two steps later) instead of the reference's per-step `loss.item()` sync (:73).  MetricLogger sees every
        # loss exactly once, in order; by the end of the epoch its statistics are those of the reference loop.
        g = optimizer.param_groups[0]

        def pre_step():
            tr.set_hyper(g["lr"], g["momentum"], g["weight_decay"])

        def post_step():
            optimizer._opt_called = True          # the step ran inside the fused kernel sequence
            lr_scheduler.step()
            metric_logger.update(lr=g["lr"])      # before the loader is advanced: log_every prints right after its yield returns

        for ready in tr.run_iter(metric_logger.log_every(data_loader, print_freq, header), pre_step, post_step):
            for v in ready:
                metric_logger.update(loss=v)
        return metric_logger.meters["loss"].global_avg, g["lr"]
    loss_weight = torch.as_tensor([1.0, 2.0], device=device) if num_classes == 2 else None
    for image, target in metric_logger.log_every(data_loader, print_freq, header):
        image, target = image.to(device, non_blocking=True), target.to(device, non_blocking=True)
        # the B200 path computes in bf16 storage / fp32 accumulate itself; autocast + GradScaler are not needed
        output = model(image)
        loss = criterion(output, target, loss_weight, num_classes=num_classes, ignore_index=255)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        lr_scheduler.step()
        lr = optimizer.param_groups[0]["lr"]
        metric_logger.update(loss=loss.item(), lr=lr)
    return metric_logger.meters["loss"].global_avg, lr


def create_lr_scheduler(optimizer, num_step: int, epochs: int, warmup=True, warmup_epochs=1, warmup_factor=1e-3):
    assert num_step > 0 and epochs > 0
    if warmup is False:
        warmup_epochs = 0

    def f(x):
        if warmup is True and x <= (warmup_epochs * num_step):
            alpha = float(x) / (warmup_epochs * num_step)
            return warmup_factor * (1 - alpha) + alpha
        return (1 - (x - warmup_epochs * num_step) / ((epochs - warmup_epochs) * num_step)) ** 0.9

    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=f)
