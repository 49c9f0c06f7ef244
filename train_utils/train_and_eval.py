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


def train_one_epoch(model, optimizer, data_loader, device, epoch, num_classes, lr_scheduler, print_freq=10, scaler=None):
    model.train()
    metric_logger = utils.MetricLogger(delimiter="  ")
    metric_logger.add_meter('lr', utils.SmoothedValue(window_size=1, fmt='{value:.6f}'))
    header = 'Epoch: [{}]'.format(epoch)
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
