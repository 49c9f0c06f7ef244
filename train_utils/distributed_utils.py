"""API mirror of the reference's train_utils/distributed_utils.py (meters :14-73/:170-260, ConfusionMatrix :76-125,
DiceCoefficient :128-167, process-group helpers :271-338).  Host-side logic is rewritten from the behaviour; the
metric arithmetic on the drop-in `evaluate` path runs in the fused CUDA kernel (egm_unet_b200.loss.EvalMetrics)."""
import builtins
import datetime
import os
import time
from collections import defaultdict, deque

import torch
import torch.distributed as dist


def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process():
    return get_rank() == 0


def save_on_master(*args, **kwargs):
    if is_main_process():
        torch.save(*args, **kwargs)


def mkdir(path):
    os.makedirs(path, exist_ok=True)


def setup_for_distributed(is_master):
    """Silence print() on non-master ranks unless called with force=True."""
    plain = builtins.print

    def rank_print(*args, **kwargs):
        if kwargs.pop('force', False) or is_master:
            plain(*args, **kwargs)

    builtins.print = rank_print


def init_distributed_mode(args):
    """Same contract as distributed_utils.py:315-338: read RANK/WORLD_SIZE/LOCAL_RANK (or SLURM_PROCID), bind the
    GPU, create the NCCL process group (NVLink/NVSwitch on a B200 box) and mute non-master prints."""
    env = os.environ
    if 'RANK' in env and 'WORLD_SIZE' in env:
        args.rank, args.world_size, args.gpu = int(env['RANK']), int(env['WORLD_SIZE']), int(env['LOCAL_RANK'])
    elif 'SLURM_PROCID' in env:
        args.rank = int(env['SLURM_PROCID'])
        args.gpu = args.rank % torch.cuda.device_count()
    elif not hasattr(args, 'rank'):
        print('Not using distributed mode')
        args.distributed = False
        return
    args.distributed = True
    torch.cuda.set_device(args.gpu)
    args.dist_backend = 'nccl'
    print('| distributed init (rank {}): {}'.format(args.rank, args.dist_url), flush=True)
    dist.init_process_group(backend=args.dist_backend, init_method=args.dist_url, world_size=args.world_size, rank=args.rank)
    setup_for_distributed(args.rank == 0)


class SmoothedValue(object):
    """Windowed + global statistics of a scalar series."""

    def __init__(self, window_size=20, fmt=None):
        self.fmt = fmt or "{value:.4f} ({global_avg:.4f})"
        self.deque = deque(maxlen=window_size)
        self.total, self.count = 0.0, 0

    def update(self, value, n=1):
        self.deque.append(value)
        self.count += n
        self.total += value * n

    def synchronize_between_processes(self):
        if not is_dist_avail_and_initialized():
            return
        t = torch.tensor([self.count, self.total], dtype=torch.float64, device='cuda')
        dist.barrier()
        dist.all_reduce(t)
        self.count, self.total = int(t[0].item()), t[1].item()

    @property
    def median(self):
        return torch.tensor(list(self.deque)).median().item()

    @property
    def avg(self):
        return torch.tensor(list(self.deque), dtype=torch.float32).mean().item()

    @property
    def global_avg(self):
        return self.total / self.count

    @property
    def max(self):
        return max(self.deque)

    @property
    def value(self):
        return self.deque[-1]

    def __str__(self):
        return self.fmt.format(median=self.median, avg=self.avg, global_avg=self.global_avg, max=self.max, value=self.value)


class ConfusionMatrix(object):
    def __init__(self, num_classes):
        self.num_classes = num_classes
        self.mat = None

    def update(self, a, b):
        """a: flattened ground truth, b: flattened prediction (API of :81-91)."""
        n = self.num_classes
        if self.mat is None:
            self.mat = torch.zeros((n, n), dtype=torch.int64, device=a.device)
        with torch.no_grad():
            k = (a >= 0) & (a < n)
            self.mat += torch.bincount(n * a[k].to(torch.int64) + b[k], minlength=n ** 2).reshape(n, n)

    def reset(self):
        if self.mat is not None:
            self.mat.zero_()

    def compute(self):
        h = self.mat.float()
        diag = torch.diag(h)
        return diag.sum() / h.sum(), diag / h.sum(1), diag / (h.sum(1) + h.sum(0) - diag)

    def reduce_from_all_processes(self):
        if is_dist_avail_and_initialized():
            dist.barrier()
            dist.all_reduce(self.mat)

    def __str__(self):
        acc_global, acc, iu = self.compute()
        return ('global correct: {:.1f}\naverage row correct: {}\nIoU: {}\nmean IoU: {:.1f}').format(
            acc_global.item() * 100, ['{:.1f}'.format(i) for i in (acc * 100).tolist()],
            ['{:.1f}'.format(i) for i in (iu * 100).tolist()], iu.mean().item() * 100)


class DiceCoefficient(object):
    def __init__(self, num_classes: int = 2, ignore_index: int = -100):
        self.cumulative_dice = None
        self.num_classes = num_classes
        self.ignore_index = ignore_index
        self.count = None

    def update(self, pred, target):
        """pred: logits [N,C,H,W]; Dice of the arg-max mask vs target over foreground classes (:135-144)."""
        from egm_unet_b200.loss import EvalMetrics
        if self.cumulative_dice is None:
            self.cumulative_dice = torch.zeros(1, dtype=torch.float32, device=pred.device)
            self.count = torch.zeros(1, dtype=torch.float32, device=pred.device)
        m = EvalMetrics(self.num_classes, self.ignore_index, pred.device)
        m.update(pred, target)
        self.cumulative_dice += m.dice
        self.count += 1

    @property
    def value(self):
        if self.count is None or self.count == 0:
            return 0
        return self.cumulative_dice / self.count

    def reset(self):
        if self.cumulative_dice is not None:
            self.cumulative_dice.zero_()
        if self.count is not None:
            self.count.zero_()

    def reduce_from_all_processes(self):
        if is_dist_avail_and_initialized():
            dist.barrier()
            dist.all_reduce(self.cumulative_dice)
            dist.all_reduce(self.count)


class MetricLogger(object):
    def __init__(self, delimiter="\t"):
        self.meters = defaultdict(SmoothedValue)
        self.delimiter = delimiter

    def update(self, **kwargs):
        for k, v in kwargs.items():
            if isinstance(v, torch.Tensor):
                v = v.item()
            assert isinstance(v, (float, int))
            self.meters[k].update(v)

    def __getattr__(self, attr):
        if attr in self.meters:
            return self.meters[attr]
        if attr in self.__dict__:
            return self.__dict__[attr]
        raise AttributeError("'{}' object has no attribute '{}'".format(type(self).__name__, attr))

    def __str__(self):
        return self.delimiter.join("{}: {}".format(name, str(meter)) for name, meter in self.meters.items())

    def synchronize_between_processes(self):
        for meter in self.meters.values():
            meter.synchronize_between_processes()

    def add_meter(self, name, meter):
        self.meters[name] = meter

    def log_every(self, iterable, print_freq, header=None):
        header = header or ''
        total = len(iterable)
        iter_time, data_time = SmoothedValue(fmt='{avg:.4f}'), SmoothedValue(fmt='{avg:.4f}')
        fields = [header, '[{0:' + str(len(str(total))) + 'd}/{1}]', 'eta: {eta}', '{meters}', 'time: {time}', 'data: {data}']
        cuda = torch.cuda.is_available()
        if cuda:
            fields.append('max mem: {memory:.0f}')
        log_msg = self.delimiter.join(fields)
        start = end = time.time()
        for i, obj in enumerate(iterable):
            data_time.update(time.time() - end)
            yield obj
            iter_time.update(time.time() - end)
            if i % print_freq == 0:
                eta = str(datetime.timedelta(seconds=int(iter_time.global_avg * (total - i))))
                kw = dict(eta=eta, meters=str(self), time=str(iter_time), data=str(data_time))
                if cuda:
                    kw['memory'] = torch.cuda.max_memory_allocated() / (1024.0 * 1024.0)
                print(log_msg.format(i, total, **kw))
            end = time.time()
        print('{} Total time: {}'.format(header, str(datetime.timedelta(seconds=int(time.time() - start)))))
