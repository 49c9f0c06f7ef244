"""Drop-in for the reference's `train_utils` package (train_utils/__init__.py:1-2)."""
from .train_and_eval import train_one_epoch, evaluate, create_lr_scheduler, criterion  # noqa: F401
from .distributed_utils import init_distributed_mode, save_on_master, mkdir  # noqa: F401
