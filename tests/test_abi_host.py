"""CPU: the C-ABI library loads and exports every symbol include/egm_b200.h declares; host-side logic."""
import ctypes
import os
import re

import pytest
import torch

import egm_unet_b200  # noqa: F401
from egm_unet_b200 import abi


def test_library_exports_every_declared_symbol():
    assert os.path.exists(abi.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    protos = abi.parse_header()
    assert len(protos) >= 60
    cdll = ctypes.CDLL(abi.LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), name
    src = open(abi.HEADER).read()
    declared = set(re.findall(r"\b(egm_\w+)\s*\(", src))
    assert declared == set(protos), declared ^ set(protos)
    L = abi.lib()
    assert L.fn["egm_abi_version"]() == 1


def test_host_only_queries_need_no_gpu():
    assert abi.query("loss_workspace_bytes", 2, 2, 8, 8) >= 2 * 8 * 8
    assert abi.query("mca_vec_len", 3, 5, 7, 8) % 4 == 0
    assert abi.query("mca_vec_off_c", 3, 5, 7) >= 3 * 5 + 3 * 7


def test_no_cpu_fallback():
    import egm_unet_b200 as E
    m = E.UNet(3, 2, base_c=8)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 16, 16))
    with pytest.raises(RuntimeError):
        E.criterion({"out": torch.zeros(1, 2, 8, 8)}, torch.zeros(1, 8, 8, dtype=torch.int64))


def test_lr_scheduler_matches_reference_formula():
    from train_utils import create_lr_scheduler
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=0.02, momentum=0.9)
    sch = create_lr_scheduler(opt, num_step=10, epochs=5, warmup=True)
    lrs = []
    for _ in range(50):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    assert abs(lrs[0] - 0.02 * 1e-3) < 1e-12 and abs(lrs[10] - 0.02) < 1e-9
    assert abs(lrs[30] - 0.02 * (1 - 20 / 40) ** 0.9) < 1e-9


def test_confusion_matrix_and_meters_cpu():
    from train_utils.distributed_utils import ConfusionMatrix, SmoothedValue, MetricLogger
    cm = ConfusionMatrix(2)
    cm.update(torch.tensor([0, 1, 1, 255, 0]), torch.tensor([0, 1, 0, 1, 0]))
    assert cm.mat.tolist() == [[2, 0], [1, 1]]
    acc, _, iu = cm.compute()
    assert abs(float(acc) - 0.75) < 1e-6 and abs(float(iu[0]) - 2 / 3) < 1e-6
    sv = SmoothedValue(window_size=2)
    for v in (1.0, 2.0, 3.0):
        sv.update(v)
    assert sv.global_avg == 2.0 and sv.value == 3.0 and sv.max == 3.0
    ml = MetricLogger()
    ml.update(loss=1.5)
    assert "loss" in str(ml)


def test_fastdiv_reciprocal_is_exact():
    """csrc/common.cuh FastDiv (33-bit Granlund-Montgomery reciprocal used by every grid-stride kernel's index decoding):
    q = (umulhi(x, m) + x) >> s with s = ceil(log2 d), m = floor(2^32 (2^s - d) / d) + 1 must equal x // d for all x < 2^32."""
    import numpy as np
    rng = np.random.default_rng(0)
    ds = np.unique(np.concatenate([np.arange(1, 2050), rng.integers(1, 2 ** 31 - 1, 3000), [2 ** 31 - 1, 2 ** 30, 3 * 2 ** 20 + 1]])).astype(np.uint64)
    xs = np.concatenate([rng.integers(0, 2 ** 32, 4000, dtype=np.uint64), np.array([0, 1, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1], dtype=np.uint64)])
    for d in ds.tolist():
        s = 0 if d <= 1 else int(d - 1).bit_length()
        m = ((1 << 32) * ((1 << s) - d)) // d + 1
        assert m < 2 ** 32
        x = xs if d > 4096 else np.concatenate([xs, np.arange(0, 5 * d, dtype=np.uint64), (np.arange(1, 40, dtype=np.uint64) * np.uint64(d)) - np.uint64(1)])
        q = (((x * np.uint64(m)) >> np.uint64(32)) + x) >> np.uint64(s)
        assert np.array_equal(q, x // np.uint64(d)), d
