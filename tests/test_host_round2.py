"""CPU: host-side logic added in round 2 -- the staged-reference loader, bench.py's roofline bookkeeping, the drop-in
train loop's routing decision, init_distributed_mode's non-distributed branch."""
import argparse
import importlib.util
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("_bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_staged_reference_equals_oracle():
    """oracle/_ref (the unmodified reference, staged by build()) and the oracle restatement agree to fp32 round-off -- the pin the
    fixtures were generated with, re-checked wherever the staged copy exists."""
    from oracle import build_ref, egm_oracle as O, synth
    if not build_ref.available():
        pytest.skip("oracle/_ref not staged (no /root/reference on this machine)")
    _, tae, _ = build_ref.load()
    import train_utils
    assert train_utils.__file__.startswith(ROOT), "loading the reference must not shadow the repo's own train_utils"
    lw = torch.tensor([1.0, 2.0])
    for variant in ("unet", "egm", "yuan"):
        m = build_ref.build_model(variant)
        sd = synth.fill_state_dict(m.state_dict())
        m.load_state_dict(sd)
        m.train()
        x, t = synth.make_inputs(2, 48, 40)
        out = m(x)["out"]
        ref = O.forward(sd, x, variant, True)
        assert float((out - ref).abs().max() / ref.abs().max()) < 2e-5, variant
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            l_ref = tae.criterion({"out": out}, t, lw, num_classes=2, ignore_index=255)
        assert abs(float(l_ref) - float(O.criterion(ref, t, lw))) <= 1e-5 * abs(float(l_ref))


def test_bench_flop_and_byte_bookkeeping():
    b = _bench()
    # 70.47 GFLOP / image forward, 211.0 GFLOP / image training at 480^2
    assert abs(b.doubleconv_flops_per_image(train=False) / 1e9 - 70.47) < 0.01
    assert abs(b.doubleconv_flops_per_image(train=True) / 1e9 - 211.0) < 0.1
    key = "conv2d_tc_view:32,0,32,32,0,32,0,16,480,480,32,32,3,3,1"
    assert b.conv_shape(key) == (16, 480, 480, 32, 32, 3, 3, 1) and b.is_doubleconv(key, 16)
    assert not b.is_doubleconv("conv2d_tc_view:16,0,16,16,0,16,0,16,240,240,16,16,3,3,12", 16)      # dilated GRFB branch
    assert not b.is_doubleconv("conv2d_tc_view:64,0,64,16,0,16,0,16,240,240,64,16,1,1,1", 16)       # 1x1
    fam, nbytes = b.hbm_family_bytes("bn_act_fwd:32,0,1,0,32,0,1,3686400,32")
    assert fam == "batchnorm" and nbytes == 3686400 * 32 * 2 * 2
    fam, nbytes = b.hbm_family_bytes("maxpool2x2_fwd:1,16,480,480,32")
    assert fam == "pool" and nbytes == 16 * 480 * 480 * 32 * 2 * 1.25
    fam, nbytes = b.hbm_family_bytes("upsample_concat_fwd:1,16,240,240,480,480,32,32")
    assert fam == "upsample_concat" and nbytes == (16 * 240 * 240 * 32 + 16 * 480 * 480 * 32 + 16 * 480 * 480 * 64) * 2
    fam, nbytes = b.hbm_family_bytes("loss_fwd_bwd:16,2,480,480,255,1,123")
    assert fam == "loss" and nbytes == 16 * 480 * 480 * (8 + 8 + 8)
    assert b.hbm_family_bytes("sgd_step:6302840")[0] is None


def test_fused_route_is_refused_without_cuda_or_plain_sgd():
    import train_utils.train_and_eval as tae
    import egm_unet_b200 as E
    model = E.UNet(3, 2, base_c=8)
    sgd = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    assert tae._fused_trainer(model, sgd, 2) is None                                   # CPU parameters: no CUDA path, no fallback
    assert tae._fused_trainer(model, torch.optim.Adam(model.parameters()), 2) is None
    assert tae._fused_trainer(torch.nn.Linear(2, 2), sgd, 2) is None
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(torch.zeros(1, 3, 16, 16))


def test_init_distributed_mode_without_rendezvous_env(capsys, monkeypatch):
    import train_utils
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "SLURM_PROCID"):
        monkeypatch.delenv(k, raising=False)
    args = argparse.Namespace(dist_url="env://")
    train_utils.init_distributed_mode(args)
    assert args.distributed is False and "Not using distributed mode" in capsys.readouterr().out


def test_integration_doc_maps_every_header_symbol():
    """INTEGRATION.md is the reference-side binding guide: every entry point include/egm_b200.h declares must be named there
    (`egm_x_fwd/bwd` shorthand counts for both directions)."""
    import re
    hdr = open(os.path.join(ROOT, "include", "egm_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    syms = list(dict.fromkeys(re.findall(r"\b(egm_\w+)\s*\(", hdr)))
    assert len(syms) >= 90
    missing = [s for s in syms if s not in doc and not (s.endswith(("_fwd", "_bwd")) and s[:-4] in doc)]
    assert not missing, missing


def test_bench_family_bookkeeping_knows_the_view_entry_points():
    """bench.py's HBM-family table keys on C-ABI call names: the strided MaxPool forms and the fused Up path must land in their families."""
    b = _bench()
    assert b.hbm_family_bytes("maxpool2x2_fwd_view:64,0,1,16,480,480,32") == ("pool", 16 * 480 * 480 * 32 * 2 * 1.25)
    assert b.hbm_family_bytes("maxpool2x2_bwd_view:32,0,32,0,1,1,16,480,480,32")[0] == "pool"
    fam, nbytes = b.hbm_family_bytes("upsample_concat_fwd:1,16,240,240,480,480,32,32")
    assert fam == "upsample_concat" and nbytes == (16 * 240 * 240 * 32 + 16 * 480 * 480 * 32 + 16 * 480 * 480 * 64) * 2
    assert b.is_doubleconv("conv2d_tc_ex:32,0,32,32,0,32,0,16,480,480,32,32,3,3,1,0") and b.is_doubleconv("conv2d_wgrad_tc_view:64,0,64,32,0,32,16,480,480,64,32,3,3,1")
    assert not b.is_doubleconv("conv2d_tc_view:16,0,16,16,0,16,0,16,240,240,16,16,1,1,1")


def test_ncu_conv_traffic_matches_launches_to_calls(tmp_path):
    """tools/ncu_conv_traffic.py: the i-th conv kernel of an ncu launch list is the i-th conv call of the step's call log; DRAM bytes of
    the DoubleConv launches are summed per step (bench.py's roofline.traffic)."""
    import json
    import subprocess
    hdr = '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"'
    rows = [hdr]

    def add(i, name, ns, rd, wr):
        base = f'"{i}","1","python","h","{name}","1","7","(192, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics"'
        rows.append(base + f',"dram__bytes_read.sum","Mbyte","{rd}"')
        rows.append(base + f',"dram__bytes_write.sum","Kbyte","{wr}"')
        rows.append(base + f',"gpu__time_duration.sum","us","{ns}"')
    add(0, "void k_nchw_to_nhwc<__nv_bfloat16>(const float *, T1 *, long long, int, long long)", 20.0, 44.0, 1.0)
    add(1, "void k_conv_tc_halo<32, 1>(CUtensorMap_st, CUtensorMap_st, __nv_bfloat16 *, const float *, ConvHaloParams)", 150.0, 236.0, 186000.0)
    add(2, "void k_conv_tc_halo<0, 1>(CUtensorMap_st, CUtensorMap_st, __nv_bfloat16 *, const float *, ConvHaloParams)", 25.0, 30.0, 100.0)
    add(3, "k_wgrad_tc_halo(CUtensorMap_st, CUtensorMap_st, float *, WgradHaloParams)", 120.0, 480.0, 4000.0)
    csvp, callp, outp = tmp_path / "l.csv", tmp_path / "c.txt", tmp_path / "o.json"
    csvp.write_text("==PROF== noise\n" + "\n".join(rows) + "\n")
    callp.write_text("nchw_to_nhwc:0,16,3,480,480\n"
                     "conv2d_tc_ex:32,0,32,32,0,32,0,16,480,480,32,32,3,3,1,0\n"
                     "conv2d_tc_view:16,0,16,16,0,16,0,16,240,240,16,16,1,1,1\n"
                     "bn_act_fwd:32,0,1,0,32,0,1,3686400,32\n"
                     "conv2d_wgrad_tc_view:32,0,32,32,0,32,16,480,480,32,32,3,3,1\n")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_conv_traffic.py"), str(csvp), str(callp), str(outp)], stdout=subprocess.DEVNULL)
    o = json.load(open(outp))
    assert o["launches_in_step"] == 4 and o["conv"]["launches"] == 3
    assert o["doubleconv"]["launches"] == 2                                    # the 16->16 1x1 is not a DoubleConv layer
    assert abs(o["doubleconv_dram_bytes_per_step"] - (236e6 + 186e6 + 480e6 + 4e6)) < 1.0
    assert abs(o["doubleconv"]["ms"] - 0.27) < 1e-9
