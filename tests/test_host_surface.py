"""CPU tests of the drop-in surface that needs no GPU: module constructors /
state_dict, launcher stubs, CLI error paths, no-CPU-fallback behaviour."""
import sys
import types

import numpy as np
import pytest
import torch

import pathnet_b200 as pb
from pathnet_b200 import cli, run


def test_module_surface_and_reference_checkpoint_keys():
    m = pb.PathNet(1703, 128, 5, 4)
    keys = set(m.state_dict().keys())
    want = {"fc0.weight", "fc0.bias", "LSTM.weight_ih_l0", "LSTM.weight_hh_l0", "LSTM.bias_ih_l0",
            "LSTM.bias_hh_l0", "fc2.weight", "fc2.bias", "attw.weight", "attw.bias"} | \
           {f"nets.{d}.{k}" for d in range(4) for k in ("weight", "bias")}
    assert keys == want
    assert pb.PAGG is pb.PathNet
    h = pb.PathNet_homo(10, 64, 3, 6, aggr="mean")
    assert len(h.nets) == 6 and h.attw.weight.shape == (1, 128)


def test_reference_checkpoint_loads(have_reference):
    if not have_reference:
        pytest.skip("needs /root/reference/saved_models/cornell.pth")
    sd = torch.load("/root/reference/saved_models/cornell.pth", map_location="cpu")
    m = pb.PathNet(1703, 128, 5, 4)
    m.load_state_dict(sd)          # strict: same names and shapes


def test_same_seed_same_init_as_reference(have_reference):
    """Construction order mirrors the reference, so torch.manual_seed gives the
    same initial weights as the reference classes."""
    if not have_reference:
        pytest.skip("needs /root/reference/PathNet_run.py")
    from oracle import pagg_oracle as po
    ns = po.load_reference_classes()
    for ours, theirs in ((pb.PathNet, "PathNet"), (pb.PathNet_homo, "PathNet_homo")):
        torch.manual_seed(5); a = ours(20, 32, 3, 4)
        torch.manual_seed(5); b = ns[theirs](20, 32, 3, 4)
        for (k1, v1), (k2, v2) in zip(a.state_dict().items(), b.state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2), k1


def test_forward_on_cpu_raises_instead_of_falling_back():
    m = pb.PathNet_homo(8, 32, 3, 2)
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.randn(5, 8), torch.zeros((1, 6), dtype=torch.int64), 3, 2, np.array([1, 0, 0, 0, 0], bool),
          torch.zeros((1, 3, 2), dtype=torch.int64), None)


def test_launcher_stub_rebinds_forward_of_reference_class_names():
    run.install_stubs()
    import torch_geometric.nn as tgnn
    assert tgnn.MessagePassing is run.MessagePassing

    class PathNet(tgnn.MessagePassing):      # what PathNet_run.py:150 declares
        def forward(self, *a):
            return "reference forward"

    class Other(tgnn.MessagePassing):
        def forward(self, *a):
            return "untouched"

    assert PathNet.forward.__qualname__.startswith("_fused_forward")
    assert Other().forward() == "untouched"
    import numpy as _np
    assert hasattr(_np, "bool")
    from scipy.sparse.linalg.eigen.arpack import eigsh  # noqa: F401  (dataset.py:23)


def test_cli_argument_errors_mirror_reference(capsys):
    assert cli.main(["gen_merw", "cornell", "40"]) == 0           # gen_merw.cpp:128-132
    assert "ERROR: Incorrect number of parameters." in capsys.readouterr().err
    assert cli.main(["nonsense"]) == 2


def test_mask_rows_cache_follows_in_place_changes():
    """RowsCache keeps nonzero(mask) per mask (torch: tensor object + in-place version; numpy: by content)."""
    import torch
    import numpy as np
    from pathnet_b200.pagg import PathNet_homo, RowsCache
    m = PathNet_homo(8, 16, 3, 4)
    rc, cpu = m._rows_cache, torch.device("cpu")
    assert isinstance(rc, RowsCache)
    mask = torch.tensor([True, False, True, True])
    r1 = rc.get(mask, cpu)
    assert r1.dtype == torch.int32 and r1.tolist() == [0, 2, 3]
    assert rc.get(mask, cpu) is r1                                 # same object, same version: reused
    mask[1] = True                                                 # in-place edit bumps the version counter
    assert rc.get(mask, cpu).tolist() == [0, 1, 2, 3]
    other = torch.tensor([False, True, False, False])              # a different tensor never hits another's entry
    assert rc.get(other, cpu).tolist() == [1]
    a = np.array([0, 1, 1, 0])
    r2 = rc.get(a, cpu)
    assert r2.tolist() == [1, 2]
    assert rc.get(np.array([0, 1, 1, 0]), cpu) is r2               # numpy masks: by content (the reference passes numpy bool masks)
    a[0] = 1
    assert rc.get(a, cpu).tolist() == [0, 1, 2]                    # a mutated array is a different key
    for _ in range(20):                                            # bounded
        rc.get(torch.tensor([True]), cpu)
    assert len(rc.ents) <= 8


def test_math_mode_resolution(monkeypatch):
    import warnings
    import pytest
    from pathnet_b200 import pagg
    assert pagg.resolve_math_mode("fp32", 128, 4) == "fp32"
    assert pagg.resolve_math_mode("tc", 128, 4) == "tc"
    assert pagg.resolve_math_mode("auto", 64, 4) == "fp32"        # no tensor-core kernels at this hidden size
    monkeypatch.setenv("PN_MATH_MODE", "fp32")
    assert pagg.resolve_math_mode("auto", 128, 4) == "fp32"       # env override of the default
    monkeypatch.delenv("PN_MATH_MODE")
    pagg._warned_auto_tc = False
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert pagg.resolve_math_mode("auto", 128, 4) == "tc3"      # parity first: the split-precision tensor-core mode
        assert pagg.resolve_math_mode("auto", 128, 4) == "tc3"
    assert len(w) == 1 and "'tc'" in str(w[0].message)            # one notice naming the faster opt-in mode
    assert pagg.resolve_math_mode("tc3", 128, 4) == "tc3"
    with pytest.raises(ValueError):
        pagg.resolve_math_mode("bf16", 128, 4)


def test_math_mode_codes_match_the_header_and_size_the_workspace():
    """pagg.MATH mirrors `pn_math` of include/pathnet_b200.h; the workspace query is host-only arithmetic, so the per-mode
    sizing can be checked without a GPU: the split-precision mode carries the gathered table as two fp16 terms (as many bytes
    as the fp32 table), the fast mode as one (half), plus the fp16 operand copies of its TMA-fed projections."""
    import ctypes as C
    import re
    from pathlib import Path
    from pathnet_b200 import _lib, pagg
    hdr = (Path(__file__).resolve().parents[1] / "include" / "pathnet_b200.h").read_text()
    codes = {m.group(1): int(m.group(2)) for m in re.finditer(r"PN_MATH_(FP32|TC3|TC)\s*=\s*(\d+)", hdr)}
    assert codes == {"FP32": pagg.MATH["fp32"], "TC": pagg.MATH["tc"], "TC3": pagg.MATH["tc3"]}
    N, F, H, Cc, S, nw, L = 1000, 40, 128, 4, 300, 5, 4

    def ws(mode, **kw):
        cfg = _lib.PaggCfg(N=N, F=F, H=H, C=Cc, S=S, nw=nw, L=L, variant=0, math_mode=pagg.MATH[mode], training=1, dropout_p=0.5,
                           dropout_seed=1, no_backward=0, row_chunk=kw.get("row_chunk", 0), dropout_seed_dev=None, x16=None, cell=0, pool=0)
        return _lib.lib().pn_pagg_workspace_bytes(C.byref(cfg))

    b32, btc, b3 = ws("fp32"), ws("tc"), ws("tc3")
    assert b32 > 0 and btc > b32 and b3 > b32
    table = N * L * H * 4
    assert b3 - b32 >= table and btc - b32 >= table // 2      # two fp16 terms vs one (+ mask words, + the fast mode's fp16 copies)
    assert ws("tc3", row_chunk=50) < b3                        # row chunking bounds the per-sequence part in this mode too
    bad = _lib.PaggCfg(N=N, F=F, H=H, C=Cc, S=S, nw=nw, L=L, variant=0, math_mode=7, training=0, dropout_p=0.0, dropout_seed=0,
                       no_backward=0, row_chunk=0, dropout_seed_dev=None, x16=None, cell=0, pool=0)
    assert _lib.lib().pn_pagg_workspace_bytes(C.byref(bad)) == 0      # unknown math mode: rejected by the config check
    gru = _lib.PaggCfg(N=N, F=F, H=H, C=Cc, S=S, nw=nw, L=L, variant=0, math_mode=pagg.MATH["tc3"], training=0, dropout_p=0.0,
                       dropout_seed=0, no_backward=0, row_chunk=0, dropout_seed_dev=None, x16=None, cell=2, pool=0)
    assert _lib.lib().pn_pagg_workspace_bytes(C.byref(gru)) == 0      # the tensor-core modes implement the LSTM cell only


def test_dist_shutdown_is_a_no_op_without_a_process_group():
    from pathnet_b200.dist import shutdown
    shutdown(1.0)   # nothing initialised: returns at once
