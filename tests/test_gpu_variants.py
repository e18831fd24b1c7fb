"""GPU tests of the SURVEY 8f-4 variants: the reference's older `class PAGG` (tanh RNN + mean pooling,
baseline/GPRGNN/src/copy.py:299-359) and the pooling ablations (mean / sum / max instead of the attention)."""
import numpy as np
import pytest
import torch

from oracle import pagg_oracle as po
from pathnet_b200 import _lib, pagg
from tests.conftest import GOLDEN
from tests.test_gpu_pagg import FWD_TOL, GRAD_TOL, _random_case, np_dropout_hash
from tests.test_gpu_tc import assert_grad_close, assert_tc_logits

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("path", sorted(GOLDEN.glob("paggrnn_*.npz")), ids=lambda p: p.stem)
def test_pagg_rnn_module_matches_reference_fixture(path):
    """pathnet_b200.PAGG_RNN (same ctor / forward / parameter names) against the reference class's own outputs."""
    z = np.load(path)
    m = pagg.PAGG_RNN(int(z["F"]), int(z["H"]), int(z["C"]), int(z["N"]))
    m.load_state_dict({k[6:]: torch.tensor(z[k]) for k in z.files if k.startswith("param:")})   # the reference's state_dict loads as is
    m = m.to(DEV).eval()
    nw, L, S = int(z["nw"]), int(z["L"]), int(z["S"])
    y = m(torch.tensor(z["X"], device=DEV), torch.tensor(z["neis"].astype(np.int64)), nw, L, z["mask"],
          torch.tensor(z["layer_type"].astype(np.int64)), torch.arange(S * nw * L, device=DEV))
    torch.testing.assert_close(y.cpu(), torch.tensor(z["logits"]), **FWD_TOL)
    (y * torch.tensor(z["G"], device=DEV)).sum().backward()
    for k, p in m.named_parameters():
        torch.testing.assert_close(p.grad.cpu(), torch.tensor(z["grad:" + k]), **GRAD_TOL, msg=lambda s: f"{k}: {s}")


def test_pagg_rnn_training_mode_with_identical_masks():
    """train(): dropout p = 0.9 on the RNN input and on the concat (copy.py:351,360), masks from the kernels' hash."""
    N, Fd, H, C, nw, L, S, p, seed = 90, 14, 64, 4, 5, 4, 33, 0.9, 0xFEED1234
    torch.manual_seed(2)
    m = pagg.PAGG_RNN(Fd, H, C, N)
    g = torch.Generator().manual_seed(8)
    X = torch.randn(N, Fd, generator=g)
    mask = torch.zeros(N, dtype=torch.bool); mask[torch.randperm(N, generator=g)[:S]] = True
    neis, lt = torch.randint(0, N, (S, nw * L), generator=g), torch.randint(0, L, (S, nw, L), generator=g)
    thr, scale = int(p * 65536), 1.0 / (1.0 - p)
    keep_seq = np_dropout_hash(seed, 0, np.arange(S * nw * L * H, dtype=np.uint64)) >= thr      # canonical order [S, nw, L, H] = [P, L, H]
    keep_cat = np_dropout_hash(seed, 1, np.arange(S * 2 * H, dtype=np.uint64)) >= thr
    ms = torch.tensor(keep_seq.reshape(S * nw, L, H).astype(np.float64) * scale).transpose(0, 1)   # the oracle wants [L, P, H]
    mc = torch.tensor(keep_cat.reshape(S, 2 * H).astype(np.float64) * scale)
    params = {k: v.detach().clone().double().requires_grad_(True) for k, v in m.state_dict().items()}
    y_ref = po.pagg_rnn_forward(params, X.double(), neis, nw, L, mask, lt, mask_seq=ms, mask_cat=mc)
    G = torch.randn(S, C)
    (y_ref * G.double()).sum().backward()
    m = m.to(DEV).train()
    m.dropout = p
    torch.manual_seed(0)
    # route the explicit seed through the functional form (the module draws its own)
    H_ = m.hidden_size
    prm = {"fc0.weight": m.fc0.weight, "fc0.bias": m.fc0.bias, "fc2.weight": m.fc2.weight, "fc2.bias": m.fc2.bias,
           "LSTM.weight_ih_l0": m.RNN.weight_ih_l0, "LSTM.weight_hh_l0": m.RNN.weight_hh_l0, "LSTM.bias_ih_l0": m.RNN.bias_ih_l0,
           "LSTM.bias_hh_l0": m.RNN.bias_hh_l0, "attw.weight": torch.zeros(1, 2 * H_, device=DEV), "attw.bias": torch.zeros(1, device=DEV)}
    for d, lin in enumerate(m.nets):
        prm[f"nets.{d}.weight"], prm[f"nets.{d}.bias"] = lin.weight, lin.bias
    sn, st, pt = pagg.canonicalize("plain", neis.to(DEV), lt.to(DEV), nw, L, N=N)
    rows = torch.nonzero(mask).reshape(-1).to(DEV, dtype=torch.int32)
    y = pagg.pagg_forward("plain", X.to(DEV), prm, sn, st, pt, rows, dropout=p, training=True, seed=seed, math_mode="fp32",
                          cell="rnn", pool="mean")
    torch.testing.assert_close(y.cpu(), y_ref.detach().float(), **FWD_TOL)
    (y * G.to(DEV)).sum().backward()
    for k, q in m.named_parameters():
        torch.testing.assert_close(q.grad.cpu(), params[k].grad.float(), **GRAD_TOL, msg=lambda s: f"{k}: {s}")


@pytest.mark.parametrize("variant", ["hetero", "homo"])
@pytest.mark.parametrize("pool", ["mean", "sum", "max"])
@pytest.mark.parametrize("mode,H", [("fp32", 64), ("tc", 128)])
def test_pooling_ablations_match_oracle(variant, pool, mode, H):
    N, Fd, Cc, nw, L, S = 130, 18, 5, 6, 4, 47
    cls = pagg.PathNet if variant == "hetero" else pagg.PathNet_homo
    torch.manual_seed(7)
    m = cls(Fd, H, Cc, L, math_mode=mode, pool=pool)
    if variant == "homo" and mode == "tc":     # keep ReLU inputs away from zero (pure precision test, as in test_gpu_tc.py)
        with torch.no_grad():
            m.fc0.bias += 4.0
            for lin in m.nets:
                lin.weight *= 0.1; lin.bias.fill_(1.0)
    g = torch.Generator().manual_seed(11)
    X = torch.randn(N, Fd, generator=g)
    mask = torch.zeros(N, dtype=torch.bool); mask[torch.randperm(N, generator=g)[:S]] = True
    neis, lt = torch.randint(0, N, (S, nw * L), generator=g), torch.randint(0, L, (S, nw, L), generator=g)
    params = {k: v.detach().clone().double().requires_grad_(True) for k, v in m.state_dict().items()}
    y_ref = po.pagg_forward_pool_variant(variant, params, X.double(), neis, nw, L, mask, lt, pool)
    G = torch.randn(S, Cc)
    (y_ref * G.double()).sum().backward()
    m = m.to(DEV).eval()
    y = m(X.to(DEV), neis, nw, L, mask, lt, None)
    (y * G.to(DEV)).sum().backward()
    if mode == "fp32":
        torch.testing.assert_close(y.cpu(), y_ref.detach().float(), **FWD_TOL)
    else:
        assert_tc_logits(y.cpu(), y_ref.detach().float())
    for k, q in m.named_parameters():
        if k.startswith("attw"):
            assert params[k].grad is None and float(q.grad.abs().max()) == 0.0      # no attention: its parameters get zero gradient
            continue
        want = params[k].grad.float()
        if mode == "fp32":
            torch.testing.assert_close(q.grad.cpu(), want, **GRAD_TOL, msg=lambda s: f"{k}: {s}")
        elif pool == "max":   # max pooling routes each gradient to ONE path: an argmax that flips under fp16 rounding moves a
            g = q.grad.cpu().double().flatten()   # whole row's worth of gradient, so hold the direction, not the entries
            cos = torch.dot(g, want.double().flatten()) / (g.norm() * want.double().norm() + 1e-30)
            assert torch.isfinite(g).all() and cos > 0.99, f"{k}: cosine to the oracle gradient {cos:.4f}"
        else:
            assert_grad_close(q.grad.cpu(), want, k, 1e-2, 1e-2)


def test_variant_argument_errors():
    import ctypes as C
    kw = dict(N=10, F=4, H=128, C=2, S=1, nw=1, L=2, training=0, dropout_p=0.0, dropout_seed=0, no_backward=0, row_chunk=0,
              dropout_seed_dev=None, x16=None)
    bytes_ = lambda **o: _lib.lib().pn_pagg_workspace_bytes(C.byref(_lib.PaggCfg(**{**kw, **o})))
    assert bytes_(variant=1, math_mode=1, cell=0, pool=1) > 0
    assert bytes_(variant=2, math_mode=0, cell=1, pool=1) > 0
    assert bytes_(variant=2, math_mode=0, cell=1, pool=0) == 0        # plain variant has no attention
    assert bytes_(variant=1, math_mode=1, cell=1, pool=0) == 0        # tensor-core kernels: LSTM only
    assert bytes_(variant=1, math_mode=0, cell=2, pool=0) > 0         # GRU (fp32 math mode)
    assert bytes_(variant=1, math_mode=1, cell=2, pool=0) == 0        # tensor-core kernels: LSTM only
    assert bytes_(variant=1, math_mode=0, cell=3, pool=0) == 0


@pytest.mark.parametrize("variant", ["hetero", "homo"])
@pytest.mark.parametrize("H", [32, 64, 128])
def test_gru_cell_matches_oracle(variant, H):
    """cell="gru" (nn.GRU under the `LSTM.*` keys): logits and every parameter gradient against the fp64 oracle, in
    training mode with the kernels' own dropout masks handed to nobody -- dropout off, like the pooling ablations."""
    N, Fd, Cc, nw, L, S = 130, 18, 5, 6, 4, 47
    cls = pagg.PathNet if variant == "hetero" else pagg.PathNet_homo
    torch.manual_seed(9)
    m = cls(Fd, H, Cc, L, cell="gru")                 # math_mode "auto" -> fp32 for a non-LSTM cell
    assert m.LSTM.weight_ih_l0.shape == (3 * H, H)
    g = torch.Generator().manual_seed(13)
    X = torch.randn(N, Fd, generator=g)
    mask = torch.zeros(N, dtype=torch.bool); mask[torch.randperm(N, generator=g)[:S]] = True
    neis, lt = torch.randint(0, N, (S, nw * L), generator=g), torch.randint(0, L, (S, nw, L), generator=g)
    params = {k: v.detach().clone().double().requires_grad_(True) for k, v in m.state_dict().items()}
    y_ref = po.pagg_forward_cell_variant(variant, params, X.double(), neis, nw, L, mask, lt, "gru")
    G = torch.randn(S, Cc)
    (y_ref * G.double()).sum().backward()
    m = m.to(DEV).eval()
    y = m(X.to(DEV), neis, nw, L, mask, lt, None)
    (y * G.to(DEV)).sum().backward()
    torch.testing.assert_close(y.cpu(), y_ref.detach().float(), **FWD_TOL)
    for k, q in m.named_parameters():
        torch.testing.assert_close(q.grad.cpu(), params[k].grad.float(), **GRAD_TOL, msg=lambda s: f"{k}: {s}")


def test_gru_cell_is_refused_by_the_tensor_core_mode():
    m = pagg.PathNet(18, 128, 5, 4, cell="gru", math_mode="tc").to(DEV)
    X = torch.randn(50, 18, device=DEV)
    mask = torch.zeros(50, dtype=torch.bool); mask[:7] = True
    with pytest.raises(Exception, match="LSTM cell only"):
        m(X, torch.randint(0, 50, (7, 12)), 3, 4, mask, torch.randint(0, 4, (7, 3, 4)), None)
