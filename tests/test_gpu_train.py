"""GPU tests of the device-side training step (pathnet_b200.train.TrainStep, pn_ce_loss_grad, pn_adam_step):
the same arithmetic as the reference's loop body (PathNet_run.py:345-352) done with torch ops."""
import ctypes as C

import numpy as np
import pytest
import torch

from pathnet_b200 import _lib, pagg
from pathnet_b200.train import TrainStep
from tests.test_gpu_pagg import _random_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("S,Cc,ldt", [(1, 2, torch.int64), (37, 3, torch.int64), (1000, 7, torch.int32), (9463, 3, torch.int64)])
def test_ce_loss_and_gradient_match_torch(S, Cc, ldt):
    g = torch.Generator().manual_seed(S + Cc)
    logits = (torch.randn(S, Cc, generator=g) * 3).to(DEV).requires_grad_(True)
    y = torch.randint(0, Cc, (S,), generator=g).to(DEV)
    want = torch.nn.functional.cross_entropy(logits, y, reduction="sum") / (2 * S)   # inv_count = 1 / (global rows) with 2 ranks
    want.backward()
    loss = torch.zeros(1, device=DEV)
    dl = torch.empty(S, Cc, device=DEV)
    yl = y.to(ldt)
    _lib.check(_lib.lib().pn_ce_loss_grad(S, Cc, logits.data_ptr(), yl.data_ptr(), pagg.IDX_DTYPE[ldt], 1.0 / (2 * S),
                                          loss.data_ptr(), dl.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.testing.assert_close(loss[0], want.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dl, logits.grad, rtol=1e-5, atol=1e-7)


def test_adam_step_matches_torch_adam():
    n = 10_007
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(n, generator=g).to(DEV)
    grads = [torch.randn(n, generator=g).to(DEV) * 10.0 ** (-k) for k in range(4)]
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=0.005, weight_decay=0.0005)
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    state = torch.zeros(2, dtype=torch.int64, device=DEV)
    cfg = _lib.AdamCfg(lr=0.005, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0005)
    st = torch.cuda.current_stream().cuda_stream
    for gk in grads:
        ref.grad = gk.clone()
        opt.step()
        _lib.check(_lib.lib().pn_step_advance(state.data_ptr(), st))
        _lib.check(_lib.lib().pn_adam_step(n, p.data_ptr(), gk.data_ptr(), m.data_ptr(), v.data_ptr(), C.byref(cfg), state.data_ptr(), st))
        torch.testing.assert_close(p, ref.detach(), rtol=2e-5, atol=1e-7)
    assert int(state[1]) == 4 and int(state[0]) != 0


def _reference_loop(variant, m0, X, mask, y, batches, nw, L, mode, steps):
    """the reference's loop body with torch's CE and Adam around OUR module (autograd path)"""
    m = type(m0)(m0.feature_length, m0.hidden_size, m0.out_size, L, math_mode=mode)
    m.load_state_dict(m0.state_dict())
    m = m.to(DEV).train()
    opt = torch.optim.Adam(m.parameters(), lr=0.005, weight_decay=0.0005)
    lossf = torch.nn.CrossEntropyLoss()
    losses = []
    for i in range(steps):
        neis, lt = batches[i % len(batches)]
        out = m(X.to(DEV), neis, nw, L, mask, lt, None)
        loss = lossf(out, y[mask].to(DEV))
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return m, losses


@pytest.mark.parametrize("variant", ["homo", "hetero"])
@pytest.mark.parametrize("mode,graph,idt", [("fp32", False, torch.int64), ("fp32", True, torch.int32), ("tc", True, torch.int64), ("tc3", True, torch.int32)])
def test_train_step_equals_module_plus_torch_optimizer(variant, mode, graph, idt):
    """dropout 0 (deterministic): K steps of TrainStep == K steps of module.forward + CrossEntropyLoss + backward +
    torch.optim.Adam — losses and final parameters."""
    N, Fd, H, Cc, nw, L, S, K = 150, 20, 128, 4, 6, 4, 70, 4
    m0, X, mask, neis, lt = _random_case(variant, N, Fd, H, Cc, nw, L, S, seed=31)
    y = torch.randint(0, Cc, (N,), generator=torch.Generator().manual_seed(2))
    g = torch.Generator().manual_seed(9)
    batches = [(neis, lt), (torch.randint(0, N, (S, nw * L), generator=g), torch.randint(0, L, (S, nw, L), generator=g))]
    ref, ref_losses = _reference_loop(variant, m0, X, mask, y, batches, nw, L, mode, K)
    m = type(m0)(Fd, H, Cc, L, math_mode=mode, dropout=0.0)
    m.load_state_dict(m0.state_dict())
    m = m.to(DEV).train()
    ts = TrainStep(m, X.to(DEV), mask, y[mask], nw, L, lr=0.005, weight_decay=0.0005, graph=graph)
    losses = []
    for i in range(K):
        a, b = batches[i % 2]
        ldt = torch.uint8 if idt == torch.int32 else torch.int64
        losses.append(ts.step(a.to(idt), b.to(ldt)).item())     # host tensors in, loss out
    tol = dict(rtol=1e-3, atol=1e-4) if mode in ("fp32", "tc3") else dict(rtol=2e-2, atol=2e-3)
    torch.testing.assert_close(torch.tensor(losses), torch.tensor(ref_losses), **tol)
    for (k, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        # Adam normalises the step: after K steps parameters differ by at most ~K * lr where gradient signs wobble
        torch.testing.assert_close(p, q, rtol=0, atol=(2e-4 if mode in ("fp32", "tc3") else 4 * 0.005 * K), msg=lambda s: f"{k}: {s}")
    sd = m.state_dict()
    assert set(sd) == set(ref.state_dict())                      # same checkpoint surface, parameters are views of the flat buffer
    assert sd["fc0.weight"].data_ptr() == ts.params.flat.data_ptr()


def test_train_step_graph_replay_draws_fresh_dropout_masks_and_learns():
    N, Fd, H, Cc, nw, L, S = 200, 16, 128, 3, 8, 4, 120
    m, X, mask, neis, lt = _random_case("homo", N, Fd, H, Cc, nw, L, S, seed=5)
    m.dropout = 0.5
    m = m.to(DEV).train()
    y = (X[:, 0] > 0).long() + (X[:, 1] > 0).long()              # learnable labels
    ts = TrainStep(m, X.to(DEV), mask, y[mask], nw, L, lr=0.01, weight_decay=0.0, math_mode="tc", seed=123, graph=True)
    frozen = ts.adam.lr
    ts.adam.lr = 0.0                                               # no learning: identical inputs and weights
    l0 = [ts.step(neis, lt).item() for _ in range(4)]            # -> differences come from the dropout masks only
    assert len(set(l0)) == 4, l0                                  # eager first step + 3 replays, all masks differ
    # (the captured graph holds lr by value; rebuild for the learning run)
    ts2 = TrainStep(m, X.to(DEV), mask, y[mask], nw, L, lr=frozen, weight_decay=0.0, math_mode="tc", seed=124, graph=True)
    l1 = [ts2.step(neis, lt).item() for _ in range(60)]
    assert np.mean(l1[-5:]) < 0.8 * np.mean(l1[:5]), (l1[:5], l1[-5:])
    m.eval()
    with torch.no_grad():
        out = m(X.to(DEV), neis, nw, L, mask, lt, None)           # the module (views of the flat buffer) sees the trained weights
    assert torch.isfinite(out).all()


def test_train_step_reports_bad_indices():
    m, X, mask, neis, lt = _random_case("homo", 60, 8, 32, 3, 4, 3, 25, seed=7)
    m = m.to(DEV).train()
    ts = TrainStep(m, X.to(DEV), mask, torch.zeros(25, dtype=torch.long), 4, 3, graph=False)
    bad = neis.clone(); bad[0, 0] = 60
    ts.step(bad, lt)
    torch.cuda.synchronize()
    with pytest.raises(IndexError):
        ts.step(neis, lt)
