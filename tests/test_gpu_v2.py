"""GPU tests of the round-2 surface of the PAGG operator: index dtypes and range checks of
pn_pagg_canonicalize, row-chunked calls, autograd robustness, the flat gradient buffer, and the
tensor-core mode held DIRECTLY to the reference-made fixtures and to the oracle at the benchmarked
(pubmed) configuration in train mode."""
import numpy as np
import pytest
import torch

from oracle import pagg_oracle as po
from pathnet_b200 import _lib, pagg
from tests.conftest import GOLDEN
from tests.test_gpu_pagg import FWD_TOL, GRAD_TOL, _fixture, _module_for, _oracle_run, _random_case, np_dropout_hash
from tests.test_gpu_tc import TC_GRAD_LOOSE, assert_grad_close, assert_tc_logits

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


# ------------------------------------------------------------------ canonicalize: dtypes + range check
@pytest.mark.parametrize("variant", ["hetero", "homo"])
@pytest.mark.parametrize("ndt,tdt", [(torch.int32, torch.uint8), (torch.int32, torch.int32), (torch.int64, torch.uint8),
                                     (torch.int32, torch.int64)])
def test_canonicalize_accepts_sampler_dtypes(variant, ndt, tdt):
    """int32 node ids / uint8 labels (what the sampler emits) canonicalise exactly like the reference's int64 tensors."""
    S, nw, L = 23, 7, 5
    g = torch.Generator().manual_seed(3)
    neis = torch.randint(0, 1000, (S, nw * L), generator=g)
    lt = torch.randint(0, L, (S, nw, L), generator=g)
    sn, st, pt = pagg.canonicalize(variant, neis.to(DEV, dtype=ndt), lt.to(DEV, dtype=tdt), nw, L, N=1000)
    o_sn, o_st, o_pt = po.canonicalize(variant, neis.numpy(), lt.numpy(), nw, L)
    assert (sn.cpu().numpy() == o_sn).all() and (st.cpu().numpy() == o_st).all() and (pt.cpu().numpy() == o_pt).all()


@pytest.mark.parametrize("variant", ["hetero", "homo"])
def test_out_of_range_indices_are_sanitised_and_reported(variant):
    """A node id outside [0,N) or a label outside [0,L): the reference raises (IndexError on CPU, a device-side assert
    on CUDA).  Here the entry becomes 0 — nothing reads or scatters out of bounds — and IndexError surfaces
    asynchronously: at check_indices(), or at the module's next forward."""
    N, Fd, H, Cc, nw, L, S = 50, 8, 32, 3, 4, 3, 20
    m, X, mask, neis, lt = _random_case(variant, N, Fd, H, Cc, nw, L, S, seed=4)
    m = m.to(DEV).eval()
    Xd = X.to(DEV)
    y_ok = m(Xd, neis, nw, L, mask, lt, None)
    m.check_indices()                      # clean inputs: nothing raised
    bad = neis.clone(); bad[3, 5] = N + 7; bad[0, 0] = -2
    y = m(Xd, bad, nw, L, mask, lt, None)
    y.square().sum().backward()            # backward scatters through the sanitised indices
    torch.cuda.synchronize()
    assert torch.isfinite(y).all() and all(torch.isfinite(p.grad).all() for p in m.parameters())
    with pytest.raises(IndexError, match="node id"):
        m.check_indices()
    m.check_indices()                      # the report is consumed
    badt = lt.clone(); badt[1, 2, 1] = L
    m(Xd, neis, nw, L, mask, badt, None)
    torch.cuda.synchronize()
    with pytest.raises(IndexError, match="label"):
        m(Xd, neis, nw, L, mask, lt, None)   # deferred report: raised by the NEXT call
    y2 = m(Xd, neis, nw, L, mask, lt, None)
    torch.testing.assert_close(y2, y_ok, rtol=0, atol=0)


def test_int32_uint8_inputs_through_the_module_match_int64():
    m, X, mask, neis, lt = _random_case("homo", 120, 16, 64, 4, 5, 4, 60, seed=6)
    m = m.to(DEV).eval()
    y64 = m(X.to(DEV), neis, 5, 4, mask, lt, None)
    y32 = m(X.to(DEV), neis.to(DEV, dtype=torch.int32), 5, 4, mask, lt.to(DEV, dtype=torch.uint8), None)   # device tensors, sampler dtypes
    torch.testing.assert_close(y32, y64, rtol=0, atol=0)


# ------------------------------------------------------------------ row-chunked calls
@pytest.mark.parametrize("variant", ["hetero", "homo"])
@pytest.mark.parametrize("mode,H", [("fp32", 64), ("fp32", 128), ("tc", 128), ("tc3", 128)])
@pytest.mark.parametrize("training", [False, True])
def test_row_chunked_call_equals_whole_call(variant, mode, H, training):
    """row_chunk > 0: forward loops over row chunks sharing one Z / T table, backward recomputes each chunk's forward
    (same dropout masks: mask indices are global).  Logits are identical to the whole call (every sequence's
    arithmetic is the same), gradients agree up to the order of the atomic sums."""
    N, Fd, Cc, nw, L, S, p, seed = 300, 24, 5, 7, 4, 157, 0.5, 0x1234ABCD
    m, X, mask, neis, lt = _random_case(variant, N, Fd, H, Cc, nw, L, S, seed=12)
    sn, st, pt = po.canonicalize(variant, neis.numpy(), lt.numpy(), nw, L)
    rows = torch.nonzero(mask).reshape(-1).numpy()
    m = m.to(DEV)
    dev = lambda a, dt: torch.tensor(a, device=DEV, dtype=dt)
    G = torch.randn(S, Cc, device=DEV)
    out = {}
    for chunk in (0, 50, 157, 1):
        if chunk == 1 and mode == "fp32" and H == 128:
            continue   # (one-row chunks are exercised by the other two combinations)
        m.zero_grad()
        y = pagg.pagg_forward(variant, X.to(DEV), dict(m.named_parameters()), dev(sn, torch.int32), dev(st, torch.uint8),
                              dev(pt, torch.int32), dev(rows, torch.int32), dropout=p, training=training, seed=seed,
                              math_mode=mode, row_chunk=chunk)
        (y * G).sum().backward()
        torch.cuda.synchronize()
        out[chunk] = (y.detach().clone(), {k: q.grad.clone() for k, q in m.named_parameters()})
    y0, g0 = out[0]
    for chunk, (y, g) in out.items():
        if chunk == 0:
            continue
        torch.testing.assert_close(y, y0, rtol=0, atol=0, msg=lambda s: f"row_chunk={chunk}: {s}")
        for k in g0:
            if mode in ("tc", "tc3"):   # the per-chunk power-of-two gradient scale differs -> fp16 rounding differs slightly
                assert_grad_close(g[k], g0[k], f"row_chunk={chunk} {k}", 5e-3, 5e-3)
            else:
                torch.testing.assert_close(g[k], g0[k], rtol=2e-3, atol=2e-5, msg=lambda s: f"row_chunk={chunk} {k}: {s}")


def test_row_chunk_bounds_the_workspace():
    import ctypes as C
    kw = dict(N=10_000_000, F=128, H=128, C=8, S=4_800_000, nw=64, L=8, variant=1, math_mode=1, training=1, dropout_p=0.7,
              dropout_seed=1, no_backward=0, dropout_seed_dev=None, x16=None)
    whole = _lib.lib().pn_pagg_workspace_bytes(C.byref(_lib.PaggCfg(row_chunk=0, **kw)))
    chunked = _lib.lib().pn_pagg_workspace_bytes(C.byref(_lib.PaggCfg(row_chunk=4096, **kw)))
    assert whole > 10 ** 13 and chunked < 140 * 2 ** 30       # BASELINE configs[4]: TBs whole, fits one B200 (180 GB) chunked


# ------------------------------------------------------------------ autograd robustness
def test_second_backward_and_inplace_updates_raise_clearly():
    m, X, mask, neis, lt = _random_case("homo", 80, 12, 32, 3, 4, 3, 30, seed=2)
    m = m.to(DEV).eval()
    y = m(X.to(DEV), neis, 4, 3, mask, lt, None)
    loss = y.square().sum()
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second time"):
        loss.backward()
    m.zero_grad()
    y = m(X.to(DEV), neis, 4, 3, mask, lt, None)
    with torch.no_grad():
        m.fc0.weight.add_(1.0)             # optimizer-style in-place update between forward and backward
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        y.sum().backward()


def test_parameter_grads_alias_one_flat_buffer():
    """pn_pagg_backward writes into slices of ONE flat buffer and autograd installs those slices as .grad: the
    data-parallel all-reduce runs on it in place (pathnet_b200.dist.FlatGradAllReduce zero-copy path)."""
    from pathnet_b200.dist import FlatGradAllReduce
    m, X, mask, neis, lt = _random_case("hetero", 80, 12, 128, 3, 4, 4, 30, seed=2)
    m = m.to(DEV).train()
    red = FlatGradAllReduce(m.parameters(), sink=m.grad_sink)
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        m(X.to(DEV), neis, 4, 4, mask, lt, None).square().sum().backward()
        assert red._aliases_sink()
        flat = m.grad_sink.flat
        assert sum(p.numel() for p in m.parameters()) == m.grad_sink.payload
        before = {k: p.grad.clone() for k, p in m.named_parameters()}
        flat.mul_(2.0)                     # what an in-place all-reduce over 2 identical ranks would do
        for k, p in m.named_parameters():
            torch.testing.assert_close(p.grad, 2 * before[k], rtol=0, atol=0)


# ------------------------------------------------------------------ tensor-core mode, directly against the reference
H128_FIXTURES = sorted(GOLDEN.glob("pagg_*_h128.npz"))


@pytest.mark.parametrize("path", H128_FIXTURES, ids=lambda p: p.stem)
def test_tc_mode_matches_reference_fixture(path):
    """The tensor-core mode against what the REFERENCE's own classes produced (fp16-exact inputs and weights, so the
    only differences are fp16 rounding of intermediate activations, tanh.approx and the summation order)."""
    z, params, grads = _fixture(path)
    variant, nw, L = str(z["variant"]), int(z["nw"]), int(z["L"])
    m = _module_for(variant, int(z["F"]), int(z["H"]), int(z["C"]), L, params, math_mode="tc").eval()
    X = torch.tensor(z["X"].astype(np.float32), device=DEV)
    neis, lt = torch.tensor(z["neis"].astype(np.int64)), torch.tensor(z["layer_type"].astype(np.int64))
    y = m(X, neis, nw, L, z["mask"], lt, None)
    want = torch.tensor(z["logits"])
    print(f"tc vs reference fixture: max abs err {(y.cpu() - want).abs().max().item():.3e} (max |logit| {want.abs().max().item():.3f})")
    assert_tc_logits(y.cpu(), want)
    (y * torch.tensor(z["G"].astype(np.float32), device=DEV)).sum().backward()
    tol = TC_GRAD_LOOSE if variant == "homo" else (None, None)
    worst = max(assert_grad_close(p.grad.cpu(), grads[k], k, *tol) for k, p in m.named_parameters())
    print(f"tc vs reference fixture: worst relative Frobenius grad err {worst:.3e}")


def _hash_masks(seed, S, nw, L, H, p):
    """dropout masks of the kernels (canonical order), as float32 tensors incl. the 1/(1-p) scale; hashed in slabs"""
    thr, scale = int(p * 65536), np.float32(1.0 / (1.0 - p))
    total = S * nw * L * H
    ms = np.empty(total, dtype=np.float32)
    step = 1 << 24
    for lo in range(0, total, step):
        hi = min(total, lo + step)
        ms[lo:hi] = (np_dropout_hash(seed, 0, np.arange(lo, hi, dtype=np.uint64)) >= thr) * scale
    mc = (np_dropout_hash(seed, 1, np.arange(S * 2 * H, dtype=np.uint64)) >= thr) * scale
    return torch.from_numpy(ms).view(S, nw, L, H), torch.from_numpy(mc.astype(np.float32)).view(S, 2 * H)


def test_benchmarked_config_train_mode_against_oracle():
    """THE benchmarked configuration (pubmed-shaped: N = 19 717, F = 500, H = 128, nw = 40, L = 4, S = 9 463 rows,
    P = 378 520 paths, PathNet_homo, train mode, dropout 0.7 with the kernels' own hash masks) compared with the
    CPU oracle directly — logits and all twelve parameter gradients — in all three math modes."""
    N, Fd, H, Cc, nw, L, S, p, seed = 19717, 500, 128, 3, 40, 4, 9463, 0.7, 0x5EED0BE0C4
    m, X, mask, neis, lt = _random_case("homo", N, Fd, H, Cc, nw, L, S, seed=9)
    sn, st, pt = po.canonicalize("homo", neis.numpy(), lt.numpy(), nw, L)
    rows = torch.nonzero(mask).reshape(-1).numpy()
    ms, mc = _hash_masks(seed, S, nw, L, H, p)
    torch.set_num_threads(min(32, torch.get_num_threads()))
    params = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}   # fp32 CPU = the reference's arithmetic
    y_ref = po.pagg_forward_canonical("homo", params, X, sn, st, pt, rows, mask_seq=ms, mask_cat=mc)
    G = torch.randn(S, Cc) / S
    (y_ref * G).sum().backward()
    del ms
    m = m.to(DEV).train()
    dev = lambda a, dt: torch.tensor(a, device=DEV, dtype=dt)
    idx = (dev(sn, torch.int32), dev(st, torch.uint8), dev(pt, torch.int32), dev(rows, torch.int32))
    for mode in ("fp32", "tc3", "tc"):
        m.zero_grad()
        y = pagg.pagg_forward("homo", X.to(DEV), dict(m.named_parameters()), *idx, dropout=p, training=True, seed=seed,
                              math_mode=mode)
        err = (y.cpu() - y_ref.detach()).abs().max().item()
        print(f"[{mode}] pubmed train-mode logits: max abs err {err:.3e} (max |logit| {y_ref.abs().max().item():.3f})")
        (y * G.to(DEV)).sum().backward()
        if mode in ("fp32", "tc3"):   # the split-precision tensor-core mode is held to the strict mode's bar
            torch.testing.assert_close(y.cpu(), y_ref.detach(), rtol=2e-4, atol=5e-5)   # two fp32 summation orders over 378 520 paths
            worst = max(assert_grad_close(q.grad.cpu(), params[k].grad, k, 2e-3, 1e-3) for k, q in m.named_parameters())
        else:
            assert_tc_logits(y.cpu(), y_ref.detach())
            worst = max(assert_grad_close(q.grad.cpu(), params[k].grad, k, *TC_GRAD_LOOSE) for k, q in m.named_parameters())
        print(f"[{mode}] pubmed train-mode worst relative Frobenius grad err {worst:.3e}")
