"""GPU tests of the split-precision tensor-core math mode (math_mode="tc3", PN_MATH_TC3).

The LSTM runs on tcgen05 tensor cores with every operand carried as two fp16 terms (hi + lo) and three tensor-core
passes per product, gates on ex2 / rcp; the projections run on the strict fp32 kernels.  The mode is held to the SAME
bar as the strict fp32 mode (tests/test_gpu_pagg.py): logits rtol 1e-4 / atol 2e-5, every parameter gradient
rtol 2e-3 / atol 2e-5 — against the reference-made fixtures, the fp64 oracle, and (train mode, the kernels' own dropout
masks) the benchmarked pubmed-shaped configuration."""
import numpy as np
import pytest
import torch

from oracle import pagg_oracle as po
from pathnet_b200 import pagg
from tests.conftest import GOLDEN
from tests.test_gpu_pagg import FWD_TOL, GRAD_TOL, _fixture, _module_for, _oracle_run, _random_case, np_dropout_hash

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
H128_FIXTURES = sorted(GOLDEN.glob("pagg_*_h128.npz"))


def _report(tag, y, y_ref, grads, g_ref):
    err = (y.double() - y_ref.double()).abs().max().item()
    worst = max((grads[k].double() - g_ref[k].double()).norm().item() / max(g_ref[k].double().norm().item(), 1e-30) for k in grads)
    print(f"[tc3 {tag}] logits max abs err {err:.3e} (max |logit| {y_ref.abs().max().item():.3f}); worst relative Frobenius grad err {worst:.3e}")


@pytest.mark.parametrize("path", H128_FIXTURES, ids=lambda p: p.stem)
def test_tc3_matches_reference_fixture_at_the_fp32_bar(path):
    """Logits and all gradients against what the REFERENCE's own classes produced (tests/golden/make_pagg_golden.py)."""
    z, params, grads = _fixture(path)
    variant, nw, L = str(z["variant"]), int(z["nw"]), int(z["L"])
    m = _module_for(variant, int(z["F"]), int(z["H"]), int(z["C"]), L, params, math_mode="tc3").eval()
    X = torch.tensor(z["X"].astype(np.float32), device=DEV)
    neis, lt = torch.tensor(z["neis"].astype(np.int64)), torch.tensor(z["layer_type"].astype(np.int64))
    y = m(X, neis, nw, L, z["mask"], lt, None)
    (y * torch.tensor(z["G"].astype(np.float32), device=DEV)).sum().backward()
    _report(path.stem, y.detach().cpu(), torch.tensor(z["logits"]), {k: p.grad.cpu() for k, p in m.named_parameters()}, grads)
    torch.testing.assert_close(y.detach().cpu(), torch.tensor(z["logits"]), **FWD_TOL)
    for k, p in m.named_parameters():
        torch.testing.assert_close(p.grad.cpu(), grads[k], **GRAD_TOL, msg=lambda s: f"{k}: {s}")


@pytest.mark.parametrize("variant", ["homo", "hetero"])
@pytest.mark.parametrize("N,Fd,Cc,nw,L,S", [(183, 50, 5, 40, 4, 87), (300, 33, 7, 5, 6, 129), (64, 16, 3, 1, 1, 64), (500, 40, 4, 3, 8, 200)])
def test_tc3_forward_backward_match_oracle(variant, N, Fd, Cc, nw, L, S):
    """Natural (default-init) weights, both variants, L = 1 .. 8, ragged last tile — against the fp64 oracle."""
    H = 128
    m, X, mask, neis, lt = _random_case(variant, N, Fd, H, Cc, nw, L, S, seed=N + S + 3)
    G = torch.randn(S, Cc)
    y_ref, g_ref = _oracle_run(variant, m, X, mask, neis, lt, nw, L, G)
    m = m.to(DEV).eval()
    m.math_mode = "tc3"
    y = m(X.to(DEV), neis, nw, L, mask, lt, None)
    (y * G.to(DEV)).sum().backward()
    _report(f"{variant} N={N} L={L}", y.detach().cpu(), y_ref, {k: p.grad.cpu() for k, p in m.named_parameters()}, g_ref)
    torch.testing.assert_close(y.detach().cpu(), y_ref, **FWD_TOL)
    for k, p in m.named_parameters():
        torch.testing.assert_close(p.grad.cpu(), g_ref[k], **GRAD_TOL, msg=lambda s: f"{k}: {s}")


@pytest.mark.parametrize("variant", ["homo", "hetero"])
def test_tc3_training_dropout_matches_oracle_with_same_masks(variant):
    N, Fd, H, Cc, nw, L, S, p, seed = 150, 21, 128, 4, 6, 4, 70, 0.7, 0xABCDEF0123
    m, X, mask, neis, lt = _random_case(variant, N, Fd, H, Cc, nw, L, S, seed=8)
    thr, scale = int(p * 65536), 1.0 / (1.0 - p)
    keep_seq = np_dropout_hash(seed, 0, np.arange(S * nw * L * H, dtype=np.uint64)) >= thr
    keep_cat = np_dropout_hash(seed, 1, np.arange(S * 2 * H, dtype=np.uint64)) >= thr
    ms = torch.tensor(keep_seq.reshape(S, nw, L, H).astype(np.float64) * scale)
    mc = torch.tensor(keep_cat.reshape(S, 2 * H).astype(np.float64) * scale)
    sn, st, pt = po.canonicalize(variant, neis.numpy(), lt.numpy(), nw, L)
    rows = torch.nonzero(mask).reshape(-1).numpy()
    params = {k: v.detach().clone().double().requires_grad_(True) for k, v in m.state_dict().items()}
    y_ref = po.pagg_forward_canonical(variant, params, X.double(), sn, st, pt, rows, mask_seq=ms, mask_cat=mc)
    G = torch.randn(S, Cc)
    (y_ref * G.double()).sum().backward()
    m = m.to(DEV).train()
    dev = lambda a, dt: torch.tensor(a, device=DEV, dtype=dt)
    y = pagg.pagg_forward(variant, X.to(DEV), dict(m.named_parameters()), dev(sn, torch.int32), dev(st, torch.uint8),
                          dev(pt, torch.int32), dev(rows, torch.int32), dropout=p, training=True, seed=seed, math_mode="tc3")
    (y * G.to(DEV)).sum().backward()
    _report(f"{variant} train", y.detach().cpu(), y_ref.detach().float(), {k: q.grad.cpu() for k, q in m.named_parameters()},
            {k: params[k].grad.float() for k in params})
    torch.testing.assert_close(y.detach().cpu(), y_ref.detach().float(), **FWD_TOL)
    for k, pr in m.named_parameters():
        torch.testing.assert_close(pr.grad.cpu(), params[k].grad.float(), **GRAD_TOL, msg=lambda s: f"{k}: {s}")


def test_tc3_is_deterministic_in_the_forward_and_close_to_fp32_mode_at_pubmed_scale():
    """Full BASELINE size (P = 378 520 sequences): two forward calls are bit-identical (no atomics in the forward) and the
    logits agree with the strict fp32 mode at its own tolerance."""
    N, Fd, H, Cc, nw, L, S = 19717, 500, 128, 3, 40, 4, 9463
    m, X, mask, neis, lt = _random_case("homo", N, Fd, H, Cc, nw, L, S, seed=9)
    m = m.to(DEV).eval()
    Xd = X.to(DEV)
    with torch.no_grad():
        y32 = m(Xd, neis, nw, L, mask, lt, None)
        m.math_mode = "tc3"
        y3a = m(Xd, neis, nw, L, mask, lt, None)
        y3b = m(Xd, neis, nw, L, mask, lt, None)
    assert torch.equal(y3a, y3b)
    print("tc3 vs fp32 mode at pubmed scale: max abs diff", (y3a - y32).abs().max().item(), "max |logit|", y32.abs().max().item())
    torch.testing.assert_close(y3a, y32, rtol=2e-4, atol=5e-5)


@pytest.mark.parametrize("flags,splitk", [(16, 1), (19, 1), (19, 5), (18, 1), (26, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (512, 128, 1000), (300, 77, 130), (1000, 512, 500)])
def test_split_precision_tcgen05_gemm_is_fp32_grade(M, N, K, flags, splitk):
    """The projections' GEMM in this mode: arbitrary fp32 operands (NOT fp16-representable), every operand layout the
    projections use, split-K and accumulate — against fp64.  Observed: 2-4x the error of a cuBLAS fp32 GEMM at K = 500 .. 1000
    (the tensor core's fp32 accumulator truncates, so the error grows with the length of the accumulation chain; with
    split-K 5 it equals cuBLAS'), i.e. ~4e-6 of max |C|: the bound below is 1e-5 relative + 1e-5 sqrt(K) absolute."""
    from pathnet_b200 import _lib
    g = torch.Generator().manual_seed(M + N + K + flags)
    a_mn, b_mn, acc = bool(flags & 1), bool(flags & 2), bool(flags & 8)
    A = torch.randn(M, K, generator=g) * 3.0
    B = torch.randn(N, K, generator=g) * 0.1
    C0 = torch.randn(M, N, generator=g) if acc else torch.zeros(M, N)
    want = A.double() @ B.double().t() + C0.double()
    Ad = (A.t().contiguous() if a_mn else A).to(DEV)
    Bd = (B.t().contiguous() if b_mn else B).to(DEV)
    C = C0.clone().to(DEV)
    _lib.check(_lib.lib().pn_debug_gemm_tc2(Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), M, N, K, flags, splitk,
                                            torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (C.cpu().double() - want).abs().max().item()
    ref32 = ((A.to(DEV) @ B.to(DEV).t()).cpu().double() + C0.double() - want).abs().max().item()   # what cuBLAS fp32 (no TF32) makes of it
    print(f"[split gemm M={M} N={N} K={K} flags={flags} splitk={splitk}] max abs err {err:.3e} (cuBLAS fp32: {ref32:.3e}, max |C| {want.abs().max().item():.2f})")
    torch.testing.assert_close(C.cpu().double(), want, rtol=1e-5, atol=1e-5 * float(K) ** 0.5)
