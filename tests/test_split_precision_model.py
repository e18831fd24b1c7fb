"""Host-side model of the arithmetic of the split-precision tensor-core mode (csrc/pagg_tc3.cu), in numpy:

* an fp32 value carried as two fp16 terms, hi = fp16(v), lo = fp16(v - hi);
* a product as three passes into an fp32 accumulator, A W ~= A_hi W_hi + A_lo W_hi + A_hi W_lo;
* the gate formulas: sigmoid / tanh through e^-x with the exponent capped at 2^40 and ONE reciprocal of the product of the
  denominators shared by i, f, g (and by o, tanh c).

No GPU: these are the identities and error bounds the kernels rely on (the kernels themselves are held to the fp32 parity bar in
tests/test_gpu_tc3.py)."""
import numpy as np


def split(v):
    v = v.astype(np.float32)
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def test_two_fp16_terms_carry_22_bits():
    rng = np.random.default_rng(0)
    sign = np.where(rng.random(200000) < 0.5, -1.0, 1.0)
    v = (sign * np.exp(rng.uniform(np.log(0.125), np.log(2000.0), 200000))).astype(np.float32)   # lo stays a normal fp16 number
    hi, lo = split(v)
    rec = hi.astype(np.float64) + lo.astype(np.float64)
    rel = np.abs(rec - v.astype(np.float64)) / np.abs(v.astype(np.float64))
    assert rel.max() <= 2.0 ** -21                                   # 11 + 11 significant bits (one lost to the sign of lo)
    small = (rng.standard_normal(100000) * 0.05).astype(np.float32)  # below 0.125 lo falls into fp16 subnormals (quantum 2^-24):
    hi, lo = split(small)                                            # the bound turns absolute -- 3e-8, far below any tolerance here
    assert np.abs(hi.astype(np.float64) + lo.astype(np.float64) - small.astype(np.float64)).max() <= 2.0 ** -25


def test_three_pass_product_is_fp32_grade():
    rng = np.random.default_rng(1)
    M, K, N = 64, 768, 64
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.uniform(-1, 1, (K, N)) / np.sqrt(128)).astype(np.float32)      # LSTM init range at H = 128
    exact = A.astype(np.float64) @ W.astype(np.float64)
    ah, al = split(A)
    wh, wl = split(W)
    f = lambda x: x.astype(np.float32)
    three = f(ah) @ f(wh) + f(al) @ f(wh) + f(ah) @ f(wl)                    # fp32 accumulation of exact fp16 x fp16 products
    one = f(ah) @ f(wh)                                                     # the fast mode's single pass
    scale = np.abs(A).astype(np.float64) @ np.abs(W).astype(np.float64)      # sum of |terms|: what rounding errors are relative to
    err3, err1 = np.abs(three - exact) / scale, np.abs(one - exact) / scale
    assert err3.max() < 2e-6 and err1.max() > 20 * err3.max()               # ~fp32 summation error vs ~2^-12 per operand
    dropped = np.abs(f(al) @ f(wl)) / scale                                  # the lo x lo term that is never computed
    assert dropped.max() < 2.0 ** -20


def _gates(ai, af, ag, ao, c_prev):
    """float32 restatement of the epilogue of k_lstm_fwd_tc3 (np.exp2 stands in for ex2.approx, 1/x for rcp.approx)"""
    f32 = np.float32
    l2e = f32(1.4426950408889634)
    exp_neg = lambda x: np.exp2(np.minimum(-l2e * x, f32(40.0)).astype(f32)).astype(f32)
    exp_neg2 = lambda x: np.exp2(np.minimum(f32(-2.0) * l2e * x, f32(40.0)).astype(f32)).astype(f32)
    di, df = f32(1) + exp_neg(ai), f32(1) + exp_neg(af)
    eg = exp_neg2(ag)
    dg, dov = f32(1) + eg, f32(1) + exp_neg(ao)
    r3 = f32(1) / (di * df * dg)
    gi, gf, gg = r3 * (df * dg), r3 * (di * dg), (f32(1) - eg) * (r3 * (di * df))
    cn = gf * c_prev + gi * gg
    ec = exp_neg2(cn)
    dc = f32(1) + ec
    r2 = f32(1) / (dov * dc)
    go = r2 * dc
    h = go * ((f32(1) - ec) * (r2 * dov))
    return gi, gf, gg, go, cn, h


def test_gate_formulas_with_shared_reciprocals():
    rng = np.random.default_rng(2)
    n = 400000
    pre = [(rng.standard_normal(n) * s).astype(np.float32) for s in (3, 3, 3, 3)]
    c_prev = (rng.standard_normal(n) * 2).astype(np.float32)
    gi, gf, gg, go, cn, h = _gates(*pre, c_prev)
    a = [x.astype(np.float64) for x in pre]
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    ri, rf, rg, ro = sig(a[0]), sig(a[1]), np.tanh(a[2]), sig(a[3])
    rc = rf * c_prev.astype(np.float64) + ri * rg
    rh = ro * np.tanh(rc)
    for got, want in ((gi, ri), (gf, rf), (gg, rg), (go, ro), (h, rh)):
        assert np.abs(got.astype(np.float64) - want).max() < 1e-6          # tanh.approx would sit at 5e-4 here
    assert np.abs(cn.astype(np.float64) - rc).max() < 4e-6
    # extremes: saturated gates, huge cell states -- the capped exponents keep every product of denominators finite
    big = np.array([-200, -90, -45, -30, 0, 30, 45, 90, 200], dtype=np.float32)
    grid = np.stack(np.meshgrid(big, big, big, big, big, indexing="ij"), -1).reshape(-1, 5)
    out = _gates(*(grid[:, k].copy() for k in range(5)))
    for x in out:
        assert np.isfinite(x).all()
    gi, gf, gg, go, cn, h = out
    assert (gi >= 0).all() and (gi <= 1).all() and (np.abs(gg) <= 1).all() and (np.abs(h) <= 1.0000001).all()
