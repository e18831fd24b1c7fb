"""PAGG operator — host-side mirror of the reference's `PathNet` /
`PathNet_homo` modules (/root/reference/PathNet_run.py:150-278).

Same constructor `(feature_length, hidden_size, out_size, wl, **kwargs)`, same
7-argument `forward(X, neis, num_w, walk_len, indices, layer_type, indxx)`,
same parameter names/shapes (so reference checkpoints such as
saved_models/cornell.pth load with `load_state_dict`), same train()/eval()
dropout behaviour.  All arithmetic runs in libpathnet_b200.so through the C
ABI (include/pathnet_b200.h); there is no PyTorch/CPU fallback — without a
CUDA device or the library, forward raises.

Two things the reference reads from module globals are attributes here:
`self.dropout` (reference global `dropout`, :71) and the device (taken from X).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from ._lib import PARAM_NAMES, PaggCfg, PaggParams, check, lib

VARIANT = {"hetero": 0, "homo": 1}
MATH = {"fp32": 0, "tc": 1}


def resolve_math_mode(mode: str, H: int, L: int) -> str:
    """"auto" (the default): tensor-core kernels where they exist (hidden 128, path_len <= 8), else the fp32
    CUDA-core kernels.  "fp32" = strict-parity mode (1e-4 against the reference's CPU fp32 math); "tc" =
    tensor cores or an error if the shape is unsupported."""
    if mode == "auto":
        return "tc" if (H == 128 and L <= 8) else "fp32"
    if mode not in MATH:
        raise ValueError(f"math_mode must be 'auto', 'fp32' or 'tc' (got {mode!r})")
    return mode


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def _params_struct(tensors) -> PaggParams:
    s = PaggParams()
    for k, t in zip(PARAM_NAMES, tensors):
        setattr(s, k, t.data_ptr())
    return s


def canonicalize(variant: str, neis: torch.Tensor, layer_type: torch.Tensor, num_w: int, walk_len: int):
    """Device int64 index tensors of the reference -> canonical (seq_node int32
    [S,nw,L], seq_type uint8 [S,nw,L], partner int32 [S,nw]) (pn_pagg_canonicalize)."""
    S = neis.shape[0]
    dev = neis.device
    neis = neis.contiguous().view(S, num_w * walk_len)
    layer_type = layer_type.contiguous()
    seq_node = torch.empty((S, num_w, walk_len), dtype=torch.int32, device=dev)
    seq_type = torch.empty((S, num_w, walk_len), dtype=torch.uint8, device=dev)
    partner = torch.empty((S, num_w), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().pn_pagg_canonicalize(VARIANT[variant], S, num_w, walk_len, neis.data_ptr(),
                                         layer_type.data_ptr(), seq_node.data_ptr(), seq_type.data_ptr(),
                                         partner.data_ptr(), _stream_ptr()))
    return seq_node, seq_type, partner


class _PaggFunction(torch.autograd.Function):
    """logits = PAGG(X; 12 parameter tensors) over canonical index tensors."""

    @staticmethod
    def forward(ctx, X, fc0_w, fc0_b, nets_w, nets_b, w_ih, w_hh, b_ih, b_hh, attw_w, attw_b, fc2_w,
                fc2_b, seq_node, seq_type, partner, rows, meta):
        params = [t.detach().contiguous().float() for t in
                  (fc0_w, fc0_b, nets_w, nets_b, w_ih, w_hh, b_ih, b_hh, attw_w, attw_b, fc2_w, fc2_b)]
        X = X.detach().contiguous().float()
        dev = X.device
        if dev.type != "cuda":
            raise RuntimeError("pathnet_b200 PAGG runs on CUDA only (no CPU fallback); X is on %s" % dev)
        S, nw, L = seq_node.shape
        H, Fd = params[0].shape
        cfg = PaggCfg(N=X.shape[0], F=Fd, H=H, C=params[10].shape[0], S=S, nw=nw, L=L,
                      variant=VARIANT[meta["variant"]], math_mode=MATH[resolve_math_mode(meta["math_mode"], H, L)],
                      training=int(meta["training"]), dropout_p=float(meta["dropout"]),
                      dropout_seed=int(meta["seed"]), no_backward=int(not meta.get("need_backward", True)))
        with torch.cuda.device(dev):
            ws_bytes = lib().pn_pagg_workspace_bytes(C.byref(cfg))
            if ws_bytes == 0:
                check(-1)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            logits = torch.empty((S, cfg.C), dtype=torch.float32, device=dev)
            ps = _params_struct(params)
            check(lib().pn_pagg_forward(C.byref(cfg), C.byref(ps), X.data_ptr(), seq_node.data_ptr(),
                                        seq_type.data_ptr(), partner.data_ptr(), rows.data_ptr(),
                                        ws.data_ptr(), logits.data_ptr(), _stream_ptr()))
        ctx.cfg, ctx.ws, ctx.params, ctx.X = cfg, ws, params, X
        ctx.idx = (seq_node, seq_type, partner, rows)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        cfg, ws, params, X = ctx.cfg, ctx.ws, ctx.params, ctx.X
        seq_node, seq_type, partner, rows = ctx.idx
        dlogits = dlogits.contiguous().float()
        grads = [torch.empty_like(p) for p in params]
        with torch.cuda.device(X.device):
            ps, gs = _params_struct(params), _params_struct(grads)
            check(lib().pn_pagg_backward(C.byref(cfg), C.byref(ps), X.data_ptr(), seq_node.data_ptr(),
                                         seq_type.data_ptr(), partner.data_ptr(), rows.data_ptr(),
                                         ws.data_ptr(), dlogits.data_ptr(), C.byref(gs), _stream_ptr()))
        ctx.ws = None
        return (None, *grads, None, None, None, None, None)


def pagg_forward(variant, X, params: dict, seq_node, seq_type, partner, rows, *, dropout=0.0,
                 training=False, seed=0, math_mode="auto"):
    """Functional form over a dict with the reference's state_dict keys."""
    L = seq_node.shape[2]
    nets_w = torch.stack([params[f"nets.{d}.weight"] for d in range(L)])
    nets_b = torch.stack([params[f"nets.{d}.bias"] for d in range(L)])
    need_bwd = torch.is_grad_enabled() and any(t.requires_grad for t in params.values() if torch.is_tensor(t))
    meta = dict(variant=variant, math_mode=math_mode, training=training, dropout=dropout, seed=seed, need_backward=need_bwd)
    return _PaggFunction.apply(X, params["fc0.weight"], params["fc0.bias"], nets_w, nets_b,
                               params["LSTM.weight_ih_l0"], params["LSTM.weight_hh_l0"],
                               params["LSTM.bias_ih_l0"], params["LSTM.bias_hh_l0"],
                               params["attw.weight"].reshape(-1), params["attw.bias"],
                               params["fc2.weight"], params["fc2.bias"], seq_node, seq_type, partner,
                               rows, meta)


class _PathNetBase(nn.Module):
    _variant = "hetero"

    def __init__(self, feature_length, hidden_size, out_size, wl, **kwargs):
        # kwargs mirrors the reference signature (`aggr` is set and ignored there, :156-157)
        self.dropout = float(kwargs.pop("dropout", 0.0))
        self.math_mode = kwargs.pop("math_mode", "auto")   # "auto" | "fp32" (strict parity) | "tc"
        kwargs.setdefault("aggr", "mean")
        super().__init__()
        self._rows_cache = {}
        self.feature_length, self.hidden_size, self.out_size = feature_length, hidden_size, out_size
        # construction order = the reference's (:162-170), so a given torch seed
        # yields the same initial weights
        self.fc0 = nn.Linear(feature_length, hidden_size)
        self.LSTM = nn.LSTM(hidden_size, hidden_size)
        self.fc2 = nn.Linear(2 * hidden_size, out_size)
        self.nets = nn.ModuleList([nn.Linear(hidden_size, hidden_size) for _ in range(wl)])
        self.attw = nn.Linear(2 * hidden_size, 1)
        self.Lrelu = nn.LeakyReLU()

    def _apply(self, fn, *a, **k):
        # nn.LSTM would otherwise try cudnn weight flattening on .to(device); harmless
        return super()._apply(fn, *a, **k)

    def _rows_of(self, indices, dev):
        """Row numbers selected by the node mask `indices` (int32, on dev).  torch.nonzero synchronises with the device,
        and training passes the same mask tensor every step (PathNet_run.py:345-347), so the result is kept per mask
        TENSOR OBJECT (held alive, so its identity cannot be recycled) for as long as its in-place version counter is
        unchanged.  numpy masks are not cached (their mutations cannot be seen)."""
        if torch.is_tensor(indices):
            ent = self._rows_cache.get(id(indices))
            if ent is not None and ent[0] is indices and ent[1] == indices._version and ent[2] == str(dev):
                return ent[3]
            mask = indices.to(dev)
        else:
            mask = torch.as_tensor(np.asarray(indices), device=dev)
        rows = (torch.nonzero(mask.reshape(-1), as_tuple=False).reshape(-1) if mask.dtype == torch.bool
                else torch.nonzero(mask.reshape(-1) != 0, as_tuple=False).reshape(-1)).to(torch.int32)
        if torch.is_tensor(indices):
            if len(self._rows_cache) >= 8:   # train / val / test masks and a few more; never grows without bound
                self._rows_cache.pop(next(iter(self._rows_cache)))
            self._rows_cache[id(indices)] = (indices, indices._version, str(dev), rows)
        return rows

    def forward(self, X, neis, num_w, walk_len, indices, layer_type, indxx=None):
        dev = X.device
        if dev.type != "cuda":
            raise RuntimeError("pathnet_b200 PAGG runs on CUDA only (no CPU fallback); X is on %s" % dev)
        rows = self._rows_of(indices, dev)
        S = rows.numel()
        neis = torch.as_tensor(neis).to(dev, dtype=torch.int64)            # reference: neis.to(device) (:176,:244)
        layer_type = torch.as_tensor(layer_type).to(dev, dtype=torch.int64)  # (:184,:248)
        if neis.shape[0] != S:
            raise ValueError(f"neis has {neis.shape[0]} rows but sum(indices) = {S}")
        if neis.numel() != S * num_w * walk_len or layer_type.numel() != S * num_w * walk_len:
            raise ValueError("neis / layer_type do not hold S*num_w*walk_len entries")
        if walk_len != len(self.nets):
            raise ValueError(f"walk_len {walk_len} != number of per-distance layers {len(self.nets)}")
        seq_node, seq_type, partner = canonicalize(self._variant, neis, layer_type.view(S, num_w, walk_len),
                                                   num_w, walk_len)
        seed = 0
        if self.training and self.dropout > 0:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # follows torch.manual_seed
        sd = {k: v for k, v in self.named_parameters()}
        return pagg_forward(self._variant, X, sd, seq_node, seq_type, partner, rows, dropout=self.dropout,
                            training=self.training, seed=seed, math_mode=self.math_mode)


class PathNet(_PathNetBase):
    """Drop-in for reference class `PathNet` (PathNet_run.py:150-211)."""
    _variant = "hetero"


class PathNet_homo(_PathNetBase):
    """Drop-in for reference class `PathNet_homo` (PathNet_run.py:214-278)."""
    _variant = "homo"

    def __init__(self, feature_length, hidden_size, out_size, wl, **kwargs):
        super().__init__(feature_length, hidden_size, out_size, wl, **kwargs)
        nn.init.xavier_uniform_(self.fc0.weight)   # :236-237
        nn.init.xavier_uniform_(self.fc2.weight)


PAGG = PathNet  # the name the README (:118) and baseline/GPRGNN/src/copy.py:299 use
