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
`self.dropout` (reference global `dropout`, :71; pass `dropout=` to the
constructor — the reference's command-line default is 0.7, :48) and the
device (taken from X).
"""
from __future__ import annotations

import ctypes as C
import os
import warnings

import numpy as np
import torch
import torch.nn as nn

from ._lib import PARAM_NAMES, PaggCfg, PaggParams, check, lib

VARIANT = {"hetero": 0, "homo": 1, "plain": 2}
MATH = {"fp32": 0, "tc": 1, "tc3": 2}
CELL = {"lstm": 0, "rnn": 1, "gru": 2}
POOL = {"attention": 0, "mean": 1, "sum": 2, "max": 3}
IDX_DTYPE = {torch.int64: 0, torch.int32: 1, torch.uint8: 2}
_warned_auto_tc = False


def resolve_math_mode(mode: str, H: int, L: int) -> str:
    """"fp32": strict-parity CUDA-core mode (1e-4 against the reference's fp32 math), any hidden size up to 128.
    "tc3": the LSTM on tcgen05 tensor cores at fp32-grade accuracy — every operand travels as two fp16 terms (hi + lo) and
    every product is three tensor-core passes into one fp32 accumulator, gates on ex2 / rcp, projections on the fp32
    kernels: held to the SAME parity bar as "fp32" at about a ninth of its step time (hidden 128, path_len <= 8).
    "tc": the fast mode — single-pass fp16 operands (11-bit significand, the width of TF32, which is what the reference itself
    runs on under cuDNN / cuBLAS on a current GPU), fp32 accumulation, tanh.approx: logits within ~1e-3 of fp32 math;
    half the step time of "tc3".  Both tensor-core modes need hidden 128 and path_len <= 8 and raise otherwise.
    "auto" (the modules' default, overridable with env PN_MATH_MODE): parity first — "tc3" where it exists, else "fp32";
    the first time it resolves to "tc3" a one-line notice names the faster opt-in mode."""
    global _warned_auto_tc
    if mode == "auto":
        mode = os.environ.get("PN_MATH_MODE", "auto")
    if mode == "auto":
        if H == 128 and L <= 8:
            if not _warned_auto_tc:
                _warned_auto_tc = True
                warnings.warn("pathnet_b200: math_mode='auto' selected 'tc3' (split-precision tensor-core kernels, the fp32 "
                              "mode's 1e-4 parity bar); pass math_mode='tc' or set PN_MATH_MODE=tc for the single-pass fp16 "
                              "kernels (2x faster, logits within ~1e-3 of fp32 math)", stacklevel=3)
            return "tc3"
        return "fp32"
    if mode not in MATH:
        raise ValueError(f"math_mode must be 'auto', 'fp32', 'tc' or 'tc3' (got {mode!r})")
    return mode


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def _params_struct(tensors) -> PaggParams:
    s = PaggParams()
    for k, t in zip(PARAM_NAMES, tensors):
        setattr(s, k, t.data_ptr())
    return s


class IndexGuard:
    """Asynchronous report of the range check pn_pagg_canonicalize does on the caller's index tensors.  On CUDA the
    reference's `X[neis]` / `nei[indxx, layer_type]` (:179,:191) turn an out-of-range index into a device-side assert
    that surfaces at a later synchronisation; here the kernel replaces the bad entry by 0 (so nothing ever indexes out
    of bounds), sets a status word, and the word is copied to pinned host memory in stream order.  `poll()` raises
    IndexError once that copy has landed — at the next forward()/backward() of the module, or immediately when
    `sync=True` (env PN_SYNC_CHECKS=1)."""

    def __init__(self, dev):
        self.dev_word = torch.zeros(1, dtype=torch.int32, device=dev)
        self.host_word = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.event = None

    def arm(self):
        self.host_word.copy_(self.dev_word, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()

    def poll(self, sync=False):
        if self.event is None:
            return
        if sync:
            self.event.synchronize()
        if not self.event.query():
            return
        self.event = None
        bits = int(self.host_word[0])
        if bits:
            self.dev_word.zero_()
            self.host_word.zero_()
            what = " and ".join(w for b, w in ((1, "a node id outside [0, N)"), (2, "a distance label outside [0, walk_len)"))
                                if bits & b)
            raise IndexError(f"pathnet_b200 PAGG: neis / layer_type held {what} (the reference raises on X[neis] / "
                             f"nei[indxx, layer_type]); the entry was replaced by 0")


_SYNC_CHECKS = os.environ.get("PN_SYNC_CHECKS", "0") not in ("", "0")


def canonicalize(variant: str, neis: torch.Tensor, layer_type: torch.Tensor, num_w: int, walk_len: int, N=None,
                 guard: IndexGuard | None = None):
    """Device index tensors of the reference (int64 as PathNet_run.py passes them, or the sampler's int32 node ids /
    uint8 labels) -> canonical (seq_node int32 [S,nw,L], seq_type uint8 [S,nw,L], partner int32 [S,nw])
    (pn_pagg_canonicalize).  N = number of nodes for the range check (None: no upper bound known -> 2^31-1)."""
    S = neis.shape[0]
    dev = neis.device
    if neis.dtype not in (torch.int64, torch.int32):
        neis = neis.long()
    if layer_type.dtype not in IDX_DTYPE:
        layer_type = layer_type.long()
    neis = neis.contiguous().view(S, num_w * walk_len)
    layer_type = layer_type.contiguous()
    seq_node = torch.empty((S, num_w, walk_len), dtype=torch.int32, device=dev)
    seq_type = torch.empty((S, num_w, walk_len), dtype=torch.uint8, device=dev)
    partner = torch.empty((S, num_w), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().pn_pagg_canonicalize(VARIANT[variant], int(N) if N is not None else 2 ** 31 - 1, S, num_w, walk_len,
                                         neis.data_ptr(), IDX_DTYPE[neis.dtype], layer_type.data_ptr(),
                                         IDX_DTYPE[layer_type.dtype], seq_node.data_ptr(), seq_type.data_ptr(),
                                         partner.data_ptr(), guard.dev_word.data_ptr() if guard is not None else None,
                                         _stream_ptr()))
        if guard is not None:
            guard.arm()
            if _SYNC_CHECKS:
                guard.poll(sync=True)
    return seq_node, seq_type, partner


class X16Cache:
    """fp16 copy of X for the tensor-core mode's TMA loads (pn_pagg_pack_x16), kept while the same tensor object is
    passed with an unchanged in-place version — X is constant over a training run (PathNet_run.py:300)."""

    def __init__(self):
        self.ent = None

    def get(self, X):
        e = self.ent
        if e is not None and e[0] is X and e[1] == X._version and e[2] == X.data_ptr():
            return e[3]
        N, F = X.shape
        buf = torch.empty(lib().pn_pagg_x16_bytes(N, F), dtype=torch.uint8, device=X.device)
        with torch.cuda.device(X.device):
            check(lib().pn_pagg_pack_x16(N, F, X.data_ptr(), buf.data_ptr(), _stream_ptr()))
        self.ent = (X, X._version, X.data_ptr(), buf)
        return buf


class GradSink:
    """Where the latest backward put the parameter gradients: ONE flat fp32 buffer (the order of PARAM_NAMES) whose
    slices are returned to autograd, so `p.grad` of every parameter aliases it.  pathnet_b200.dist all-reduces
    `flat` in place — no pack / unpack copies."""
    flat = None      # torch tensor [sum of 4-element-padded sizes]
    offsets = None   # element offset of each PARAM_NAMES entry
    payload = 0      # number of gradient elements in it (without the alignment padding)


class _PaggFunction(torch.autograd.Function):
    """logits = PAGG(X; 12 parameter tensors) over canonical index tensors."""

    @staticmethod
    def forward(ctx, X, fc0_w, fc0_b, nets_w, nets_b, w_ih, w_hh, b_ih, b_hh, attw_w, attw_b, fc2_w,
                fc2_b, seq_node, seq_type, partner, rows, meta):
        raw = (fc0_w, fc0_b, nets_w, nets_b, w_ih, w_hh, b_ih, b_hh, attw_w, attw_b, fc2_w, fc2_b)
        params = [t.detach().contiguous().float() for t in raw]
        Xc = X.detach().contiguous().float()
        dev = Xc.device
        if dev.type != "cuda":
            raise RuntimeError("pathnet_b200 PAGG runs on CUDA only (no CPU fallback); X is on %s" % dev)
        S, nw, L = seq_node.shape
        H, Fd = params[0].shape
        cfg = PaggCfg(N=Xc.shape[0], F=Fd, H=H, C=params[10].shape[0], S=S, nw=nw, L=L,
                      variant=VARIANT[meta["variant"]], math_mode=MATH[resolve_math_mode(meta["math_mode"], H, L)],
                      training=int(meta["training"]), dropout_p=float(meta["dropout"]),
                      dropout_seed=int(meta["seed"]), no_backward=int(not meta.get("need_backward", True)),
                      row_chunk=int(meta.get("row_chunk", 0) or 0), dropout_seed_dev=None, x16=None,
                      cell=CELL[meta.get("cell", "lstm")], pool=POOL[meta.get("pool", "attention")])
        x16 = None
        if cfg.math_mode == MATH["tc"] and meta.get("x16_cache") is not None and X.dtype == torch.float32 and X.is_contiguous():
            x16 = meta["x16_cache"].get(X)     # keyed on the caller's tensor object: valid while X is not modified in place
            cfg.x16 = x16.data_ptr()
        with torch.cuda.device(dev):
            ws_bytes = lib().pn_pagg_workspace_bytes(C.byref(cfg))
            if ws_bytes == 0:
                check(-1)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            logits = torch.empty((S, cfg.C), dtype=torch.float32, device=dev)
            ps = _params_struct(params)
            check(lib().pn_pagg_forward(C.byref(cfg), C.byref(ps), Xc.data_ptr(), seq_node.data_ptr(),
                                        seq_type.data_ptr(), partner.data_ptr(), rows.data_ptr(),
                                        ws.data_ptr(), logits.data_ptr(), _stream_ptr()))
        # the tensors backward re-reads go through save_for_backward, so an in-place update of a parameter (or of X)
        # between forward and backward trips autograd's version check instead of silently changing the gradients
        ctx.save_for_backward(X, *raw)
        ctx.cfg, ctx.ws, ctx.x16 = cfg, ws, x16
        ctx.idx = (seq_node, seq_type, partner, rows)
        ctx.sink = meta.get("grad_sink")
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        if ctx.ws is None:
            raise RuntimeError("pathnet_b200 PAGG: backward was called a second time on the same forward; the saved "
                               "activations are released after the first backward (run forward again)")
        saved = ctx.saved_tensors   # raises if any was modified in place since forward
        X, raw = saved[0], saved[1:]
        params = [t.detach().contiguous().float() for t in raw]
        Xc = X.detach().contiguous().float()
        cfg, ws = ctx.cfg, ctx.ws
        seq_node, seq_type, partner, rows = ctx.idx
        dlogits = dlogits.contiguous().float()
        sizes = [p.numel() for p in params]
        padded = [(n + 3) // 4 * 4 for n in sizes]   # every slice starts 16-byte aligned
        flat = torch.zeros(sum(padded), dtype=torch.float32, device=Xc.device)
        grads, offs, off = [], [], 0
        for p, n, np_ in zip(params, sizes, padded):
            grads.append(flat[off:off + n].view_as(p))
            offs.append(off)
            off += np_
        with torch.cuda.device(Xc.device):
            ps, gs = _params_struct(params), _params_struct(grads)
            check(lib().pn_pagg_backward(C.byref(cfg), C.byref(ps), Xc.data_ptr(), seq_node.data_ptr(),
                                         seq_type.data_ptr(), partner.data_ptr(), rows.data_ptr(),
                                         ws.data_ptr(), dlogits.data_ptr(), C.byref(gs), _stream_ptr()))
        ctx.ws = None   # the (large) workspace is released as soon as its one backward has run
        if ctx.sink is not None:
            ctx.sink.flat, ctx.sink.offsets, ctx.sink.payload = flat, offs, sum(sizes)
        out = [g if g.dtype == r.dtype else g.to(r.dtype) for g, r in zip(grads, raw)]   # autograd wants the input dtype
        return (None, *out, None, None, None, None, None)


def pagg_forward(variant, X, params: dict, seq_node, seq_type, partner, rows, *, dropout=0.0,
                 training=False, seed=0, math_mode="auto", row_chunk=0, grad_sink=None, x16_cache=None, cell="lstm",
                 pool="attention"):
    """Functional form over a dict with the reference's state_dict keys.  cell / pool: the SURVEY 8f-4 variants
    ("rnn": LSTM.* keys then hold nn.RNN-shaped tensors [H,H] / [H]; "gru": nn.GRU-shaped [3H,H] / [3H];
    "mean" | "sum" | "max": attw.* unused)."""
    L = seq_node.shape[2]
    if cell != "lstm" and math_mode == "auto":
        math_mode = "fp32"   # the tensor-core kernels implement the LSTM cell only
    nets_w = torch.stack([params[f"nets.{d}.weight"] for d in range(L)])
    nets_b = torch.stack([params[f"nets.{d}.bias"] for d in range(L)])
    need_bwd = torch.is_grad_enabled() and any(t.requires_grad for t in params.values() if torch.is_tensor(t))
    meta = dict(variant=variant, math_mode=math_mode, training=training, dropout=dropout, seed=seed,
                need_backward=need_bwd, row_chunk=row_chunk, grad_sink=grad_sink, x16_cache=x16_cache, cell=cell, pool=pool)
    return _PaggFunction.apply(X, params["fc0.weight"], params["fc0.bias"], nets_w, nets_b,
                               params["LSTM.weight_ih_l0"], params["LSTM.weight_hh_l0"],
                               params["LSTM.bias_ih_l0"], params["LSTM.bias_hh_l0"],
                               params["attw.weight"].reshape(-1), params["attw.bias"],
                               params["fc2.weight"], params["fc2.bias"], seq_node, seq_type, partner,
                               rows, meta)


class RowsCache:
    """Row numbers selected by a node mask (int32, on the device).  torch.nonzero synchronises with the device, and a
    training loop passes the same mask every step (PathNet_run.py:345-347), so the result is kept per mask: torch
    tensors by object identity (held alive, so the id cannot be recycled) + in-place version counter; numpy masks by
    their bytes (the reference's masks are numpy bool arrays of N entries — hashing them costs microseconds)."""

    def __init__(self, limit=8):
        self.limit, self.ents = limit, {}

    def get(self, indices, dev):
        if torch.is_tensor(indices):
            key = ("t", id(indices), str(dev))
            ent = self.ents.get(key)
            if ent is not None and ent[0] is indices and ent[1] == indices._version:
                return ent[2]
            mask, keep, ver = indices.to(dev), indices, indices._version
        else:
            arr = np.ascontiguousarray(np.asarray(indices))
            key = ("n", arr.shape, arr.dtype.str, hash(arr.tobytes()), str(dev))
            ent = self.ents.get(key)
            if ent is not None and np.array_equal(ent[0], arr):
                return ent[2]
            mask, keep, ver = torch.as_tensor(arr, device=dev), arr.copy(), 0
        rows = (torch.nonzero(mask.reshape(-1), as_tuple=False).reshape(-1) if mask.dtype == torch.bool
                else torch.nonzero(mask.reshape(-1) != 0, as_tuple=False).reshape(-1)).to(torch.int32)
        if len(self.ents) >= self.limit:   # train / val / test masks and a few more; never grows without bound
            self.ents.pop(next(iter(self.ents)))
        self.ents[key] = (keep, ver, rows)
        return rows


def fused_forward(self, variant, X, neis, num_w, walk_len, indices, layer_type, *, dropout, math_mode, row_chunk=0,
                  rows_cache=None, guard=None, grad_sink=None, x16_cache=None, cell="lstm", pool="attention", params=None):
    """The 7-argument forward of both reference classes on the fused operator; shared by the drop-in modules below
    and by the launcher that rebinds the forward of the reference's own classes (pathnet_b200/run.py)."""
    dev = X.device
    if dev.type != "cuda":
        raise RuntimeError("pathnet_b200 PAGG runs on CUDA only (no CPU fallback); X is on %s" % dev)
    if guard is not None:
        guard.poll()   # an index error of an earlier call surfaces here at the latest
    rows = (rows_cache or RowsCache()).get(indices, dev)
    S = rows.numel()
    # the reference moves both index tensors to the device inside forward (:176,:184,:244,:248); int32 / uint8
    # tensors (what gen_merw() hands over) are taken as they are — no int64 blow-up on either side of the copy
    neis = torch.as_tensor(neis)
    layer_type = torch.as_tensor(layer_type)
    neis = neis.to(dev, non_blocking=True) if neis.dtype in (torch.int64, torch.int32) else neis.to(dev, dtype=torch.int64)
    layer_type = (layer_type.to(dev, non_blocking=True) if layer_type.dtype in IDX_DTYPE
                  else layer_type.to(dev, dtype=torch.int64))
    if neis.shape[0] != S:
        raise ValueError(f"neis has {neis.shape[0]} rows but sum(indices) = {S}")
    if neis.numel() != S * num_w * walk_len or layer_type.numel() != S * num_w * walk_len:
        raise ValueError("neis / layer_type do not hold S*num_w*walk_len entries")
    if walk_len != len(self.nets):
        raise ValueError(f"walk_len {walk_len} != number of per-distance layers {len(self.nets)}")
    seq_node, seq_type, partner = canonicalize(variant, neis, layer_type.view(S, num_w, walk_len), num_w, walk_len,
                                               N=X.shape[0], guard=guard)
    seed = 0
    if self.training and dropout > 0:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # follows torch.manual_seed (host generator: no device sync)
    sd = params if params is not None else {k: v for k, v in self.named_parameters()}
    return pagg_forward(variant, X, sd, seq_node, seq_type, partner, rows, dropout=dropout, training=self.training,
                        seed=seed, math_mode=math_mode, row_chunk=row_chunk, grad_sink=grad_sink, x16_cache=x16_cache,
                        cell=cell, pool=pool)


class _PathNetBase(nn.Module):
    _variant = "hetero"

    def __init__(self, feature_length, hidden_size, out_size, wl, **kwargs):
        # kwargs mirrors the reference signature (`aggr` is set and ignored there, :156-157)
        self.dropout = float(kwargs.pop("dropout", 0.0))     # the reference reads a module global (:71); its CLI default is 0.7
        self.math_mode = kwargs.pop("math_mode", "auto")     # "auto" (= "tc3" where it exists, else "fp32") | "fp32" | "tc3" | "tc" (fast)
        self.row_chunk = int(kwargs.pop("row_chunk", 0))     # > 0: bounded workspace, see pn_pagg_cfg.row_chunk
        self.pool = kwargs.pop("pool", "attention")          # ablations: "mean" | "sum" | "max" instead of the attention
        if self.pool not in POOL:
            raise ValueError(f"pool must be one of {sorted(POOL)} (got {self.pool!r})")
        self.cell = kwargs.pop("cell", "lstm")               # sequence-encoder ablation: "gru" = nn.GRU under the same `LSTM.*` keys
        if self.cell not in ("lstm", "gru"):
            raise ValueError(f"cell must be 'lstm' or 'gru' (got {self.cell!r}); the tanh-RNN variant is class PAGG_RNN")
        kwargs.setdefault("aggr", "mean")
        super().__init__()
        self._rows_cache = RowsCache()
        self._guard = None
        self.grad_sink = GradSink()
        self._x16 = X16Cache()
        self.feature_length, self.hidden_size, self.out_size = feature_length, hidden_size, out_size
        # construction order = the reference's (:162-170), so a given torch seed
        # yields the same initial weights
        self.fc0 = nn.Linear(feature_length, hidden_size)
        self.LSTM = nn.LSTM(hidden_size, hidden_size) if self.cell == "lstm" else nn.GRU(hidden_size, hidden_size)
        self.fc2 = nn.Linear(2 * hidden_size, out_size)
        self.nets = nn.ModuleList([nn.Linear(hidden_size, hidden_size) for _ in range(wl)])
        self.attw = nn.Linear(2 * hidden_size, 1)
        self.Lrelu = nn.LeakyReLU()

    def check_indices(self):
        """Synchronise and raise IndexError if any earlier forward() saw an out-of-range node id / label."""
        if self._guard is not None:
            self._guard.poll(sync=True)

    def forward(self, X, neis, num_w, walk_len, indices, layer_type, indxx=None):
        if X.device.type != "cuda":
            raise RuntimeError("pathnet_b200 PAGG runs on CUDA only (no CPU fallback); X is on %s" % X.device)
        if self._guard is None or self._guard.dev_word.device != X.device:
            self._guard = IndexGuard(X.device)
        return fused_forward(self, self._variant, X, neis, num_w, walk_len, indices, layer_type, dropout=self.dropout,
                             math_mode=self.math_mode, row_chunk=self.row_chunk, rows_cache=self._rows_cache,
                             guard=self._guard, grad_sink=self.grad_sink, x16_cache=self._x16, pool=self.pool, cell=self.cell)


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


PAGG = PathNet  # the name the README (:118) uses for the aggregator


class PAGG_RNN(nn.Module):
    """Drop-in for the OLDER aggregator `class PAGG` of the reference (baseline/GPRGNN/src/copy.py:299-359; the
    "RNN / mean" ablation named in performance.jpg): fc0 -> gather -> one of four Linear(H,H) by distance label ->
    dropout -> tanh nn.RNN over the path -> plain mean over the paths -> concat with the ego embedding -> dropout ->
    fc2.  Same constructor `(feature_length, hidden_size, out_size, node_num)`, same 7-argument forward, same
    parameter names (fc0, RNN, fc2, nei0..nei3).  The reference hard-codes dropout p = 0.9 (:347,:354)."""

    def __init__(self, feature_length, hidden_size, out_size, node_num, **kwargs):
        self.dropout = float(kwargs.pop("dropout", 0.9))
        self.math_mode = kwargs.pop("math_mode", "fp32")
        super().__init__()
        self.feature_length, self.hidden_size, self.out_size, self.node_num = feature_length, hidden_size, out_size, node_num
        self.fc0 = nn.Linear(feature_length, hidden_size)
        self.RNN = nn.RNN(hidden_size, hidden_size)
        self.fc2 = nn.Linear(2 * hidden_size, out_size)
        self.nei0 = nn.Linear(hidden_size, hidden_size)
        self.nei1 = nn.Linear(hidden_size, hidden_size)
        self.nei2 = nn.Linear(hidden_size, hidden_size)
        self.nei3 = nn.Linear(hidden_size, hidden_size)
        for lin in (self.fc0, self.fc2, self.nei0, self.nei1, self.nei2, self.nei3):   # :312-317
            nn.init.xavier_uniform_(lin.weight)
        self.nets = [self.nei0, self.nei1, self.nei2, self.nei3]    # plain list: no extra state_dict keys
        self._rows_cache, self._guard, self.grad_sink, self._x16 = RowsCache(), None, GradSink(), X16Cache()

    def forward(self, X, neis, num_w, walk_len, indices, layer_type, indxx=None):
        if X.device.type != "cuda":
            raise RuntimeError("pathnet_b200 PAGG runs on CUDA only (no CPU fallback); X is on %s" % X.device)
        if walk_len != 4:
            raise ValueError("the reference's PAGG class has exactly four per-distance layers (walk_len must be 4)")
        if self._guard is None or self._guard.dev_word.device != X.device:
            self._guard = IndexGuard(X.device)
        H = self.hidden_size
        zero = X.new_zeros(1)
        params = {"fc0.weight": self.fc0.weight, "fc0.bias": self.fc0.bias, "fc2.weight": self.fc2.weight, "fc2.bias": self.fc2.bias,
                  "LSTM.weight_ih_l0": self.RNN.weight_ih_l0, "LSTM.weight_hh_l0": self.RNN.weight_hh_l0,
                  "LSTM.bias_ih_l0": self.RNN.bias_ih_l0, "LSTM.bias_hh_l0": self.RNN.bias_hh_l0,
                  "attw.weight": X.new_zeros(1, 2 * H), "attw.bias": zero}     # no attention in this class
        for d, lin in enumerate(self.nets):
            params[f"nets.{d}.weight"], params[f"nets.{d}.bias"] = lin.weight, lin.bias
        return fused_forward(self, "plain", X, neis, num_w, walk_len, indices, layer_type, dropout=self.dropout,
                             math_mode=self.math_mode, rows_cache=self._rows_cache, guard=self._guard, grad_sink=self.grad_sink,
                             x16_cache=self._x16, cell="rnn", pool="mean", params=params)
