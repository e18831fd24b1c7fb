"""The reference's training step as device work only.

`train_fixed_indices` (/root/reference/PathNet_run.py:345-352) does, per epoch:
forward on the train rows -> CrossEntropyLoss -> zero_grad / backward -> Adam
step, as ~60 separate host-launched PyTorch ops.  `TrainStep` runs the same
arithmetic through the C ABI — pn_pagg_canonicalize, pn_pagg_forward,
pn_ce_loss_grad, pn_pagg_backward, pn_adam_step — over preallocated buffers and
captures it as ONE CUDA graph: a step is an H2D copy of that epoch's index
tensors (int64 as the reference holds them, or the sampler's int32 / uint8)
plus one graph launch.  Parameters live in one flat buffer (the module's
parameters become views of it, so `state_dict()`, eval forwards and
`torch.save` keep working), gradients in another: the data-parallel exchange is
one in-place all-reduce of that buffer.

Dropout masks stay fresh under replay: the kernels add a device-resident
counter (pn_step_state.seed_add, advanced inside the graph) to the seed.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import pagg
from ._lib import PARAM_NAMES, AdamCfg, PaggCfg, PaggParams, check, lib

_STATE_KEYS = {"fc0_w": "fc0.weight", "fc0_b": "fc0.bias", "w_ih": "LSTM.weight_ih_l0", "w_hh": "LSTM.weight_hh_l0",
               "b_ih": "LSTM.bias_ih_l0", "b_hh": "LSTM.bias_hh_l0", "attw_w": "attw.weight", "attw_b": "attw.bias",
               "fc2_w": "fc2.weight", "fc2_b": "fc2.bias"}


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class FlatParams:
    """All twelve parameter tensors of a PathNet / PathNet_homo module in ONE fp32 buffer, in the order of the C ABI's
    pn_pagg_params (every slice 16-byte aligned); the module's parameters are re-pointed at views of it."""

    def __init__(self, model):
        sd = dict(model.named_parameters())
        L = len(model.nets)
        dev = sd["fc0.weight"].device
        shapes = {}
        for k in PARAM_NAMES:
            if k == "nets_w":
                shapes[k] = (L,) + tuple(sd["nets.0.weight"].shape)
            elif k == "nets_b":
                shapes[k] = (L,) + tuple(sd["nets.0.bias"].shape)
            else:
                shapes[k] = tuple(sd[_STATE_KEYS[k]].shape)
        self.shapes, self.offsets, off = shapes, {}, 0
        for k in PARAM_NAMES:
            self.offsets[k] = off
            off += (int(np.prod(shapes[k])) + 3) // 4 * 4
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for k in PARAM_NAMES:
                v = self.view(self.flat, k)
                if k == "nets_w":
                    for d in range(L):
                        v[d].copy_(sd[f"nets.{d}.weight"]); sd[f"nets.{d}.weight"].data = v[d]
                elif k == "nets_b":
                    for d in range(L):
                        v[d].copy_(sd[f"nets.{d}.bias"]); sd[f"nets.{d}.bias"].data = v[d]
                else:
                    p = sd[_STATE_KEYS[k]]
                    v.copy_(p); p.data = v

    def view(self, flat, k):
        n = int(np.prod(self.shapes[k]))
        return flat[self.offsets[k]:self.offsets[k] + n].view(self.shapes[k])

    def struct(self, flat) -> PaggParams:
        s = PaggParams()
        for k in PARAM_NAMES:
            setattr(s, k, flat.data_ptr() + 4 * self.offsets[k])
        return s


class TrainStep:
    """step(neis, layer_type) -> device scalar loss: one training step of PathNet_run.py:345-352 on `rows` (the
    train nodes), their labels `y_rows`, with Adam(lr, weight_decay) — the reference's optimiser (:295-296).

    neis / layer_type: that epoch's [S, nw*L] / [S, nw, L] index tensors, on the host (pinned or not) or the device,
    int64 (the reference's) or int32 / uint8 (the sampler's); dtype and shape are fixed by the first call.
    world > 1 (torch.distributed initialised): the flat gradient buffer is all-reduced in place between backward and
    the optimiser step.  Two ways to split the rows: weak scaling — every rank passes its own rows / index tensors
    and `s_global` = the global row count; strong scaling — every rank passes the SAME full-batch `indices` and
    index tensors plus `shard=(lo, hi)`: canonicalisation covers the whole batch (the hetero variant's output rows
    pool other nodes' paths, so only the canonical tensors can be cut), forward / backward run on rows [lo, hi)."""

    def __init__(self, model, X, indices, y_rows, num_w, walk_len, *, lr=0.005, weight_decay=0.0005, betas=(0.9, 0.999),
                 eps=1e-8, dropout=None, math_mode=None, row_chunk=None, seed=None, graph=True, s_global=None, group=None,
                 shard=None, graph_collective=None):
        if X.device.type != "cuda":
            raise RuntimeError("pathnet_b200 TrainStep runs on CUDA only (no CPU fallback); X is on %s" % X.device)
        self.model, self.dev = model, X.device
        self.X = X.detach().contiguous().float()
        self.variant = model._variant if hasattr(model, "_variant") else ("homo" if type(model).__name__ == "PathNet_homo" else "hetero")
        self.nw, self.L = int(num_w), int(walk_len)
        self.dropout = float(model.dropout if dropout is None else dropout)
        H, F = model.fc0.weight.shape
        mode = pagg.resolve_math_mode(getattr(model, "math_mode", "auto") if math_mode is None else math_mode, H, self.L)
        self.rows = pagg.RowsCache().get(indices, self.dev) if not (torch.is_tensor(indices) and indices.dtype == torch.int32 and indices.device == self.dev) else indices
        S = self.rows.numel()
        self.S_total, self.group = S, group
        self.lo, self.hi = (0, S) if shard is None else (int(shard[0]), int(shard[1]))
        if not (0 <= self.lo <= self.hi <= S):
            raise ValueError(f"shard {shard} outside the {S} rows")
        self.s_global = int(s_global if s_global is not None else S)
        self.y = torch.as_tensor(y_rows).to(self.dev)
        if self.y.dtype not in (torch.int64, torch.int32):
            self.y = self.y.long()
        if self.y.numel() != S:
            raise ValueError(f"y_rows has {self.y.numel()} labels for {S} rows")
        self.y = self.y[self.lo:self.hi].contiguous()
        S_all, S = S, self.hi - self.lo
        self.S = S
        Cc = model.fc2.weight.shape[0]
        self.params = FlatParams(model)
        n = self.params.numel
        # gradients: one flat buffer in the parameters' layout, + one 16-byte slot at the end where the CE kernel accumulates
        # the loss — so the data-parallel exchange (gradients AND the loss value) is ONE all-reduce
        self.grad_all = torch.zeros(n + 4, dtype=torch.float32, device=self.dev)
        self.grad = self.grad_all[:n]
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=self.dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=self.dev)
        # p.grad = views of the flat gradient buffer (so code that inspects .grad keeps working)
        for k, p in model.named_parameters():
            p.grad = None
        self.adam = AdamCfg(lr=lr, beta1=betas[0], beta2=betas[1], eps=eps, weight_decay=weight_decay)
        base_seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)
        self.state = torch.zeros(2, dtype=torch.int64, device=self.dev)   # pn_step_state {seed_add, step}
        self.cfg = PaggCfg(N=self.X.shape[0], F=F, H=H, C=Cc, S=S, nw=self.nw, L=self.L, variant=pagg.VARIANT[self.variant],
                           math_mode=pagg.MATH[mode], training=1, dropout_p=self.dropout, dropout_seed=base_seed,
                           no_backward=0, row_chunk=int(getattr(model, "row_chunk", 0) if row_chunk is None else row_chunk),
                           dropout_seed_dev=self.state.data_ptr(), x16=None, cell=0,
                           pool=pagg.POOL[getattr(model, "pool", "attention")])
        self.x16 = None
        if mode == "tc":   # X is constant over the run: its fp16 copy for the TMA-loaded fc0 GEMM is made once
            self.x16 = torch.empty(lib().pn_pagg_x16_bytes(self.cfg.N, F), dtype=torch.uint8, device=self.dev)
            with torch.cuda.device(self.dev):
                check(lib().pn_pagg_pack_x16(self.cfg.N, F, self.X.data_ptr(), self.x16.data_ptr(), torch.cuda.current_stream().cuda_stream))
            self.cfg.x16 = self.x16.data_ptr()
        with torch.cuda.device(self.dev):
            ws_bytes = lib().pn_pagg_workspace_bytes(C.byref(self.cfg))
            if ws_bytes == 0:
                check(-1)
        self.ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.dev)
        self.seq_node = torch.empty((S_all, self.nw, self.L), dtype=torch.int32, device=self.dev)
        self.seq_type = torch.empty((S_all, self.nw, self.L), dtype=torch.uint8, device=self.dev)
        self.partner = torch.empty((S_all, self.nw), dtype=torch.int32, device=self.dev)
        self.logits = torch.empty((S, Cc), dtype=torch.float32, device=self.dev)
        self.dlogits = torch.empty_like(self.logits)
        self.loss = self.grad_all[n:n + 1]
        self.guard = pagg.IndexGuard(self.dev)
        self.neis_buf = self.lt_buf = None
        self.use_graph, self.graph = bool(graph), None
        # multi-GPU: capture the gradient all-reduce and the Adam kernel in the same graph as forward / backward (one launch per
        # step and no host gap before the collective); PN_GRAPH_NCCL=0 keeps the collective outside the graph
        import os
        self.graph_collective = (os.environ.get("PN_GRAPH_NCCL", "1") != "0") if graph_collective is None else bool(graph_collective)
        self._copy_stream = None
        self.steps = 0
        self.launches_per_step = 0

    # ------------------------------------------------------------------ device work of one step
    def _enqueue_fwd_bwd(self):
        L_, st = lib(), torch.cuda.current_stream().cuda_stream
        ps, gs = self.params.struct(self.params.flat), self.params.struct(self.grad)
        check(L_.pn_step_advance(self.state.data_ptr(), st))
        per = self.lo * self.nw
        sn, sty = self.seq_node.data_ptr() + 4 * per * self.L, self.seq_type.data_ptr() + per * self.L
        pt, rw = self.partner.data_ptr() + 4 * per, self.rows.data_ptr() + 4 * self.lo
        check(L_.pn_pagg_canonicalize(self.cfg.variant, self.cfg.N, self.S_total, self.nw, self.L, self.neis_buf.data_ptr(),
                                      pagg.IDX_DTYPE[self.neis_buf.dtype], self.lt_buf.data_ptr(),
                                      pagg.IDX_DTYPE[self.lt_buf.dtype], self.seq_node.data_ptr(), self.seq_type.data_ptr(),
                                      self.partner.data_ptr(), self.guard.dev_word.data_ptr(), st))
        check(L_.pn_pagg_forward(C.byref(self.cfg), C.byref(ps), self.X.data_ptr(), sn, sty, pt, rw, self.ws.data_ptr(),
                                 self.logits.data_ptr(), st))
        self.loss.zero_()
        check(L_.pn_ce_loss_grad(self.S, self.cfg.C, self.logits.data_ptr(), self.y.data_ptr(), pagg.IDX_DTYPE[self.y.dtype],
                                 1.0 / self.s_global, self.loss.data_ptr(), self.dlogits.data_ptr(), st))
        check(L_.pn_pagg_backward(C.byref(self.cfg), C.byref(ps), self.X.data_ptr(), sn, sty, pt, rw, self.ws.data_ptr(),
                                  self.dlogits.data_ptr(), C.byref(gs), st))

    def _enqueue_adam(self):
        check(lib().pn_adam_step(self.params.numel, self.params.flat.data_ptr(), self.grad.data_ptr(),
                                 self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), C.byref(self.adam),
                                 self.state.data_ptr(), torch.cuda.current_stream().cuda_stream))

    def _enqueue(self, single):
        n0 = lib().pn_launch_count()
        self._enqueue_fwd_bwd()
        if single:
            self._enqueue_adam()
        elif self.graph_collective:   # the all-reduce and the optimiser step belong to the same graph (NCCL supports capture)
            self._reduce_and_update()
        self.launches_per_step = int(lib().pn_launch_count() - n0)   # kernels of this library per step (a replayed graph launches the same nodes)

    # ------------------------------------------------------------------ public
    def _stage(self, neis, layer_type):
        neis, layer_type = torch.as_tensor(neis), torch.as_tensor(layer_type)
        if neis.dtype not in (torch.int64, torch.int32):
            neis = neis.long()
        if layer_type.dtype not in pagg.IDX_DTYPE:
            layer_type = layer_type.long()
        if self.neis_buf is None:
            if neis.numel() != self.S_total * self.nw * self.L or layer_type.numel() != self.S_total * self.nw * self.L:
                raise ValueError("neis / layer_type do not hold S*num_w*walk_len entries")
            self.neis_buf = torch.empty(neis.numel(), dtype=neis.dtype, device=self.dev)
            self.lt_buf = torch.empty(layer_type.numel(), dtype=layer_type.dtype, device=self.dev)
        if neis.dtype != self.neis_buf.dtype or layer_type.dtype != self.lt_buf.dtype or neis.numel() != self.neis_buf.numel():
            raise ValueError("TrainStep: the index tensors changed dtype or size since the first step")
        self.neis_buf.copy_(neis.reshape(-1), non_blocking=True)
        self.lt_buf.copy_(layer_type.reshape(-1), non_blocking=True)

    def prefetch(self, neis, layer_type):
        """Start the H2D copy of the NEXT step's index tensors on a side stream (landing buffers alternate), so it
        overlaps the step that is running; the following `step()` (no arguments) consumes it."""
        neis, layer_type = torch.as_tensor(neis), torch.as_tensor(layer_type)
        with torch.cuda.device(self.dev):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream()
                self._landing = [None, None]
                self._land_free = [None, None]
            i = self._land_i = 1 - getattr(self, "_land_i", 1)
            if self._landing[i] is None:
                self._landing[i] = (torch.empty(neis.numel(), dtype=neis.dtype, device=self.dev),
                                    torch.empty(layer_type.numel(), dtype=layer_type.dtype, device=self.dev))
            with torch.cuda.stream(self._copy_stream):
                if self._land_free[i] is not None:
                    self._copy_stream.wait_event(self._land_free[i])   # the step that read this landing buffer is done with it
                self._landing[i][0].copy_(neis.reshape(-1), non_blocking=True)
                self._landing[i][1].copy_(layer_type.reshape(-1), non_blocking=True)
                self._land_ready = torch.cuda.Event()
                self._land_ready.record()

    def step(self, neis=None, layer_type=None):
        self.guard.poll()
        with torch.cuda.device(self.dev):
            if neis is None:
                if getattr(self, "_land_ready", None) is None:
                    raise ValueError("TrainStep.step() without arguments needs a prefetch() first")
                torch.cuda.current_stream().wait_event(self._land_ready)
                i = self._land_i
                self._stage(*self._landing[i])                         # device -> device into the graph's static buffers
                self._land_free[i] = torch.cuda.Event()
                self._land_free[i].record()
                self._land_ready = None
            else:
                self._stage(neis, layer_type)
            single = _world(self.group) == 1
            if not self.use_graph:
                self._enqueue_fwd_bwd()
                if single:
                    self._enqueue_adam()
            else:
                if self.graph is None:
                    # one eager step first (function attributes, lazy one-time queries), on a side stream as capture requires
                    s = torch.cuda.Stream()
                    s.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(s):
                        self._enqueue(single)
                    torch.cuda.current_stream().wait_stream(s)
                    self.steps += 1
                    if not single and not self.graph_collective:
                        self._reduce_and_update()
                    self.guard.arm()
                    self.graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph):
                        self._enqueue(single)
                    return self.loss
                self.graph.replay()
            if not single and not (self.use_graph and self.graph_collective):
                self._reduce_and_update()
            self.guard.arm()
        self.steps += 1
        return self.loss

    def _reduce_and_update(self):
        dist.all_reduce(self.grad_all, op=dist.ReduceOp.SUM, group=self.group)   # gradients + the loss slot, in place
        self._enqueue_adam()

    def grads(self):
        """{state_dict key: view of the flat gradient buffer} (after a step: the gradients that step applied)."""
        out, L = {}, len(self.model.nets)
        for k in PARAM_NAMES:
            v = self.params.view(self.grad, k)
            if k in ("nets_w", "nets_b"):
                for d in range(L):
                    out[f"nets.{d}.{'weight' if k == 'nets_w' else 'bias'}"] = v[d]
            else:
                out[_STATE_KEYS[k]] = v
        return out
