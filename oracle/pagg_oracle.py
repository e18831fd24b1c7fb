"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference PAGG operator.

A plain-PyTorch (CPU, fp32 or fp64) restatement of `PathNet.forward`
(/root/reference/PathNet_run.py:172-211) and `PathNet_homo.forward`
(:239-278), written op by op from the reference's semantics, with the LSTM
spelled out (PyTorch nn.LSTM semantics: gate order i,f,g,o, both biases, zero
initial state) so autograd supplies the backward the reference gets from
`loss.backward()` (:351).  Only tests/, __graft_entry__.smoke() and bench.py's
CPU-baseline legs may import it — never the product.

Parity pinning: tests/golden/make_pagg_golden.py runs the reference's OWN
classes (AST-extracted from the read-only mount, never copied) on seeded
inputs and commits inputs + outputs + gradients; tests/test_oracle_pinning.py
checks this restatement against those fixtures and, when /root/reference is
present, against the live classes.  The third-party arithmetic involved
(torch nn.LSTM / nn.Linear / F.softmax, torch pinned to 1.11 in the
reference's requirements.txt:7, 2.11 here) is pinned only through those runs.

Dropout: the reference draws masks from torch's global Philox stream
(F.dropout, :194,:209,:264,:276), which nothing else can reproduce, so the
restatement takes the two masks explicitly (already scaled by 1/(1-p)).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

PARAM_KEYS = ("fc0.weight", "fc0.bias", "LSTM.weight_ih_l0", "LSTM.weight_hh_l0",
              "LSTM.bias_ih_l0", "LSTM.bias_hh_l0", "fc2.weight", "fc2.bias", "attw.weight",
              "attw.bias")


def lstm_last_hidden(x, w_ih, w_hh, b_ih, b_hh):
    """x [L, B, H] -> h_L [B, H]; nn.LSTM(hidden, hidden) single layer (:164,:195)."""
    L, B, H = x.shape
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    for t in range(L):
        gates = x[t] @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
        i, f, g, o = gates.split(H, dim=1)
        i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
        c = f * c + i * g
        h = o * torch.tanh(c)
    return h


def _nets(params, L):
    return [(params[f"nets.{d}.weight"], params[f"nets.{d}.bias"]) for d in range(L)]


def pagg_forward(variant, params, X, neis, num_w, walk_len, indices, layer_type,
                 mask_seq=None, mask_cat=None):
    """variant: 'hetero' (class PathNet) or 'homo' (class PathNet_homo).
    params: dict with the reference's state_dict keys.  neis [S, nw*L] int64,
    layer_type [S, nw, L] int64, indices bool [N].  mask_seq [L, P, H] /
    mask_cat [S, 2H]: dropout masks incl. the 1/(1-p) scale, or None."""
    H = params["fc0.weight"].shape[0]
    indices = torch.as_tensor(np.asarray(indices)) if not torch.is_tensor(indices) else indices
    split = int(indices.sum())
    P = split * num_w
    Z = X @ params["fc0.weight"].t() + params["fc0.bias"]                      # :175 / :242
    if variant == "homo":
        Z = F.relu(Z)                                                           # :243
    lt = layer_type.reshape(P * walk_len)
    if variant == "hetero":
        nei = Z[neis].view(P, walk_len, H)                                      # :179
        neis_t = nei.transpose(0, 1)                                            # :180  [L, P, H]
        rows = torch.flip(neis_t, dims=[0]).reshape(P * walk_len, H)            # :182-183
    else:
        rows = Z[neis].view(P * walk_len, H)                                    # :246
    # every Linear on every row, then select by label (:185-191 / :249-255)
    proj = torch.stack([rows @ w.t() + b for (w, b) in _nets(params, walk_len)], dim=1)
    sel = proj[torch.arange(P * walk_len), lt].view(P, walk_len, H)
    if variant == "homo":
        sel = F.relu(sel)                                                       # :257
        ego_full = sel.reshape(split, num_w, walk_len, H)[:, :, 0, :]           # :259-260
    seq = sel.transpose(0, 1)                                                   # [L, P, H]
    if mask_seq is not None:
        seq = seq * mask_seq                                                    # :194 / :264
    h_n = lstm_last_hidden(seq, params["LSTM.weight_ih_l0"], params["LSTM.weight_hh_l0"],
                           params["LSTM.bias_ih_l0"], params["LSTM.bias_hh_l0"])  # :195 / :265
    aw, ab = params["attw.weight"], params["attw.bias"]
    if variant == "hetero":
        h_v = h_n.view(num_w, split, H)                                         # :196-197
        partner = neis_t[0].reshape(num_w, split, H)                            # :199
        score = F.leaky_relu(torch.cat((h_v, partner), dim=-1) @ aw.t() + ab, 0.01)
        att = F.softmax(score, dim=0)        # implicit dim of a 3-D tensor is 0 (:200-201)
        pooled = (att * h_v).mean(dim=0)                                        # :202-204
    else:
        h_v = h_n.view(split, num_w, H)                                         # :266-267
        score = torch.cat((h_v, ego_full), dim=-1) @ aw.t() + ab                # :269-270
        pooled = ((1 + score) * h_v).mean(dim=1)                                # :271-273
    layer1 = torch.cat((Z[indices], pooled), dim=1)                             # :206-207 / :274-275
    if mask_cat is not None:
        layer1 = layer1 * mask_cat                                              # :209 / :276
    return layer1 @ params["fc2.weight"].t() + params["fc2.bias"]               # :210 / :277


def canonicalize(variant, neis, layer_type, num_w, walk_len):
    """Integer restatement of the index bookkeeping above: which (type, node)
    table entry feeds step t of the w-th sequence pooled into output row s,
    and which node's embedding is that sequence's attention partner.
    Returns seq_node [S,nw,L] int32, seq_type [S,nw,L] uint8, partner [S,nw] int32."""
    neis = np.asarray(neis).astype(np.int64)
    S = neis.shape[0]
    L, nw = walk_len, num_w
    P = S * nw
    idx = neis.reshape(P, L)
    typ = np.asarray(layer_type).astype(np.int64).reshape(P * L)
    if variant == "homo":
        seq_node = idx.reshape(S, nw, L)
        seq_type = typ.reshape(S, nw, L)
        partner = idx[:, 0].reshape(S, nw)
    else:
        r = np.arange(P * L)
        node_r = idx[r % P, L - 1 - r // P]          # flip + reshape (:182-183)
        rows_node = node_r.reshape(P, L)             # LSTM batch b, step j <- row b*L + j (:191)
        rows_type = typ.reshape(P, L)
        b = (np.arange(nw)[None, :] * S + np.arange(S)[:, None])   # [S, nw]: b_k = k*S + s (:196-197)
        seq_node = rows_node[b]
        seq_type = rows_type[b]
        partner = idx[b, 0]                                         # :199
    return seq_node.astype(np.int32), seq_type.astype(np.uint8), partner.astype(np.int32)


def pagg_forward_canonical(variant, params, X, seq_node, seq_type, partner, rows,
                           mask_seq=None, mask_cat=None):
    """The same operator evaluated from the canonical index form through the
    hoisted (type, node) table — what the CUDA path computes.  Used to show the
    canonical form equals the literal restatement (tests) and as the
    device-layout reference for dropout masks: mask_seq here is [S, nw, L, H]
    in canonical order."""
    H = params["fc0.weight"].shape[0]
    S, nw, L = seq_node.shape
    homo = variant == "homo"
    Z = X @ params["fc0.weight"].t() + params["fc0.bias"]
    if homo:
        Z = F.relu(Z)
    T = torch.stack([Z @ w.t() + b for (w, b) in _nets(params, L)], dim=0)      # [L, N, H]
    if homo:
        T = F.relu(T)
    sn = torch.as_tensor(seq_node.astype(np.int64))
    stp = torch.as_tensor(seq_type.astype(np.int64))
    seq = T[stp, sn]                                                            # [S, nw, L, H]
    part = seq[:, :, 0, :] if homo else Z[torch.as_tensor(partner.astype(np.int64))]
    if mask_seq is not None:
        seq = seq * mask_seq
    x = seq.reshape(S * nw, L, H).transpose(0, 1)
    h = lstm_last_hidden(x, params["LSTM.weight_ih_l0"], params["LSTM.weight_hh_l0"],
                         params["LSTM.bias_ih_l0"], params["LSTM.bias_hh_l0"]).view(S, nw, H)
    score = torch.cat((h, part), dim=-1) @ params["attw.weight"].t() + params["attw.bias"]
    if homo:
        pooled = ((1 + score) * h).mean(dim=1)
    else:
        att = F.softmax(F.leaky_relu(score, 0.01), dim=1)
        pooled = (att * h).mean(dim=1)
    layer1 = torch.cat((Z[torch.as_tensor(np.asarray(rows).astype(np.int64))], pooled), dim=1)
    if mask_cat is not None:
        layer1 = layer1 * mask_cat
    return layer1 @ params["fc2.weight"].t() + params["fc2.bias"]


def load_reference_classes(dropout=0.0, device="cpu"):
    """The reference's own classes, AST-extracted from the read-only mount
    (authoring container only).  Nothing is copied into this repo."""
    import ast
    import torch.nn as nn
    src = open("/root/reference/PathNet_run.py").read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("PathNet", "PathNet_homo")]

    class MessagePassing(nn.Module):  # the reference never calls propagate()
        def __init__(self, *a, **k):
            super().__init__()

    ns = dict(torch=torch, nn=nn, F=F, MessagePassing=MessagePassing, device=device, dropout=dropout)
    exec(compile(ast.Module(body=keep, type_ignores=[]), "PathNet_run.py", "exec"), ns)
    return ns


# --------------------------------------------------------------------------- SURVEY 8f-4 variants
def rnn_last_hidden(x, w_ih, w_hh, b_ih, b_hh):
    """x [L, B, H] -> h_L [B, H]; nn.RNN(hidden, hidden), tanh, single layer (baseline/GPRGNN/src/copy.py:307,348)."""
    L, B, H = x.shape
    h = x.new_zeros(B, H)
    for t in range(L):
        h = torch.tanh(x[t] @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh)
    return h


def pagg_rnn_forward(params, X, neis, num_w, walk_len, indices, layer_type, mask_seq=None, mask_cat=None):
    """Restatement of the reference's OLDER aggregator, `class PAGG` (baseline/GPRGNN/src/copy.py:331-359): fc0, gather,
    all four nei{d} Linears then select by label, dropout, tanh RNN, mean over the paths, concat, dropout, fc2.
    params: that class's state_dict keys (fc0.*, RNN.*_l0, fc2.*, nei0..nei3.*).  mask_seq [L, P, H], mask_cat [S, 2H]."""
    H = params["fc0.weight"].shape[0]
    indices = torch.as_tensor(np.asarray(indices)) if not torch.is_tensor(indices) else indices
    split = int(indices.sum())
    Z = X @ params["fc0.weight"].t() + params["fc0.bias"]                                   # :333
    nei = Z[neis].view(split * num_w * walk_len, H)                                          # :337
    lt = layer_type.reshape(split * num_w * walk_len)                                        # :340
    proj = torch.stack([nei @ params[f"nei{d}.weight"].t() + params[f"nei{d}.bias"] for d in range(4)], dim=1)   # :342-346
    seq = proj[torch.arange(nei.shape[0]), lt].view(split * num_w, walk_len, H).transpose(0, 1)                 # :348-349
    if mask_seq is not None:
        seq = seq * mask_seq                                                                 # :351
    h_n = rnn_last_hidden(seq, params["RNN.weight_ih_l0"], params["RNN.weight_hh_l0"], params["RNN.bias_ih_l0"],
                          params["RNN.bias_hh_l0"]).view(split, num_w, H)                    # :352-354
    layer1 = torch.cat((Z[indices], h_n.mean(dim=1)), dim=1)                                # :356-358
    if mask_cat is not None:
        layer1 = layer1 * mask_cat                                                           # :360
    return layer1 @ params["fc2.weight"].t() + params["fc2.bias"]                            # :361


def gru_last_hidden(x, w_ih, w_hh, b_ih, b_hh):
    """x [L, B, H] -> h_L [B, H]; torch nn.GRU(hidden, hidden), single layer, h_0 = 0, written out step by step (the "GRU"
    aggregator ablation of performance.jpg; the reference ships no code for it, so the definition is torch's):
        r = sigmoid(W_ir x + b_ir + W_hr h + b_hr);  z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
        n = tanh(W_in x + b_in + r * (W_hn h + b_hn));  h' = (1 - z) * n + z * h
    Pinned against torch.nn.GRU itself in tests/test_oracle_pinning.py."""
    L, B, H = x.shape
    h = x.new_zeros(B, H)
    for t in range(L):
        gi = x[t] @ w_ih.t() + b_ih
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
    return h


def pagg_forward_cell_variant(variant, params, X, neis, num_w, walk_len, indices, layer_type, cell="gru"):
    """PathNet / PathNet_homo with the LSTM replaced by another sequence encoder under the same `LSTM.*` keys (cell = "gru"):
    everything else -- projections, the variant's attention, pooling, classifier -- as in pagg_forward_canonical()."""
    assert cell == "gru"
    sn, st, pt = canonicalize(variant, neis.numpy(), layer_type.numpy(), num_w, walk_len)
    idx = torch.as_tensor(np.asarray(indices)) if not torch.is_tensor(indices) else indices
    rows = torch.nonzero(idx).reshape(-1).numpy()
    H = params["fc0.weight"].shape[0]
    S, nw, L = sn.shape
    homo = variant == "homo"
    Z = X @ params["fc0.weight"].t() + params["fc0.bias"]
    if homo:
        Z = F.relu(Z)
    T = torch.stack([Z @ w.t() + b for (w, b) in _nets(params, L)], dim=0)
    if homo:
        T = F.relu(T)
    seq = T[torch.as_tensor(st.astype(np.int64)), torch.as_tensor(sn.astype(np.int64))]
    part = seq[:, :, 0, :] if homo else Z[torch.as_tensor(pt.astype(np.int64))]
    h = gru_last_hidden(seq.reshape(S * nw, L, H).transpose(0, 1), params["LSTM.weight_ih_l0"], params["LSTM.weight_hh_l0"],
                        params["LSTM.bias_ih_l0"], params["LSTM.bias_hh_l0"]).view(S, nw, H)
    score = torch.cat((h, part), dim=-1) @ params["attw.weight"].t() + params["attw.bias"]
    if homo:
        pooled = ((1 + score) * h).mean(dim=1)
    else:
        pooled = (F.softmax(F.leaky_relu(score, 0.01), dim=1) * h).mean(dim=1)
    return torch.cat((Z[torch.as_tensor(rows.astype(np.int64))], pooled), dim=1) @ params["fc2.weight"].t() + params["fc2.bias"]


def pagg_forward_pool_variant(variant, params, X, neis, num_w, walk_len, indices, layer_type, pool):
    """PathNet / PathNet_homo with the attention replaced by mean / sum / max over the paths (the pooling ablations
    of performance.jpg): everything up to h_n as in pagg_forward(), then the plain reduction."""
    H = params["fc0.weight"].shape[0]
    S = int(torch.as_tensor(np.asarray(indices)).sum()) if not torch.is_tensor(indices) else int(indices.sum())
    sn, st, pt = canonicalize(variant, neis.numpy(), layer_type.numpy(), num_w, walk_len)
    homo = variant == "homo"
    Z = X @ params["fc0.weight"].t() + params["fc0.bias"]
    if homo:
        Z = F.relu(Z)
    T = torch.stack([Z @ w.t() + b for (w, b) in _nets(params, walk_len)], dim=0)
    if homo:
        T = F.relu(T)
    seq = T[torch.as_tensor(st.astype(np.int64)), torch.as_tensor(sn.astype(np.int64))]       # [S, nw, L, H]
    h = lstm_last_hidden(seq.reshape(S * num_w, walk_len, H).transpose(0, 1), params["LSTM.weight_ih_l0"],
                         params["LSTM.weight_hh_l0"], params["LSTM.bias_ih_l0"], params["LSTM.bias_hh_l0"]).view(S, num_w, H)
    pooled = h.mean(dim=1) if pool == "mean" else h.sum(dim=1) if pool == "sum" else h.max(dim=1).values
    idx = torch.as_tensor(np.asarray(indices)) if not torch.is_tensor(indices) else indices
    return torch.cat((Z[idx], pooled), dim=1) @ params["fc2.weight"].t() + params["fc2.bias"]


def load_reference_pagg_rnn_class(device="cpu"):
    """The reference's own `class PAGG`, AST-extracted from baseline/GPRGNN/src/copy.py (authoring container only)."""
    import ast
    import torch.nn as nn
    tree = ast.parse(open("/root/reference/baseline/GPRGNN/src/copy.py").read())
    keep = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "PAGG"]

    class MessagePassing(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    ns = dict(torch=torch, nn=nn, F=F, MessagePassing=MessagePassing, device=device)
    exec(compile(ast.Module(body=keep, type_ignores=[]), "copy.py", "exec"), ns)
    return ns["PAGG"]
