"""Run the UNMODIFIED reference script with the fused operator behind it:

    python -m pathnet_b200.run /path/to/PathNet_run.py -data=cornell -hid=128 -c cuda:0 ...

The script imports a few packages it never uses on this path and that are not
installed here (torch_geometric, matplotlib, an old scipy module path); stub
modules stand in for them.  The stub `MessagePassing` base class rebinds
`forward` of the script's own `PathNet` / `PathNet_homo` classes
(PathNet_run.py:150,214) to pathnet_b200's operator: the reference constructors
still create the parameters (same init, same state_dict), and `dropout` /
`device` are read from the script's module globals where the reference reads
them (:71,:83).
"""
from __future__ import annotations

import os
import runpy
import sys
import types

import numpy as np
import torch
import torch.nn as nn


def _fused_forward(variant):
    def forward(self, X, neis, num_w, walk_len, indices, layer_type, indxx=None):
        from . import pagg
        mod = sys.modules.get(type(self).__module__)
        drop = float(getattr(mod, "dropout", 0.0)) if mod is not None else 0.0   # the script's module global (:71)
        st = self.__dict__.get("_pn_state")
        if st is None:   # per-module state of the fused path: mask -> rows cache (no per-call nonzero sync), index guard
            st = self.__dict__["_pn_state"] = dict(rows=pagg.RowsCache(), guard=pagg.IndexGuard(X.device), sink=pagg.GradSink(), x16=pagg.X16Cache())
        return pagg.fused_forward(self, variant, X, neis, num_w, walk_len, indices, layer_type, dropout=drop,
                                  math_mode=os.environ.get("PN_MATH_MODE", "auto"), rows_cache=st["rows"],
                                  guard=st["guard"], grad_sink=st["sink"], x16_cache=st["x16"])
    return forward


class MessagePassing(nn.Module):
    """Stand-in for torch_geometric.nn.MessagePassing (the reference subclasses
    it but never calls propagate())."""

    def __init__(self, *a, **k):
        super().__init__()

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        if cls.__name__ == "PathNet":
            cls.forward = _fused_forward("hetero")
        elif cls.__name__ == "PathNet_homo":
            cls.forward = _fused_forward("homo")


def install_stubs():
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    class _Anything:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k): return self
        def __getattr__(self, k):
            if k.startswith("__"):          # keep inspect / copy / pickle protocols honest
                raise AttributeError(k)
            return _Anything()

    def _module_getattr(k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything()

    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mod("matplotlib", pyplot=mod("matplotlib.pyplot", __getattr__=_module_getattr))
        mod("mpl_toolkits"); mod("mpl_toolkits.mplot3d", Axes3D=_Anything)
    try:
        import torch_geometric  # noqa: F401
        import torch_geometric.nn as tgnn
        tgnn.MessagePassing = MessagePassing
    except ImportError:
        mod("torch_geometric")
        mod("torch_geometric.datasets", Planetoid=_Anything)
        mod("torch_geometric.nn", MessagePassing=MessagePassing, GCNConv=_Anything)
        mod("torch_geometric.utils", add_self_loops=_Anything(), degree=_Anything(),
            from_scipy_sparse_matrix=_Anything())
    # dataset.py:23 `from scipy.sparse.linalg.eigen.arpack import eigsh`, :177 np.bool
    try:
        import scipy.sparse.linalg.eigen.arpack  # noqa: F401
    except Exception:
        from scipy.sparse.linalg import eigsh
        mod("scipy.sparse.linalg.eigen"); mod("scipy.sparse.linalg.eigen.arpack", eigsh=eigsh)
    for alias, typ in (("bool", bool), ("int", int), ("float", float)):
        if alias not in np.__dict__:
            setattr(np, alias, typ)


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__, file=sys.stderr)
        return 2
    script = os.path.abspath(argv[0])
    install_stubs()
    sys.argv = [script] + argv[1:]
    sys.path.insert(0, os.path.dirname(script))
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
