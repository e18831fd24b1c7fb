"""Stand-in with the STRUCTURE of the reference's PathNet_run.py (module-level
imports of packages that are not installed, module globals `device` and
`dropout`, classes named PathNet / PathNet_homo deriving from
torch_geometric's MessagePassing whose forward must never run) — written for
tests/test_gpu_launcher.py; none of the reference's code is in here."""
import sys

import matplotlib.pyplot as plt  # noqa: F401  (stubbed by pathnet_b200.run)
import numpy as np
import torch
import torch.nn as nn
from mpl_toolkits.mplot3d import Axes3D  # noqa: F401
from scipy.sparse.linalg.eigen.arpack import eigsh  # noqa: F401
from torch_geometric.datasets import Planetoid  # noqa: F401
from torch_geometric.nn import MessagePassing, GCNConv  # noqa: F401
from torch_geometric.utils import add_self_loops, degree  # noqa: F401

device = sys.argv[1] if len(sys.argv) > 1 else "cuda:0"
dropout = 0.0
out_path = sys.argv[2]


def _ctor(self, feature_length, hidden_size, out_size, wl, **kwargs):
    kwargs.setdefault('aggr', 'mean')
    MessagePassing.__init__(self)
    self.fc0 = nn.Linear(feature_length, hidden_size)
    self.LSTM = nn.LSTM(hidden_size, hidden_size)
    self.fc2 = nn.Linear(2 * hidden_size, out_size)
    self.nets = nn.ModuleList([nn.Linear(hidden_size, hidden_size) for _ in range(wl)])
    self.attw = nn.Linear(2 * hidden_size, 1)
    self.Lrelu = nn.LeakyReLU()


class PathNet(MessagePassing):
    __init__ = _ctor

    def forward(self, *a):
        raise AssertionError("the launcher should have replaced this forward")


class PathNet_homo(MessagePassing):
    __init__ = _ctor

    def forward(self, *a):
        raise AssertionError("the launcher should have replaced this forward")


torch.manual_seed(0)
N, F, H, C, nw, L, S = 80, 10, 32, 3, 4, 3, 25
X = torch.randn(N, F).to(device)
mask = np.zeros(N, dtype=bool)
mask[np.random.RandomState(0).permutation(N)[:S]] = True
neis = torch.randint(0, N, (S, nw * L))
ltype = torch.randint(0, L, (S, nw, L))
res = {}
for cls in (PathNet, PathNet_homo):
    model = cls(F, H, C, L).to(device)
    model.eval()
    y = model(X, neis, nw, L, mask, ltype, torch.arange(S * nw * L).to(device))
    loss = y.square().mean()
    loss.backward()
    res[cls.__name__] = dict(y=y.detach().cpu(), sd={k: v.detach().cpu() for k, v in model.state_dict().items()},
                             g={k: p.grad.detach().cpu() for k, p in model.named_parameters()})
torch.save(dict(res=res, X=X.cpu(), mask=mask, neis=neis, ltype=ltype, dims=(N, F, H, C, nw, L, S)), out_path)
print("fake reference script finished")
