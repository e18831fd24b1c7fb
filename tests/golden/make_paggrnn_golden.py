"""Generate the committed fixture of the reference's older aggregator (`class PAGG`,
baseline/GPRGNN/src/copy.py:299-359: tanh RNN + plain mean) from the reference's OWN class.

Authoring container only (needs /root/reference):  python tests/golden/make_paggrnn_golden.py
Eval mode (its dropouts are identities), CPU fp32; inputs, state_dict, logits and every parameter
gradient of sum(logits * G) are saved to tests/golden/paggrnn_<tag>.npz.
"""
import sys
import warnings
import zlib
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pagg_oracle as po  # noqa: E402

OUT = Path(__file__).resolve().parent
warnings.filterwarnings("ignore")
CASES = [("small", 60, 24, 32, 5, 5, 13), ("h128", 150, 40, 128, 7, 8, 37)]   # tag, N, F, H, C, nw, S (walk_len is 4 in this class)


def main():
    cls = po.load_reference_pagg_rnn_class()
    for tag, N, Fd, H, C, nw, S in CASES:
        L = 4
        g = torch.Generator().manual_seed(zlib.crc32(f"paggrnn-{tag}".encode()))
        torch.manual_seed(4321)
        model = cls(Fd, H, C, N).eval()
        X = torch.randn(N, Fd, generator=g)
        mask = torch.zeros(N, dtype=torch.bool)
        mask[torch.randperm(N, generator=g)[:S]] = True
        neis = torch.randint(0, N, (S, nw * L), generator=g)
        neis.view(S, nw, L)[:, :, 0] = torch.nonzero(mask).view(S, 1)
        ltype = torch.randint(0, L, (S, nw, L), generator=g)
        G = torch.randn(S, C, generator=g)
        y = model(X, neis, nw, L, mask, ltype, torch.arange(S * nw * L))
        (y * G).sum().backward()
        out = dict(N=N, F=Fd, H=H, C=C, nw=nw, L=L, S=S, X=X.numpy(), mask=mask.numpy(), neis=neis.numpy().astype(np.int32),
                   layer_type=ltype.numpy().astype(np.uint8), G=G.numpy(), logits=y.detach().numpy())
        for k, v in model.state_dict().items():
            out["param:" + k] = v.numpy()
        for k, p in model.named_parameters():
            out["grad:" + k] = p.grad.numpy()
        np.savez_compressed(OUT / f"paggrnn_{tag}.npz", **out)
        print(tag, "logits", y.shape, float(y.abs().mean()))


if __name__ == "__main__":
    main()
