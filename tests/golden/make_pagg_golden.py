"""Generate the committed PAGG fixtures from the reference's OWN classes.

Authoring container only (needs /root/reference):
    python tests/golden/make_pagg_golden.py
The classes PathNet / PathNet_homo are AST-extracted from
/root/reference/PathNet_run.py (oracle.pagg_oracle.load_reference_classes),
run on seeded inputs in eval mode (dropout identity) on CPU fp32, and inputs,
state_dict, logits and every parameter gradient of sum(logits * G) are saved.
Weights and features are rounded to fp16-representable values so the same
fixture is exact input for the tensor-core math mode.
Output: tests/golden/pagg_<variant>_<tag>.npz
"""
import sys
import zlib
import warnings
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pagg_oracle as po  # noqa: E402

OUT = Path(__file__).resolve().parent
warnings.filterwarnings("ignore")

CASES = [
    # tag, N, F, H, C, nw, L, S
    ("small", 60, 24, 32, 5, 5, 4, 13),
    ("h128", 150, 40, 128, 7, 8, 4, 37),
    ("l6", 90, 20, 64, 3, 6, 6, 29),
]


def q16(t):
    return t.half().float()


def main():
    ns = po.load_reference_classes(dropout=0.0)
    for variant, cls in (("hetero", "PathNet"), ("homo", "PathNet_homo")):
        for tag, N, Fd, H, C, nw, L, S in CASES:
            g = torch.Generator().manual_seed(zlib.crc32(f"{variant}-{tag}".encode()))
            torch.manual_seed(1234)
            model = ns[cls](Fd, H, C, L).eval()
            with torch.no_grad():
                for p in model.parameters():
                    p.copy_(q16(p * 1.5))
            X = q16(torch.randn(N, Fd, generator=g))
            mask = torch.zeros(N, dtype=torch.bool)
            mask[torch.randperm(N, generator=g)[:S]] = True
            neis = torch.randint(0, N, (S, nw * L), generator=g)
            neis.view(S, nw, L)[:, :, 0] = torch.nonzero(mask).view(S, 1)   # walks start at their node
            ltype = torch.randint(0, L, (S, nw, L), generator=g)
            ltype[:, :, 0] = 0
            G = q16(torch.randn(S, C, generator=g))
            y = model(X, neis, nw, L, mask, ltype, torch.arange(S * nw * L))
            (y * G).sum().backward()
            out = dict(variant=variant, N=N, F=Fd, H=H, C=C, nw=nw, L=L, S=S,
                       X=X.numpy().astype(np.float16), mask=mask.numpy(), neis=neis.numpy().astype(np.int32),
                       layer_type=ltype.numpy().astype(np.uint8), G=G.numpy().astype(np.float16),
                       logits=y.detach().numpy())
            for k, v in model.state_dict().items():
                out["param:" + k] = v.numpy().astype(np.float16)
            for k, p in model.named_parameters():
                out["grad:" + k] = p.grad.numpy()
            np.savez_compressed(OUT / f"pagg_{variant}_{tag}.npz", **out)
            print(variant, tag, "logits", y.shape, float(y.abs().mean()))


if __name__ == "__main__":
    main()
