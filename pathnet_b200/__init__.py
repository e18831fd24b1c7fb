"""pathnet_b200 — B200-native MERW sampler + PAGG operator behind the
reference's own surfaces."""
from .sampler import (MerwGraph, gen_merw, read_edge_input, write_paths_txt, format_paths_txt,  # noqa: F401
                      parse_paths_txt, read_paths_txt)

try:  # PAGG needs the same library; keep the sampler importable on its own
    from .pagg import PathNet, PathNet_homo, PAGG, PAGG_RNN, pagg_forward  # noqa: F401
except ImportError:  # pragma: no cover
    pass
