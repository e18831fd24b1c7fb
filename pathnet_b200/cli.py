"""CLI shims with the reference sampler's argv, cwd-relative paths and file
names (preprocess/gen_merw.cpp:125-213, gen_epoch_merw.cpp):

    python -m pathnet_b200.cli gen_merw       <data_name> <path_num> <path_len> [--seed N] [--epochs N]
    python -m pathnet_b200.cli gen_epoch_merw <data_name> <path_num> <path_len> [--seed N] [--epochs N]
    python -m pathnet_b200.cli gen | gen_epoch <data_name> <path_num> <path_len> ...   uniform-random-walk variants
        (preprocess/gen.cpp reads ../edge_input/<name>_nsl.in -> ./<name>_<nw>_<L>_nsl.txt;
         preprocess/gen_epoch.cpp reads ../edge_input/<name>.in (2 columns) -> ./<name>_<nw>_<L>_<epoch>.txt)

Run from `preprocess/`: reads `../edge_input/<name>.in`, writes
`./<name>_<nw>_<L>_merw.txt` (gen_merw) or `./<name>_<nw>_<L>_<epoch>_merw.txt`
(gen_epoch_merw).  Output is byte-identical to the reference binary's for the
same srand() seed; --seed defaults to time(0) like gen_merw.cpp:161.
"""
from __future__ import annotations

import sys
import time


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("gen_merw", "gen_epoch_merw", "gen", "gen_epoch"):
        print(__doc__, file=sys.stderr)
        return 2
    tool, rest = argv[0], argv[1:]
    seed, epochs, device, pos = None, 1000, "cuda:0", []
    i = 0
    while i < len(rest):
        if rest[i] == "--seed":
            seed = int(rest[i + 1]); i += 2
        elif rest[i] == "--epochs":
            epochs = int(rest[i + 1]); i += 2
        elif rest[i] == "--device":
            device = rest[i + 1]; i += 2
        else:
            pos.append(rest[i]); i += 1
    if len(pos) != 3:                                  # gen_merw.cpp:128-132
        print("ERROR: Incorrect number of parameters. ", file=sys.stderr)
        return 0
    name, nw, L = pos[0], int(pos[1]), int(pos[2])
    uniform = tool in ("gen", "gen_epoch")                 # preprocess/gen.cpp / gen_epoch.cpp (uniform random walk)
    # gen.cpp:57-60 opens ../edge_input/<name>_nsl.in; gen_epoch.cpp:52-54 and both MERW tools open ../edge_input/<name>.in
    fin = f"../edge_input/{name}_nsl.in" if tool == "gen" else f"../edge_input/{name}.in"
    if seed is None:
        seed = int(time.time())
    from . import sampler
    from ._lib import PnError
    suffix = "nsl" if uniform else "merw"
    if tool in ("gen_merw", "gen"):                    # :156-157 (the per-epoch variants print nothing)
        print(f"File input: {fin}")
        print(f"File output: ./{name}_{nw}_{L}_{suffix}.txt")
    if uniform:
        n_, u_, v_ = sampler.read_edge_input_nsl(fin)
        g = sampler.MerwGraph(n_, u_, v_, device=device)
        if tool == "gen":
            print(f"{g.n}", file=sys.stderr)           # gen.cpp:81
        else:
            print(f"{n_} {len(u_)}")                   # gen_epoch.cpp:64 prints "n m" to stdout before the first freopen
    else:
        g = sampler.MerwGraph.from_file(fin, device=device)
        print(f"{name}: {g.n}", file=sys.stderr)      # gen_merw.cpp:164
    batch = max(1, min(epochs, (64 << 20) // max(1, g.n * nw * L)))
    out = open(f"./{name}_{nw}_{L}_{suffix}.txt", "wb") if tool in ("gen_merw", "gen") else None
    try:
        for e0 in range(0, epochs, batch):
            e1 = min(epochs, e0 + batch)
            try:
                nodes, types = g.sample(nw, L, epochs=e1 - e0, seed=seed, epoch_begin=e0)
            except PnError as ex:
                if ex.code == -3:                      # :84-87
                    print("ERROR:: A.size() == 0 in Alias Table", file=sys.stderr)
                    return 0
                raise
            nodes, types = nodes.cpu().numpy(), types.cpu().numpy()      # int32 / uint8, C-contiguous
            if out is not None:
                out.write(memoryview(sampler.format_paths_array(nodes, types)))
            else:
                for e in range(e0, e1):
                    with open(f"./{name}_{nw}_{L}_{e}_merw.txt" if not uniform else f"./{name}_{nw}_{L}_{e}.txt", "wb") as f:
                        f.write(memoryview(sampler.format_paths_array(nodes[e - e0], types[e - e0])))
    finally:
        if out is not None:
            out.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
