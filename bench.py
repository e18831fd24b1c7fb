#!/usr/bin/env python
"""Benchmark of the PathNet hot path (MERW sampler -> PAGG fwd+bwd) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload pubmed|cora|bgp] [--scaling weak|strong] [--math tc|tc3|fp32]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Headline (`value`): PAGG paths aggregated per second, forward+backward of one
training step (the reference's step: PathNet_run.py:345-352 — forward, CE loss,
backward, Adam step), on the pubmed-shaped configuration of BASELINE.json
(N=19 717 nodes, F=500, C=3, H=128, path_num=40, path_len=4, S=9 463 rows ->
P=378 520 paths per step and per GPU).  The second metric of BASELINE.json,
MERW walks/sec (labels included), is measured in the same run on the same
graph and reported under "merw".  Data are synthetic (seeded) at those shapes.

A "step" = one pass over one batch through pathnet_b200.train.TrainStep (the
whole step captured as one CUDA graph): inputs resident in HBM for `value`;
pinned HOST index tensors copied in and the loss read back every step for
`e2e`.  Prints ONE JSON line (rank 0).

Math modes: the headline runs the fast tensor-core mode ("tc": single fp16 pass, logits within ~1e-3 of fp32 math); the
same step is timed in the split-precision tensor-core mode ("tc3": the fp32 mode's 1e-4 parity bar, what the modules'
"auto" selects) under `e2e.strict_tc3` and in the CUDA-core fp32 mode under `e2e.strict`.  `gpu_status_quo` = the
reference's own classes on the same GPU through stock PyTorch; `cpu_baseline` / `--impl reference` = on the host CPU.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # BASELINE.json configs[2] (the default; the metric is quoted on it), [1] and [3]; shapes from SURVEY §8(d)
    "pubmed": dict(name="pubmed-shaped", n=19717, undirected_edges=44324, F=500, C=3, H=128, nw=40, L=4, S=9463,
                   variant="homo", dropout=0.7),
    "cora": dict(name="cora-shaped", n=2708, undirected_edges=5278, F=1433, C=7, H=128, nw=40, L=4, S=1192,
                 variant="homo", dropout=0.7),
    "bgp": dict(name="bgp-shaped", n=63977, undirected_edges=142814, F=287, C=8, H=128, nw=40, L=6, S=4884,
                variant="hetero", dropout=0.7),
}
WORKLOAD = dict(WORKLOADS["pubmed"])
MERW_EPOCHS_PER_STEP = 16
STAGED_REF = ROOT / "oracle" / "_ref" / "pathnet" / "PathNet_run.py"


def config_of(world, scaling):
    """The SAME dict in both arms (the driver compares them): what is computed, not how."""
    w = WORKLOAD
    rows = f"S={w['S']} rows" + ("/GPU" if scaling == "weak" and world > 1 else " in total" if world > 1 else "")
    return {"workload": f"{w['name']} N={w['n']} F={w['F']} H={w['H']} C={w['C']} path_num={w['nw']} path_len={w['L']} {rows} "
                        f"-> {w['S'] * w['nw']} paths/step ({'PathNet_homo' if w['variant'] == 'homo' else 'PathNet'}, train mode, "
                        f"dropout {w['dropout']}, CE + Adam)",
            "parallelism": f"dp{world} ({scaling} scaling)" if world > 1 else "single GPU"}


# --------------------------------------------------------------------------- data
def shaped_graph(seed=0):
    """n nodes, rows like edge_input/pubmed.in: every directed edge and every self loop appears
    twice, skewed degrees, p > 0 with row sums 2."""
    w = WORKLOAD
    rng = np.random.default_rng(seed)
    n, ne = w["n"], w["undirected_edges"]
    weights = 1.0 / np.sqrt(np.arange(1, n + 1))
    weights /= weights.sum()
    a = rng.choice(n, size=ne, p=weights)
    b = rng.integers(0, n, ne)
    b = np.where(a == b, (b + 1) % n, b)
    perm = rng.permutation(n)  # decorrelate node id from degree
    a, b = perm[a], perm[b]
    loops = np.arange(n)
    u = np.concatenate([a, b, a, b, loops, loops]).astype(np.int32)
    v = np.concatenate([b, a, b, a, loops, loops]).astype(np.int32)
    order = rng.permutation(len(u))
    u, v = u[order], v[order]
    p = rng.random(len(u)) + 0.05
    rowsum = np.bincount(u, weights=p, minlength=n)
    p = 2.0 * p / rowsum[u]
    return n, u, v, p


def synthetic_inputs(rank=0):
    w = WORKLOAD
    g = torch.Generator().manual_seed(100 + rank)
    X = torch.randn(w["n"], w["F"], generator=torch.Generator().manual_seed(0))
    mask = torch.zeros(w["n"], dtype=torch.bool)
    mask[torch.randperm(w["n"], generator=g)[: w["S"]]] = True
    y = torch.randint(0, w["C"], (w["n"],), generator=torch.Generator().manual_seed(1))
    return X, mask, y


# ------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows)}


def ncu_traffic(slot):
    """DRAM bytes per launch of a kernel from the newest committed `ncu --set full` capture
    (profiles/r0X/ncu_traffic.json, written by tools/ncu_summary.py), or None."""
    for rnd in ("r02", "r01"):
        f = ROOT / "profiles" / rnd / "ncu_traffic.json"
        if f.exists():
            d = json.loads(f.read_text()).get(slot)
            if d:
                return d["dram_bytes_per_launch"], rnd
    return None, None


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------ reference arm
def reference_classes(device, dropout):
    """The reference's OWN PathNet / PathNet_homo classes, AST-extracted from the staged, unmodified script
    (oracle/_ref/pathnet/PathNet_run.py, see oracle/stage_reference.py); None when it did not travel."""
    if not STAGED_REF.exists():
        return None
    import ast
    import torch.nn as nn
    import torch.nn.functional as F
    tree = ast.parse(STAGED_REF.read_text())
    keep = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("PathNet", "PathNet_homo")]

    class MessagePassing(nn.Module):   # the reference subclasses torch_geometric's class and never calls propagate()
        def __init__(self, *a, **k):
            super().__init__()

    ns = dict(torch=torch, nn=nn, F=F, MessagePassing=MessagePassing, device=device, dropout=dropout)
    exec(compile(ast.Module(body=keep, type_ignores=[]), "PathNet_run.py", "exec"), ns)
    return ns


def reference_train_step_fn(device, rows_limit=None, seed=5):
    """One training step of the reference (PathNet_run.py:345-352) with its own classes on `device`: returns
    (step_fn, paths_per_step, description).  Falls back to the oracle port when the script is not staged."""
    w = WORKLOAD
    X, mask, y = synthetic_inputs()
    rows = torch.nonzero(mask).reshape(-1)
    if rows_limit is not None:
        rows = rows[:rows_limit]
    m2 = torch.zeros_like(mask)
    m2[rows] = True
    S, nw, L = rows.numel(), w["nw"], w["L"]
    g = torch.Generator().manual_seed(seed)
    neis = torch.randint(0, w["n"], (S, nw * L), generator=g)          # CPU int64, as the reference holds them
    lt = torch.randint(0, L, (S, nw, L), generator=g)
    ns = reference_classes(device, w["dropout"])
    torch.manual_seed(0)
    Xd, yd = X.to(device), y[m2].to(device)
    if ns is not None:
        cls = ns["PathNet_homo"] if w["variant"] == "homo" else ns["PathNet"]
        model = cls(w["F"], w["H"], w["C"], L).to(device).train()
        opt = torch.optim.Adam(model.parameters(), lr=0.005, weight_decay=0.0005)
        lossf = torch.nn.CrossEntropyLoss()
        m2_np = m2.numpy()

        def step():
            indxx = torch.arange(S * nw * L, dtype=torch.long, device=device)       # :342-343
            out = model(Xd, neis, nw, L, m2_np, lt, indxx)                            # :345-347
            loss = lossf(out, yd)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return float(loss.item())
        return step, S * nw, "reference", "unmodified PathNet_run.py classes (AST-extracted from oracle/_ref/pathnet), torch Adam + CrossEntropyLoss"
    from oracle import pagg_oracle as po
    import pathnet_b200.pagg as pg
    mod = (pg.PathNet_homo if w["variant"] == "homo" else pg.PathNet)(w["F"], w["H"], w["C"], L)
    params = {k: v.detach().clone().to(device).requires_grad_(True) for k, v in mod.state_dict().items()}
    opt = torch.optim.Adam(params.values(), lr=0.005, weight_decay=0.0005)
    p_drop = w["dropout"]

    def step():
        ms = (torch.rand(L, S * nw, w["H"], device=device) >= p_drop).float() / (1 - p_drop)
        mc = (torch.rand(S, 2 * w["H"], device=device) >= p_drop).float() / (1 - p_drop)
        out = po.pagg_forward(w["variant"], params, Xd, neis.to(device), nw, L, m2, lt.to(device), mask_seq=ms, mask_cat=mc)
        loss = torch.nn.functional.cross_entropy(out, yd)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss.item())
    return step, S * nw, "port", "oracle/pagg_oracle.py (torch restatement of the reference forward), torch Adam"


def cpu_pagg_baseline(steps=2, warmup=1, budget_s=150.0, threads=None):
    """The reference PAGG train step on the host cores: full S when it fits the time budget, else a bounded
    prefix of the rows (same N, F, H, nw, L)."""
    w = WORKLOAD
    # all the host threads it can use, but not past the point where torch's CPU kernels stop scaling on these
    # per-op shapes (measured on the box: 128 threads were 40x slower than 8-32)
    threads = threads or min(os.cpu_count() or 1, 32)
    torch.set_num_threads(threads)
    rows_limit = None
    step, paths, kind, what = reference_train_step_fn("cpu", rows_limit)
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    if first * (steps + warmup) > budget_s:       # too slow for the budget: shrink the sample proportionally
        rows_limit = max(64, int(w["S"] * budget_s / (first * (steps + warmup + 1))))
        step, paths, kind, what = reference_train_step_fn("cpu", rows_limit)
        step()
    for _ in range(max(0, warmup - 1)):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    S_used = paths // w["nw"]
    base = dict(value=paths / mean, unit="paths/s", cores=threads, kind=kind,
                sample=f"{S_used} of {w['S']} rows ({paths} paths/step), mean of {steps} step(s) after {warmup} warm-up, {what}, "
                       f"torch CPU fp32 on {threads} threads (best {paths / min(times):.0f} paths/s)")
    return base, mean, paths


def gpu_status_quo(dev, steps=5, warmup=2):
    """BASELINE.md §3 PAGG-2: the reference module itself on the same B200 through stock PyTorch (cuBLAS / cuDNN)."""
    w = WORKLOAD
    try:
        step, paths, kind, what = reference_train_step_fn(dev)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": paths / (ms * 1e-3), "unit": "paths/s", "ms_per_step": ms, "kind": kind, "what": what,
                "note": f"stock PyTorch {torch.__version__} ops on the same GPU, host int64 index tensors in + loss.item() per step "
                        f"(its own e2e), matmul TF32 {'on' if torch.backends.cuda.matmul.allow_tf32 else 'off'}, cuDNN TF32 "
                        f"{'on' if torch.backends.cudnn.allow_tf32 else 'off'}"}
    except Exception as ex:   # out of memory on the 3 GB stack etc.: report, never fail the bench
        return {"unavailable": f"{type(ex).__name__}: {str(ex)[:200]}"}


def cpu_merw_baseline(graph, epochs=1):
    """MERW on the host: the UNMODIFIED reference binary (oracle/_ref, 1 thread
    by construction) when it travelled with the repo, else the C oracle port."""
    from oracle import merw_oracle as mo
    w = WORKLOAD
    n, u, v, p = graph
    walks = n * w["nw"] * epochs
    if mo.have_ref():
        import tempfile
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
            f = Path(d) / "bench.in"
            mo.write_edge_input(f, n, u, v, p)
            # two timings: E epochs and 2E epochs of the per-epoch binary; the
            # difference is the steady-state walk+print rate without setup
            t = []
            for e in (epochs, 2 * epochs):
                t0 = time.perf_counter()
                mo.run_reference(f, w["nw"], w["L"], 12345, epochs=e, per_epoch=True, name="bench")
                t.append(time.perf_counter() - t0)
            steady = max(t[1] - t[0], 1e-9)
        return dict(value=walks / steady, unit="walks/s", cores=1, kind="reference",
                    sample=f"{epochs} epoch(s) = {walks} walks, oracle/_ref/gen_epoch_merw steady state "
                           f"(walk loop + printf to /dev/shm; setup {max(t[0] - steady, 0):.2f}s excluded)")
    g = mo.OracleGraph(n, u, v, p)
    t0 = time.perf_counter()
    g.sample(w["nw"], w["L"], 12345, 0, epochs)
    dt = time.perf_counter() - t0
    return dict(value=walks / dt, unit="walks/s", cores=1, kind="port",
                sample=f"{epochs} epoch(s) = {walks} walks, oracle/merw_oracle.c incl. BFS labels, no text output")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOAD
    base, mean, paths = cpu_pagg_baseline(steps=max(1, args.steps), warmup=max(1, args.warmup), budget_s=200.0)
    merw = cpu_merw_baseline(shaped_graph())
    line = {
        "impl": "reference", "metric": "pagg_paths_per_sec_fwd_bwd", "value": base["value"], "unit": "paths/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(max(1, args.gpus), args.scaling),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "merw": {"metric": "merw_walks_per_sec", "value": merw["value"], "unit": "walks/s", "cpu_baseline": merw},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------ our arm
def main():
    if os.environ.get("PN_BENCH_DEADLINE"):   # debug: dump every thread's Python stack and exit instead of hanging past N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["PN_BENCH_DEADLINE"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default=os.environ.get("PN_BENCH_MATH", "tc"), choices=["fp32", "tc", "tc3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the strict / module-API / status-quo / MERW legs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="multi-GPU: weak = every rank its own S rows; strong = the S rows of ONE batch sharded across ranks")
    ap.add_argument("--workload", default="pubmed", choices=sorted(WORKLOADS),
                    help="pubmed (default, BASELINE configs[2]) | cora (configs[1]) | bgp (configs[3], hetero, L=6)")
    args = ap.parse_args()
    WORKLOAD.clear()
    WORKLOAD.update(WORKLOADS[args.workload])
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    import pathnet_b200 as pb
    from pathnet_b200 import _lib
    from pathnet_b200.dist import FlatGradAllReduce, shard_bounds
    from pathnet_b200.train import TrainStep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOAD
    peaks = measured_peaks()
    L_, nw, S, H = w["L"], w["nw"], w["S"], w["H"]
    strong = args.scaling == "strong" and world > 1
    lo, hi = shard_bounds(S, rank, world) if strong else (0, S)
    S_local = hi - lo
    P_local, P_job = S_local * nw, (S * nw if strong else S * nw * world)
    cls = pb.PathNet_homo if w["variant"] == "homo" else pb.PathNet

    # ---- data: graph + one epoch of OUR sampled paths per rank ----
    graph = shaped_graph()
    g = pb.MerwGraph(*graph, device=dev)
    X, mask, y = synthetic_inputs(0 if strong else rank)
    nodes, types = g.sample(nw, L_, epochs=1, seed=12345, epoch_begin=0 if strong else rank)
    maskd = mask.to(dev)
    neis_dev = nodes[0].view(w["n"], nw * L_)[maskd].contiguous()           # int32 [S, nw*L]: the sampler's own dtypes,
    ltype_dev = types[0][maskd].contiguous()                                #   uint8 [S, nw, L]: resident in HBM
    neis_host = neis_dev.cpu().pin_memory()                                 # 4 + 1 bytes per path position in pinned host memory
    ltype_host = ltype_dev.cpu().pin_memory()
    Xd, y_rows = X.to(dev), y[mask]
    S_global = S if strong else S * world

    def make_step(math, graph_mode=True, seed=7):
        torch.manual_seed(0)
        model = cls(w["F"], H, w["C"], L_, dropout=w["dropout"], math_mode=math).to(dev).train()
        ts = TrainStep(model, Xd, mask, y_rows, nw, L_, lr=0.005, weight_decay=0.0005, seed=seed + rank, graph=graph_mode,
                       s_global=S_global, shard=(lo, hi) if strong else None)     # Adam as PathNet_run.py:295
        return model, ts

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- value: inputs resident in HBM, K graph-launched steps ----
    model, ts = make_step(args.math)
    for _ in range(args.warmup):
        ts.step(neis_dev, ltype_dev)
    with ClockSampler(local) as clocks:
        total_ms = timed(lambda: ts.step(neis_dev, ltype_dev), args.steps)
    ms_per_step = total_ms / args.steps
    value = P_job / (ms_per_step * 1e-3)
    launches = ts.launches_per_step * args.steps      # kernels of libpathnet_b200.so per step (graph nodes replayed) x steps

    # ---- e2e: the public training-step call with pinned HOST index tensors in and the loss out, every step.  The
    #      H2D copy of step k+1 is issued (prefetch) before step k's loss is read, so it overlaps step k's kernels;
    #      every step's copy and read-back happen inside the timed region ----
    def e2e_step():
        ts.step()
        ts.prefetch(neis_host, ltype_host)
        return float(ts.loss.item())
    ts.prefetch(neis_host, ltype_host)
    for _ in range(2):
        e2e_step()
    e2e_ms = timed(e2e_step, args.steps) / args.steps
    ts.step()   # consume the last prefetch
    h2d = neis_host.numel() * neis_host.element_size() + ltype_host.numel() * ltype_host.element_size()
    e2e = {"value": P_job / (e2e_ms * 1e-3), "unit": "paths/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms,
           "note": "pathnet_b200.train.TrainStep: prefetch(pinned host int32 node ids + uint8 labels, the dtypes gen_merw() hands over) "
                   "-> step() (one CUDA graph: canonicalize, forward, CE, backward, Adam) -> loss.item(); the copy of the next step's "
                   "indices overlaps the current step; X stays resident as in the reference (:300)"}

    extras = {}
    roofline = None
    merw = None
    if not args.no_extras:
        # ---- the reference's own calling convention: module.forward with CPU int64 tensors, torch CE + Adam (autograd path) ----
        neis64, ltype64 = neis_host.long().pin_memory(), ltype_host.long().pin_memory()
        torch.manual_seed(0)
        model2 = cls(w["F"], H, w["C"], L_, dropout=w["dropout"], math_mode=args.math).to(dev).train()
        opt = torch.optim.Adam(model2.parameters(), lr=0.005, weight_decay=0.0005, fused=True)
        reducer = FlatGradAllReduce(model2.parameters(), sink=model2.grad_sink) if world > 1 else None
        lossf = torch.nn.CrossEntropyLoss(reduction="sum")
        yd = y_rows.to(dev)
        if not strong:
            def module_step():
                out = model2(Xd, neis64, nw, L_, mask, ltype64, None)
                loss = lossf(out, yd) / S_global
                opt.zero_grad()
                loss.backward()
                if reducer is not None:
                    reducer()
                opt.step()
                return loss.item()
            for _ in range(2):
                module_step()
            mod_ms = timed(module_step, args.steps) / args.steps
            extras["module_api"] = {"value": P_job / (mod_ms * 1e-3), "unit": "paths/s", "ms_per_step": mod_ms,
                                    "h2d_bytes_per_step": int(neis64.numel() * 16),
                                    "note": "nn.Module.forward with CPU int64 neis/layer_type exactly as PathNet_run.py:345-352 passes them, "
                                            "torch CrossEntropyLoss + autograd + fused torch Adam, loss.item() every step"
                                            + (f"; zero-copy flat all-reduce used on {reducer.zero_copy_steps} steps" if reducer else "")}
        del model2, opt

        # ---- roofline of the dominant kernels (in-stream CUDA events inside the library; eager launches, not the graph) ----
        _, ts_prof = make_step(args.math, graph_mode=False)
        for _ in range(2):
            ts_prof.step(neis_dev, ltype_dev)
        _lib.profile_enable(True)
        for _ in range(3):
            ts_prof.step(neis_dev, ltype_dev)
        torch.cuda.synchronize()
        prof = {k: _lib.profile_read(k) for k in ("lstm_fwd", "lstm_bwd", "wgrad")}
        _lib.profile_enable(False)
        del ts_prof
        flop_pass = 16.0 * L_ * H * H * P_local      # 16 L H^2 per path per LSTM GEMM pass (forward gates, dgrad, wgrad)
        peak = peaks["bf16_sustained"] or peaks["bf16"]
        kern = {}
        for name in ("lstm_fwd", "lstm_bwd", "wgrad"):
            ms, n = prof[name]
            if n:
                kern[name] = {"ms": ms / n, "tflops": flop_pass / (ms / n * 1e-3) / 1e12}
                tr, rnd = ncu_traffic(name) if args.workload == "pubmed" and args.math == "tc" else (None, None)
                if tr:   # second roof: the kernel's DRAM traffic (committed ncu capture of this workload) over the live duration
                    kern[name]["dram_gbs"] = tr / (ms / n * 1e-3) / 1e9
                    kern[name]["hbm_frac"] = kern[name]["dram_gbs"] / peaks["hbm"]
                kern[name]["tensor_frac"] = kern[name]["tflops"] / peak
        dom = max(kern, key=lambda k: kern[k]["ms"]) if kern else None
        if dom:
            tr, rnd = ncu_traffic(dom)
            roofline = {"kernel": dom, "bound": "tensor", "achieved": kern[dom]["tflops"], "peak": peak,
                        "unit": "TFLOP/s", "frac": kern[dom]["tflops"] / peak, "traffic": tr,
                        "traffic_note": f"DRAM read+write bytes per launch from the committed ncu --set full capture (profiles/{rnd})",
                        "peak_source": f"{peaks['src']} bf16 sustained under the power cap (MEASURED_PEAKS.json); these kernels run at full clock",
                        "math_mode": args.math, "kernels": kern,
                        "whole_step_tflops": 3 * flop_pass / (ms_per_step * 1e-3) / 1e12,
                        "note": "algorithmic FLOPs = 16*L*H^2 per path per LSTM GEMM pass (SURVEY 8d): forward gates, "
                                "backward dgrad, backward wgrad; math_mode tc = tcgen05 fp16 operands / fp32 accumulate, "
                                "fp32 = CUDA-core FMA (strict parity mode).  kernels[*].hbm_frac = the same kernels against "
                                "the HBM roof (saved-activation / operand-image traffic)"}

        # ---- strict-parity mode (CUDA-core fp32, 1e-4 against the reference) on the same step ----
        if args.math == "tc" and world == 1:
            _, ts32 = make_step("fp32")
            for _ in range(2):
                ts32.step(neis_dev, ltype_dev)
            n32 = max(3, min(args.steps, 5))
            ms32 = timed(lambda: ts32.step(neis_dev, ltype_dev), n32) / n32
            ts32.prefetch(neis_host, ltype_host)

            def e2e32():
                ts32.step()
                ts32.prefetch(neis_host, ltype_host)
                return float(ts32.loss.item())
            e2e32()
            ms32e = timed(e2e32, n32) / n32
            extras["strict"] = {"math_mode": "fp32", "dtype": "f32", "value": P_job / (ms32 * 1e-3), "ms_per_step": ms32,
                                "e2e": P_job / (ms32e * 1e-3), "e2e_ms_per_step": ms32e, "unit": "paths/s", "steps": n32,
                                "note": "same step in the strict-parity math mode (CUDA-core fp32 FMA; logits within 1e-4 of the reference's fp32 math)"}
            del ts32

            # ---- split-precision tensor-core mode: the strict mode's parity bar on tcgen05 (three fp16 passes per product) ----
            _, ts3 = make_step("tc3")
            for _ in range(3):
                ts3.step(neis_dev, ltype_dev)
            n3 = max(5, args.steps)
            ms3 = timed(lambda: ts3.step(neis_dev, ltype_dev), n3) / n3
            ts3.prefetch(neis_host, ltype_host)

            def e2e3():
                ts3.step()
                ts3.prefetch(neis_host, ltype_host)
                return float(ts3.loss.item())
            e2e3()
            ms3e = timed(e2e3, n3) / n3
            extras["strict_tc3"] = {"math_mode": "tc3", "dtype": "2xf16 split operands x 3 passes, f32 accumulate (fp32-grade)",
                                    "value": P_job / (ms3 * 1e-3), "ms_per_step": ms3, "e2e": P_job / (ms3e * 1e-3),
                                    "e2e_ms_per_step": ms3e, "unit": "paths/s", "steps": n3,
                                    "note": "same step in the split-precision tensor-core mode: LSTM operands as hi + lo fp16 terms, "
                                            "A W = A_hi W_hi + A_lo W_hi + A_hi W_lo on tcgen05, ex2/rcp gates, projections on the fp32 "
                                            "kernels; held to the fp32 mode's parity bar (tests/test_gpu_tc3.py)"}
            del ts3

        # ---- MERW walks/sec (labels included), same graph ----
        E = MERW_EPOCHS_PER_STEP
        out_n = torch.empty((E, w["n"], nw, L_), dtype=torch.int32, device=dev)
        out_t = torch.empty((E, w["n"], nw, L_), dtype=torch.uint8, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for i in range(3):
            g.sample(nw, L_, epochs=E, seed=777, epoch_begin=i * E, out=(out_n, out_t))
        merw_ms = []
        _lib.profile_enable(True)
        for i in range(args.steps):
            flush.fill_(i & 0xFF)          # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.sample(nw, L_, epochs=E, seed=777, epoch_begin=(3 + i) * E, out=(out_n, out_t))
            e1.record()
            torch.cuda.synchronize()
            merw_ms.append(e0.elapsed_time(e1))
        kms, kn = _lib.profile_read("merw_walk")
        _lib.profile_enable(False)
        walks = E * w["n"] * nw
        merw_step = sum(merw_ms) / len(merw_ms)
        t_m = torch.tensor([merw_step], device=dev)
        if world > 1:
            dist.all_reduce(t_m, op=dist.ReduceOp.MAX)
        merw_step = float(t_m.item())
        # e2e of the sampler = the hand-off the next stage consumes: the tensors stay on the device (gen_merw() -> TrainStep);
        # for the reference's text boundary, D2H + the threaded text writer are timed by tools/bench_handoff.py
        bytes_per_walk = (L_ - 1) * (8 + 16) + L_ * (4 + 1)   # 92 B at L = 4
        mtr, mrnd = ncu_traffic("merw_walk")
        merw = {"metric": "merw_walks_per_sec", "value": world * walks / (merw_step * 1e-3), "unit": "walks/s",
                "ms_per_step": merw_step, "walks_per_step": walks, "epochs_per_step": E,
                "labels_included": True, "rng": "glibc_compat (bit-exact srand/rand stream)", "l2_flush": True,
                "roofline": {"kernel": "k_merw_walk", "bound": "hbm",
                             "achieved": walks * bytes_per_walk / ((kms / max(kn, 1)) * 1e-3) / 1e9 if kn else None,
                             "peak": peaks["hbm"], "unit": "GB/s",
                             "frac": (walks * bytes_per_walk / ((kms / kn) * 1e-3) / 1e9 / peaks["hbm"]) if kn else None,
                             "traffic": mtr, "bytes_per_walk": bytes_per_walk,
                             "peak_source": f"{peaks['src']} HBM copy bandwidth"}}
        del out_n, out_t, flush

    line = {
        "metric": "pagg_paths_per_sec_fwd_bwd", "value": value, "unit": "paths/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "tc": "f16xf32acc", "tc3": "2xf16-split x3 passes, f32acc"}[args.math], "data": "synthetic",
        "config": config_of(world, args.scaling),
        "math_mode": args.math,
        "l2": "per-step working set ~9 GB >> 126 MB L2 (no explicit flush for PAGG); MERW leg flushes L2 with a 256 MiB write between launches",
        "inputs": "paths drawn by the on-GPU MERW sampler (int32 node ids, uint8 labels); one CUDA graph per step",
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "launches_per_step": ts.launches_per_step,
        "roofline": roofline, "merw": merw,
    }
    if strong:
        line["rows_per_rank"] = [shard_bounds(S, r, world)[1] - shard_bounds(S, r, world)[0] for r in range(world)]
    line.update(extras)
    if roofline is not None and merw is not None:
        roofline["merw"] = {"value": merw["value"], "unit": "walks/s", "frac_of_hbm_roof": merw["roofline"]["frac"]}
    if "strict" in extras:
        e2e["strict"] = extras["strict"]
    if "strict_tc3" in extras:
        e2e["strict_tc3"] = extras["strict_tc3"]
    if world > 1:
        dist.barrier()
    if rank == 0:
        if not args.no_cpu_baseline and not args.no_extras and world == 1:
            del ts, model
            torch.cuda.empty_cache()
            line["gpu_status_quo"] = gpu_status_quo(dev)
            base, _, _ = cpu_pagg_baseline(steps=2, warmup=1, budget_s=40.0)
            base["gpu_status_quo"] = line["gpu_status_quo"]
            line["cpu_baseline"] = base
            if merw is not None:
                line["merw"]["cpu_baseline"] = cpu_merw_baseline(graph)
        print(json.dumps(line), flush=True)
    if world > 1:
        del ts, model          # the captured graphs (they hold NCCL kernels) go before the communicator
        from pathnet_b200.dist import shutdown
        shutdown()


if __name__ == "__main__":
    main()
