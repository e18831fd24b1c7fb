#!/usr/bin/env python
"""Benchmark of the PathNet hot path (MERW sampler -> PAGG fwd+bwd) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Headline (`value`): PAGG paths aggregated per second, forward+backward of one
training step (the reference's step: PathNet_run.py:345-352 — forward, CE loss,
backward, Adam step), on the pubmed-shaped configuration of BASELINE.json
(N=19 717 nodes, F=500, C=3, H=128, path_num=40, path_len=4, S=9 463 rows ->
P=378 520 paths per step and per GPU).  The second metric of BASELINE.json,
MERW walks/sec (labels included), is measured in the same run on the same
graph and reported under "merw".  Data are synthetic (seeded) at those shapes.

A "step" = one pass over one batch: inputs resident in HBM for `value`; host
index tensors copied in and the loss read back every step for `e2e`.
Prints ONE JSON line (rank 0).
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


# --------------------------------------------------------------------------- data
def pubmed_shaped_graph(seed=0):
    """n = 19 717, 216 730 rows like edge_input/pubmed.in: every
    directed edge and every self loop appears twice, skewed degrees, p > 0 with
    row sums 2."""
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
            self._stop.wait(0.2)

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
    """DRAM bytes per launch of a kernel from the committed `ncu --set full` capture
    (profiles/r01/ncu_traffic.json, written by tools/ncu_summary.py), or None."""
    f = ROOT / "profiles" / "r01" / "ncu_traffic.json"
    if f.exists():
        d = json.loads(f.read_text()).get(slot)
        if d:
            return d["dram_bytes_per_launch"]
    return None


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------ reference arm
def cpu_pagg_baseline(sample_rows=512, repeats=2, threads=None):
    """The oracle port of the reference PAGG train step (torch CPU fp32, all host
    threads) on a bounded sample of the same workload: the first `sample_rows`
    of the S rows (same N, F, H, nw, L)."""
    from oracle import pagg_oracle as po
    w = WORKLOAD
    # all the host threads it can use, but not past the point where torch's CPU
    # kernels stop scaling on these small per-op shapes (measured on the box:
    # 128 threads were 40x slower than 8-32)
    threads = threads or min(os.cpu_count(), 32)
    torch.set_num_threads(threads)
    X, mask, y = synthetic_inputs()
    rows = torch.nonzero(mask).reshape(-1)[:sample_rows]
    m2 = torch.zeros_like(mask)
    m2[rows] = True
    S, nw, L = rows.numel(), w["nw"], w["L"]
    g = torch.Generator().manual_seed(5)
    neis = torch.randint(0, w["n"], (S, nw * L), generator=g)
    lt = torch.randint(0, L, (S, nw, L), generator=g)
    import pathnet_b200.pagg as pg
    torch.manual_seed(0)
    mod = (pg.PathNet_homo if w["variant"] == "homo" else pg.PathNet)(w["F"], w["H"], w["C"], L)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    opt = torch.optim.Adam(params.values(), lr=0.005, weight_decay=0.0005)
    p_drop = w["dropout"]
    best = None
    for it in range(repeats + 1):
        t0 = time.perf_counter()
        ms = (torch.rand(L, S * nw, w["H"]) >= p_drop).float() / (1 - p_drop)
        mc = (torch.rand(S, 2 * w["H"]) >= p_drop).float() / (1 - p_drop)
        out = po.pagg_forward(w["variant"], params, X, neis, nw, L, m2, lt, mask_seq=ms, mask_cat=mc)
        loss = torch.nn.functional.cross_entropy(out, y[m2])
        opt.zero_grad()
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it > 0:
            best = dt if best is None else min(best, dt)
    return dict(value=S * nw / best, unit="paths/s", cores=threads, kind="port",
                sample=f"{S} of {w['S']} rows ({S * nw} paths) of the same workload, best of {repeats} after 1 warm-up, "
                       f"oracle/pagg_oracle.py (torch CPU fp32) train step incl. dropout masks + Adam"), best, S * nw


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
    base, best, paths = cpu_pagg_baseline(sample_rows=512, repeats=max(1, min(args.steps, 3)))
    merw = cpu_merw_baseline(pubmed_shaped_graph())
    line = {
        "impl": "reference", "metric": "pagg_paths_per_sec_fwd_bwd", "value": base["value"], "unit": "paths/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": best * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{w['name']} N={w['n']} F={w['F']} H={w['H']} C={w['C']} path_num={w['nw']} "
                               f"path_len={w['L']} S={w['S']} ({'PathNet_homo' if w['variant'] == 'homo' else 'PathNet'}, dropout {w['dropout']}); "
                               f"bounded sample: {paths} paths/step"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "merw": {"metric": "merw_walks_per_sec", "value": merw["value"], "unit": "walks/s", "cpu_baseline": merw},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default=os.environ.get("PN_BENCH_MATH", "tc"), choices=["fp32", "tc"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from pathnet_b200.dist import FlatGradAllReduce

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
    P = S * nw

    # ---- data: graph + one epoch of OUR sampled paths per rank (weak scaling) ----
    graph = pubmed_shaped_graph()
    g = pb.MerwGraph(*graph, device=dev)
    X, mask, y = synthetic_inputs(rank)
    nodes, types = g.sample(nw, L_, epochs=1, seed=12345, epoch_begin=rank)
    neis_host = nodes[0].view(w["n"], nw * L_).long().cpu()[mask].contiguous().pin_memory()
    ltype_host = types[0].long().cpu()[mask].contiguous().pin_memory()
    Xd, yd = X.to(dev), y.to(dev)[mask.to(dev)]
    torch.manual_seed(0)
    cls = pb.PathNet_homo if w["variant"] == "homo" else pb.PathNet
    model = cls(w["F"], H, w["C"], L_, dropout=w["dropout"], math_mode=args.math).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=0.005, weight_decay=0.0005, fused=True)   # same update as PathNet_run.py:295, one kernel
    reducer = FlatGradAllReduce(model.parameters()) if world > 1 else None
    lossf = torch.nn.CrossEntropyLoss(reduction="sum")
    S_global = S * world

    # device-resident inputs for `value`
    neis_dev, ltype_dev = neis_host.to(dev), ltype_host.to(dev)

    def step(neis, ltype, read_loss):
        out = model(Xd, neis, nw, L_, mask, ltype, None)
        loss = lossf(out, yd) / S_global
        opt.zero_grad()   # as PathNet_run.py:350 (grads set to None: autograd assigns instead of accumulating)
        loss.backward()
        if reducer is not None:
            reducer()
        opt.step()
        return loss.item() if read_loss else loss

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

    for _ in range(args.warmup):
        step(neis_dev, ltype_dev, False)
    launches0 = _lib.lib().pn_launch_count()
    with ClockSampler(local) as clocks:
        total_ms = timed(lambda: step(neis_dev, ltype_dev, False), args.steps)
    launches = _lib.lib().pn_launch_count() - launches0
    ms_per_step = total_ms / args.steps
    value = world * P / (ms_per_step * 1e-3)

    # ---- e2e: host index tensors in (pinned), loss out, every step ----
    for _ in range(2):
        step(neis_host, ltype_host, True)
    e2e_ms = timed(lambda: step(neis_host, ltype_host, True), args.steps) / args.steps
    h2d = neis_host.numel() * 8 + ltype_host.numel() * 8
    e2e = {"value": world * P / (e2e_ms * 1e-3), "unit": "paths/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms,
           "note": "module.forward with CPU int64 neis/layer_type (as PathNet_run.py passes them) + loss.item(); "
                   "X stays resident as in the reference (:300)"}

    # ---- roofline of the dominant kernels (in-stream CUDA events inside the library) ----
    _lib.profile_enable(True)
    nprof = 3
    for _ in range(nprof):
        step(neis_dev, ltype_dev, False)
    torch.cuda.synchronize()
    prof = {k: _lib.profile_read(k) for k in ("lstm_fwd", "lstm_bwd", "wgrad")}
    _lib.profile_enable(False)
    flop_fwd = 16.0 * L_ * H * H * P            # 16 L H^2 per path forward
    flop_bwd_dgrad = 16.0 * L_ * H * H * P      # da . Wp^T
    flop_wgrad = 16.0 * L_ * H * H * P          # [x;h]^T da
    kern = {}
    for name, fl in (("lstm_fwd", flop_fwd), ("lstm_bwd", flop_bwd_dgrad), ("wgrad", flop_wgrad)):
        ms, n = prof[name]
        if n:
            kern[name] = {"ms": ms / n, "tflops": fl / (ms / n * 1e-3) / 1e12}
            tr = ncu_traffic(name) if args.workload == "pubmed" and args.math == "tc" else None
            if tr:   # second roof: the kernel's DRAM traffic (committed ncu capture of this workload) over the live duration
                kern[name]["dram_gbs"] = tr / (ms / n * 1e-3) / 1e9
                kern[name]["hbm_frac"] = kern[name]["dram_gbs"] / peaks["hbm"]
            kern[name]["tensor_frac"] = kern[name]["tflops"] / (peaks["bf16_sustained"] or peaks["bf16"])
    dom = max(kern, key=lambda k: kern[k]["ms"]) if kern else None
    peak = peaks["bf16_sustained"] or peaks["bf16"]
    roofline = None
    if dom:
        roofline = {"kernel": dom, "bound": "tensor", "achieved": kern[dom]["tflops"], "peak": peak,
                    "unit": "TFLOP/s", "frac": kern[dom]["tflops"] / peak, "traffic": ncu_traffic(dom),
                    "traffic_note": "DRAM read+write bytes per launch from the committed ncu --set full capture (profiles/r01)",
                    "peak_source": f"{peaks['src']} bf16 sustained (MEASURED_PEAKS.json)",
                    "math_mode": args.math, "kernels": kern,
                    "note": "algorithmic FLOPs = 16*L*H^2 per path per LSTM GEMM pass (SURVEY 8d): forward gates, "
                            "backward dgrad, backward wgrad; math_mode tc = tcgen05 fp16 operands / fp32 accumulate, "
                            "fp32 = CUDA-core FMA (strict parity mode).  kernels[*].hbm_frac = the same kernels against "
                            "the HBM roof (saved-activation / operand-image traffic): the LSTM kernels sit between the two roofs"}

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
    bytes_per_walk = (L_ - 1) * (8 + 16) + L_ * (4 + 1)   # 92 B at L = 4
    merw = {"metric": "merw_walks_per_sec", "value": world * walks / (merw_step * 1e-3), "unit": "walks/s",
            "ms_per_step": merw_step, "walks_per_step": walks, "epochs_per_step": E,
            "labels_included": True, "rng": "glibc_compat (bit-exact srand/rand stream)", "l2_flush": True,
            "roofline": {"kernel": "k_merw_walk", "bound": "hbm",
                         "achieved": walks * bytes_per_walk / ((kms / max(kn, 1)) * 1e-3) / 1e9 if kn else None,
                         "peak": peaks["hbm"], "unit": "GB/s",
                         "frac": (walks * bytes_per_walk / ((kms / kn) * 1e-3) / 1e9 / peaks["hbm"]) if kn else None,
                         "traffic": ncu_traffic("merw_walk"), "bytes_per_walk": bytes_per_walk,
                         "peak_source": f"{peaks['src']} HBM copy bandwidth"}}

    line = {
        "metric": "pagg_paths_per_sec_fwd_bwd", "value": value, "unit": "paths/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.math == "fp32" else "f16xf32acc",
        "data": "synthetic",
        "config": {"workload": f"{w['name']} N={w['n']} F={w['F']} H={w['H']} C={w['C']} path_num={nw} path_len={L_} "
                               f"S={S} rows/GPU -> {P} paths/step/GPU ({'PathNet_homo' if w['variant'] == 'homo' else 'PathNet'}, train mode, dropout {w['dropout']}, "
                               f"CE + Adam), paths drawn by the on-GPU MERW sampler",
                   "parallelism": f"dp{world} (rows sharded, flat NCCL grad all-reduce)" if world > 1 else "single GPU",
                   "l2": "per-step working set ~9 GB >> 126 MB L2 (no explicit flush for PAGG); MERW leg flushes L2 "
                         "with a 256 MiB write between launches",
                   "math_mode": args.math},
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "merw": merw,
    }
    if world > 1:
        dist.barrier()
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            base, _, _ = cpu_pagg_baseline(sample_rows=512, repeats=2)
            line["cpu_baseline"] = base
            line["merw"]["cpu_baseline"] = cpu_merw_baseline(graph)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
