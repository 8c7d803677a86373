#!/usr/bin/env python
"""Phase timeline of the two step kernels inside a replayed CUDA graph (development tool, needs a GPU).

    python -m leibnizgym_b200.build -DLG_TRACE -o ab/trace.so
    LG_LIB_PATH=ab/trace.so python scripts/trace_step.py [--envs 16384] [--workload c2] [--post-only]

The -DLG_TRACE build stamps %globaltimer at named points of pre_physics_kernel / post_physics_kernel (lane 0 of
selected warps, every CTA).  This script replays a graph of consecutive steps, reads the stamps of the LAST launch of
each kernel and prints, per point, when it was reached relative to the first CTA's entry: min / median / p90 / max over the
CTAs.  The step time from CUDA events next to it tells how much of a step is the gap between kernels.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, workload_config  # noqa: E402
from leibnizgym_b200 import _native as nat  # noqa: E402
from leibnizgym_b200.env import TrifingerEnv  # noqa: E402
from leibnizgym_b200.graph_runner import GraphRunner  # noqa: E402
from leibnizgym_b200.sim import SyntheticSim  # noqa: E402
from leibnizgym_b200.synthetic import bernoulli_masks, make_sequence  # noqa: E402

SLOTS = 32
POST = {0: "entry (w0)", 1: "after dependency wait", 10: "loads issued, before trigger", 11: "windows stored (w0)",
        22: "windows barrier passed", 2: "loads issued, after trigger", 23: "stats: ballots done", 24: "stats: lane sums done",
        25: "stats: scaled (before atomic)", 3: "staged (data arrived, w0)", 4: "after barrier",
        5: "sub-task done (w0)", 16: "sub-task done (w1)", 17: "sub-task done (w2)", 18: "sub-task done (w3)",
        6: "after reward barrier", 7: "combined (w0)", 8: "statistics issued", 9: "emit done (w0)",
        19: "emit done (w1)", 20: "emit done (w2)", 21: "emit done (w3)",
        12: "entry (w4)", 13: "after dependency wait (w4)", 14: "after barrier (w4)", 15: "emit done (w4)"}
PRE = {0: "entry", 1: "slabs issued", 2: "after dependency wait", 3: "aggregate published (scan group)",
       4: "look-back + id lists done (scan group)", 5: "slabs landed", 6: "action row done", 7: "resets done",
       8: "torque done", 9: "bulk stores issued", 11: "exit"}


def read_trace(lib, which, ctas):
    buf = np.zeros((ctas, 2 * SLOTS), dtype=np.uint64)
    lib.lg_trace_read.restype = C.c_int
    lib.lg_trace_read.argtypes = [C.c_int, C.c_void_p, C.c_int]
    nat.check(lib.lg_trace_read(which, buf.ctypes.data, ctas), "lg_trace_read")
    return buf[:, :SLOTS].astype(np.int64), buf[:, SLOTS:].astype(np.int64)


def show(name, names, t, t_ref=None):
    t0 = t[:, 0].min() if t_ref is None else t_ref
    print(f"--- {name}: {t.shape[0]} CTAs; ns after the first CTA's entry (min / median / p90 / max over CTAs)")
    for slot in sorted(names, key=lambda s: np.median(t[:, s])):
        col = t[:, slot] - t0
        print(f"  {names[slot]:<32s} {col.min():7d} {int(np.median(col)):7d} {int(np.percentile(col, 90)):7d} {col.max():7d}")
    return t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--envs", type=int, default=None)
    ap.add_argument("--ring", type=int, default=32)
    ap.add_argument("--post-only", action="store_true")
    a = ap.parse_args()
    wl = dict(WORKLOADS[a.workload])
    N = a.envs or wl["envs"]
    dev = "cuda:0"
    lib = nat.load()
    cfg = workload_config(wl, N)
    ring = make_sequence(wl["seed"], a.ring, N, device=dev)
    masks = bernoulli_masks(wl["seed"], a.ring, N, wl["reset_p"], device=dev)
    env = TrifingerEnv(cfg, device=dev, verbose=False, sim=SyntheticSim(ring, dev))
    env.reset()
    runner = GraphRunner(env, ring, rotate_outputs=True, inject_reset_masks=masks)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        runner.capture(a.ring * 2, post_only=a.post_only)
        for _ in range(3):
            runner.graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            runner.graph.replay()
        e1.record()
        torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (20 * a.ring * 2)
    print(f"{N} envs, workload {a.workload}, post_only={a.post_only}: {us:.2f} us per step (CUDA events)")
    post_ctas = min(8192, (N + 27) // 28 if N <= 16576 else (N + 31) // 32)
    pre_ctas = min(8192, (N + 127) // 128)
    tp, cp = read_trace(lib, 0, post_ctas)
    ok = tp[:, 0] > 0
    tp, cp = tp[ok], cp[ok]
    # SM clock actually seen inside the kernel: cycles between two stamps of a CTA over the nanoseconds between them
    dt, dc = (tp[:, 9] - tp[:, 1]).astype(np.float64), (cp[:, 9] - cp[:, 1]).astype(np.float64)
    good = dt > 500
    print(f"SM clock inside the post kernel (dependency wait -> last store, per CTA): median {np.median(dc[good] / dt[good]) * 1e3:.0f} MHz")
    for k in [k for k in POST if (tp[:, k] == 0).all()]:
        POST.pop(k)    # points the traced kernel variant does not stamp
    if not a.post_only:
        tq, _ = read_trace(lib, 1, pre_ctas)
        tq = tq[tq[:, 0] > 0]
        ref = show("pre_physics_kernel (last launch)", PRE, tq)
        show("post_physics_kernel (last launch), relative to the pre kernel's first entry", POST, tp, ref)
        span = tp[:, [9, 15, 19, 20, 21]].max() - tq[:, 0].min()
    else:
        show("post_physics_kernel (last launch)", POST, tp)
        span = tp[:, [9, 15, 19, 20, 21]].max() - tp[:, 0].min()
    print(f"first entry -> last exit stamp: {span} ns; step {1e3 * us:.0f} ns -> {1e3 * us - span:.0f} ns outside the stamps")
    res = np.diff(np.unique(tp.ravel()))
    print(f"globaltimer resolution seen: {res[res > 0].min()} ns")


if __name__ == "__main__":
    main()
