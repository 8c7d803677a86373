#!/usr/bin/env python
"""Development timing: the observation / state fill alone (lg_fill_observations: load -> scale -> store + history
shift, no reward chain) back to back over the ring, against the full post-physics kernel — how much of the kernel's
time is the streaming part.  Needs a GPU."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, workload_config  # noqa: E402
from leibnizgym_b200 import _native as nat  # noqa: E402
from leibnizgym_b200.env import TrifingerEnv  # noqa: E402
from leibnizgym_b200.graph_runner import GraphRunner  # noqa: E402
from leibnizgym_b200.sim import SyntheticSim  # noqa: E402
from leibnizgym_b200.synthetic import make_sequence  # noqa: E402


def main():
    wl = dict(WORKLOADS["c2"])
    N, R, dev = wl["envs"], 32, "cuda:0"
    ring = make_sequence(wl["seed"], R, N, device=dev)
    env = TrifingerEnv(workload_config(wl, N), device=dev, verbose=False, sim=SyntheticSim(ring, dev))
    env.reset()
    runner = GraphRunner(env, ring, rotate_outputs=True)
    lib = env._lib
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for mode in ("fill", "post"):
            def launch(t, stream):
                s = t % R
                if mode == "fill":
                    nat.check(lib.lg_fill_observations(runner.P, runner._S[s], runner._B[s], stream), "fill")
                else:
                    nat.check(lib.lg_post_physics(runner.P, runner._S[s], runner._B[s], 0.0, stream), "post")
            for t in range(2):
                launch(t, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for t in range(128):
                    launch(t, torch.cuda.current_stream().cuda_stream)
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(40):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            print(f"{mode}: {1e3 * e0.elapsed_time(e1) / (40 * 128):.2f} us per launch")


if __name__ == "__main__":
    main()
