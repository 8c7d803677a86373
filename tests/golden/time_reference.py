"""Times the UNMODIFIED reference on the CPU for BASELINE config 2 (build container only; needs /root/reference).
Compare with `python bench.py --impl reference`: on the 8-vCPU build container the reference runs 14.1 ms/step
(1.16 M env-steps/s) and the oracle port 13.0 ms/step (1.26 M env-steps/s), i.e. the port used as `cpu_baseline`
is within 8 % of the real thing."""
import sys, time, os, torch
HERE=os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0,os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0,HERE)
from leibnizgym_b200.config import difficulty_config
from leibnizgym_b200.synthetic import make_sequence
from ref_harness import build_reference_env
N=16384; T=4
cfg=difficulty_config(2,N,asymmetric_obs=True,seed=1002)
seq=make_sequence(1002,T,N)
env,fake=build_reference_env(cfg,seq)
orig=fake.simulate
sim_t=[0.0]
def timed(sim):
    t0=time.perf_counter(); fake.cursor%=T; orig(sim); sim_t[0]+=time.perf_counter()-t0
fake.simulate=timed
env.reset()
for t in range(3): env.step(seq.action[t%T].clone())
sim_t[0]=0.0; n=40
t0=time.perf_counter()
for t in range(n): env.step(seq.action[t%T].clone())
wall=time.perf_counter()-t0
hot=wall-sim_t[0]
print("reference:", torch.get_num_threads(), "threads", round(1e3*hot/n,2), "ms/step", round(n*N/hot/1e6,3), "M env-steps/s")
