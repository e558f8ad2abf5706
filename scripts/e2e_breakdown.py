"""Where the end-to-end time of a C3 / C4 bench step goes (host buffers in, sampled layers + energies out)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from isingmodel_jl_b200 import synth, SpinSystems, OnBipartiteGraph, SamplingHelper, _lib as L

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
nst = int(sys.argv[2]) if len(sys.argv) > 2 else bench.SCA_STEPS[which]
W, h, b, R, sched, desc = bench.sca_workload(which)
nv, nh = W.shape
T = sched(nst)
pv = torch.empty((R, nv), dtype=torch.int8).pin_memory().numpy()
ph = torch.empty((R, nh), dtype=torch.int8).pin_memory().numpy()
pv[:] = synth.spins(11, R, nv); ph[:] = synth.spins(12, R, nh)
ss = SpinSystems.SpinSystemOnBipartiteGraph(pv, ph, W, h, b, device=0, prec=L.PREC_I8X3)
ua = OnBipartiteGraph.StochasticCellularAutomata(ss, float(T[0]))
ens = ss._ensemble()
def sync(): torch.cuda.synchronize()
for rep in range(3):
    t = [time.perf_counter()]
    def lap(): sync(); t.append(time.perf_counter())
    ss.spinConfiguration = pv; lap()
    ss.hiddenLayer = ph; lap()
    SamplingHelper.run_(ua, nst, seed=777, step_offset=rep * nst, temperatures=T); lap()
    st = ens.last_stats(); lap()
    S = ss.spinConfiguration; lap()
    Tm = ss.hiddenLayer; lap()
    E = SpinSystems.calcEnergy(ua); lap()
    names = ["set spins", "set hidden", "run_", "last_stats", "get spins", "get hidden", "calcEnergy"]
    d = np.diff(t) * 1e3
    print(which, "rep", rep, " ".join(f"{n}={x:.1f}ms" for n, x in zip(names, d)), f"total={d.sum():.1f}ms kernel={st['kernel_ms']:.1f}ms")
