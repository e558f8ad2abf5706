"""Cycle accounting of the tensor kernel's sampling warps on the C4 workload (measurement build: scripts/build_variant.sh
timing -DISB_TC_TIMING; run with ISING_B200_LIB=scratch_ab/lib_timing.so).  Prints, per sampling warp and averaged over the
CTAs, the share of the kernel spent waiting for an accumulator, in the chunk loop, and in the per-tile prologue / epilogue."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isingmodel_jl_b200 import synth, SpinSystems, OnBipartiteGraph, _lib

prec_name = sys.argv[1] if len(sys.argv) > 1 else "i8x3"
nst = int(sys.argv[2]) if len(sys.argv) > 2 else 200
L = _lib
W, h, b = synth.bipartite_W(784, 512, 4, 0.1)
R = 16384
prec = {"i8x3": L.PREC_I8X3, "i8x2": L.PREC_I8X2, "bf16x1": L.PREC_BF16X1}[prec_name]
sv, sh = synth.spins(11, R, 784), synth.spins(12, R, 512)
ss = SpinSystems.SpinSystemOnBipartiteGraph(sv, sh, W, h, b, device=0, prec=prec)
ua = OnBipartiteGraph.StochasticCellularAutomata(ss, 1.0)
ens = ss._ensemble()
T = np.ones(nst)
for k in range(3):
    ens.bip_run(L.BIP_SCA, nst, seed=777, step_offset=k * nst, T=T)
torch.cuda.synchronize()
lib = ctypes.CDLL(L.LIB_PATH)
n = 296 * 16 * 4
buf = (ctypes.c_longlong * n)()
rc = lib.isb_debug_tc_timing(buf, n)
a = np.array(buf[:], dtype=np.int64).reshape(296, 16, 4)[:148]
st = ens.last_stats()
print(f"{prec_name}: {nst} SCA steps, kernel {st['kernel_ms']:.3f} ms = {1e3 * st['kernel_ms'] / (2 * nst):.2f} us per half-step (rc {rc})")
tot = a[:, 4:, 0].astype(float)
print("sampling warps, mean over CTAs and warps: total %.0f cycles = %.0f per half-step" % (tot.mean(), tot.mean() / (2 * nst)))
for name, i in (("wait for accumulator", 1), ("chunk loop", 2), ("tile epilogue (arrivals, fences)", 3)):
    print("  %-34s %5.1f %%" % (name, 100 * (a[:, 4:, i] / np.maximum(tot, 1)).mean()))
rest = 1 - (a[:, 4:, 1:4].sum(axis=2) / np.maximum(tot, 1))
print("  %-34s %5.1f %%" % ("tile prologue + job iteration", 100 * rest.mean()))
print("per quadrant (warp %% 4) chunk-loop share:", [round(float(100 * (a[:, 4 + q::4, 2] / np.maximum(a[:, 4 + q::4, 0], 1)).mean()), 1) for q in range(4)])
