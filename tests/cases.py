"""Inputs of the golden cases whose big matrices are regenerated from seeds (see tests/golden/make_golden.py)."""
import numpy as np


def golden_J(name, g, synth):
    if "J" in g:
        return g["J"]
    if name == "ssf_c1_32x32_metropolis":
        J = synth.lattice_J(32)
    elif name == "ssf_c2_sk1024_glauber":
        J = synth.sk_J(1024, 2)
    else:
        raise KeyError(name)
    from conftest import sha
    assert sha(J) == str(g["J_sha"]), "synthetic J drifted from the one the golden file was made with"
    return J


def golden_W(name, g, synth):
    if "W" in g:
        return g["W"]
    if name.startswith("bip_c4_784x512"):
        W, _, _ = synth.bipartite_W(784, 512, 4)
    else:
        raise KeyError(name)
    from conftest import sha
    assert sha(W) == str(g["W_sha"]), "synthetic W drifted from the one the golden file was made with"
    return W


SSF_GOLDEN = ["ssf_2spin_hopfield", "ssf_2spin_glauber", "ssf_2spin_metropolis", "ssf_3x3_glauber",
              "ssf_3x3_metropolis", "ssf_c1_32x32_metropolis", "ssf_sk64_glauber", "ssf_sk64_hopfield",
              "ssf_c2_sk1024_glauber"]
BIP_GOLDEN = ["bip_2x3_sca", "bip_2x3_ma", "bip_24x17_sca", "bip_24x17_ma", "bip_c4_784x512_sca",
              "bip_c4_784x512_ma"]


def nodes_of(g):
    return None if g["nodes"].size == 0 else g["nodes"].astype(np.int32)
