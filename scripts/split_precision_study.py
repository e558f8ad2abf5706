"""CPU study (numpy): how closely do k-term low-precision splits of the couplings represent W, and what does that do to
the block-Gibbs decision quantity 2 (W' s + b)?  bf16 terms are what the tcgen05 path ships (ISB_PREC_BF16X1/2/3);
fp16 terms (11 significant bits each, W pre-scaled by a power of two so that the residual terms stay normal) are the
candidate for a 2-pass mode with near-fp32 couplings.  Prints, for the C4 and C3 coupling distributions, the largest
relative representation error and the largest error of the decision quantity over random +-1 inputs."""
import numpy as np


def bf16_round(x):
    u = x.astype(np.float32).view(np.uint32)
    r = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return r.view(np.float32).astype(np.float64)


def fp16_round(x):
    return x.astype(np.float16).astype(np.float64)


def split(W, rnd, terms, scale=1.0):
    res = W * scale
    parts = []
    for _ in range(terms):
        t = rnd(res)
        parts.append(t)
        res = res - t
    return sum(parts) / scale


def study(name, W, nsamp=64, seed=0):
    rng = np.random.default_rng(seed)
    S = rng.choice([-1.0, 1.0], size=(nsamp, W.shape[0]))
    exact = 2.0 * S @ W
    amax = np.abs(W).max()
    # power-of-two pre-scale that puts max|W| near 2^14: the third-order residual of fp16 stays far above 2^-14
    sc = 2.0 ** (14 - np.ceil(np.log2(amax)))
    rows = []
    for label, rnd, k, scale in (("bf16x1", bf16_round, 1, 1.0), ("bf16x2", bf16_round, 2, 1.0),
                                 ("bf16x3", bf16_round, 3, 1.0), ("fp16x1", fp16_round, 1, sc),
                                 ("fp16x2", fp16_round, 2, sc), ("fp32", lambda x: x.astype(np.float32).astype(np.float64), 1, 1.0)):
        Wq = split(W, rnd, k, scale)
        rel = np.abs(Wq - W).max() / amax
        dec = np.abs(2.0 * S @ Wq - exact).max()
        rows.append((label, rel, dec, dec / np.abs(W).sum(0).max()))
    print(f"{name}: W {W.shape}, max|W| = {amax:.4g}, max column sum of |W| = {np.abs(W).sum(0).max():.4g}")
    for label, rel, dec, decrel in rows:
        print(f"  {label:7s} max|W~ - W| / max|W| = {rel:9.3e} (2^{np.log2(rel):6.1f})   max decision error = {dec:9.3e}"
              f"  (= {decrel:8.2e} x sum|W|)")


if __name__ == "__main__":
    rng = np.random.default_rng(4)
    study("C4 (784 x 512, N(0, 0.01))", rng.normal(0.0, 0.1, (784, 512)))
    n = 4096
    J = rng.normal(0.0, 1.0 / np.sqrt(n), (n, n))
    J = np.triu(J, 1)
    J = J + J.T
    study("C3 (W = (J + qI)/2, N = 4096)", 0.5 * (J + 1.0 * np.eye(n)), nsamp=16)
