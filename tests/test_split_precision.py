"""CPU-only: the arithmetic behind the coupling-storage modes of the tensor path (include/ising_b200.h:
ISB_PREC_BF16X1/2/3, ISB_PREC_FP16X1/2), restated in numpy (scripts/split_precision_study.py): how closely k
low-precision terms represent W, and that small-integer couplings are exact in a single term."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("split_study", os.path.join(ROOT, "scripts", "split_precision_study.py"))
study = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(study)


def _scale(W):
    return 2.0 ** (14 - np.ceil(np.log2(np.abs(W).max())))


def test_split_terms_represent_gaussian_couplings():
    rng = np.random.default_rng(4)
    W = rng.normal(0.0, 0.1, (784, 512))
    amax = np.abs(W).max()
    err = lambda Wq: np.abs(Wq - W).max() / amax
    assert err(study.split(W, study.bf16_round, 1)) < 2.0 ** -8
    assert err(study.split(W, study.bf16_round, 2)) < 2.0 ** -16
    assert err(study.split(W, study.bf16_round, 3)) < 2.0 ** -24
    sc = _scale(W)
    assert err(study.split(W, study.fp16_round, 1, sc)) < 2.0 ** -11
    # two fp16 terms: 11 + 11 bits and the sign of the second term -> as good as W rounded to float
    assert err(study.split(W, study.fp16_round, 2, sc)) < 2.0 ** -23
    assert err(W.astype(np.float32).astype(np.float64)) < 2.0 ** -24


def test_small_integer_couplings_are_exact_in_one_term():
    rng = np.random.default_rng(5)
    W = np.round(rng.normal(0.0, 1.5, (192, 128)))
    assert np.array_equal(study.split(W, study.bf16_round, 1), W)
    assert np.array_equal(study.split(W, study.fp16_round, 1, _scale(W)), W)
