"""A true C caller of the C ABI: tests/c/c_abi_replay.c is compiled with gcc against include/ising_b200.h and linked to
libising_b200.so — no Python between the caller and the library.  On the CPU box it must compile, link and fail loudly
without a device; on the GPU box it replays two golden trajectories (tests/golden/) and must reproduce them."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

SRC = os.path.join(ROOT, "tests", "c", "c_abi_replay.c")
LIBDIR = os.path.join(ROOT, "isingmodel.jl_b200")


def build_exe(tmp_path):
    exe = str(tmp_path / "c_abi_replay")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                           "-L", LIBDIR, "-lising_b200", "-lm", "-Wl,-rpath," + LIBDIR])
    return exe


def write_blob(path):
    g = load_golden("ssf_sk64_glauber")
    N = g["J"].shape[0]
    nsteps = int(g["nsteps"])
    with open(path, "wb") as f:
        np.array([N, 1, nsteps, int(g["rule"]), g["T"].size, int(g["steps_per_T"]), int(g["trace_every"]), 1], dtype=np.int64).tofile(f)
        np.asfortranarray(g["J"]).T.copy().tofile(f)          # column-major J
        g["h"].astype(np.float64).tofile(f)
        g["s0"].astype(np.int8).tofile(f)
        g["nodes"].astype(np.int32).tofile(f)
        g["fluct"].astype(np.float64).tofile(f)
        g["T"].astype(np.float64).tofile(f)
        g["s_final"].astype(np.int8).tofile(f)
        np.array([int(g["flips"])], dtype=np.int64).tofile(f)
        g["E"].astype(np.float64).tofile(f)
        b = load_golden("bip_24x17_ma")
        nv, nh = b["W"].shape
        np.array([nv, nh, 1, int(b["nsteps"]), int(b["rule"]), b["T"].size], dtype=np.int64).tofile(f)
        np.ascontiguousarray(b["W"].T).tofile(f)              # column-major W (nv x nh)
        for k, dt in (("h", np.float64), ("b", np.float64), ("s0", np.int8), ("t0", np.int8), ("Fv", np.float64),
                      ("Fh", np.float64), ("T", np.float64), ("s_final", np.int8), ("t_final", np.int8)):
            np.ascontiguousarray(b[k], dtype=dt).tofile(f)


def test_c_caller_compiles_links_and_fails_loudly_without_a_device(tmp_path):
    exe = build_exe(tmp_path)
    out = subprocess.run([exe, "--symbols"], capture_output=True, text=True)
    assert out.returncode == 0 and "22 entry points linked" in out.stdout, out.stdout + out.stderr
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the replay itself runs in the gpu test")
    blob = str(tmp_path / "blob.bin")
    write_blob(blob)
    out = subprocess.run([exe, blob], capture_output=True, text=True)
    assert out.returncode == 3 and "no CPU fallback" in out.stderr, out.stdout + out.stderr


@pytest.mark.gpu
def test_c_caller_replays_the_goldens(tmp_path):
    exe = build_exe(tmp_path)
    blob = str(tmp_path / "blob.bin")
    write_blob(blob)
    out = subprocess.run([exe, blob], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
    assert out.stdout.count("identical") == 2
