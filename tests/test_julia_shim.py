"""CPU-only: the Julia shim (isingmodel.jl_b200/julia/IsingModelB200.jl) cannot be executed here (no julia in the
image), so its `ccall` sites are checked STATICALLY against the prototypes of include/ising_b200.h: every called
symbol is declared, the argument counts agree, and each Julia argument type is ABI-compatible with the C type in
the same position (Cint <-> int, Int64 <-> int64_t, Ptr{Float64} <-> double *, handles <-> Ptr{Cvoid}, ...).
It also checks that the shim keeps the reference's public names (src/IsingModel.jl:3-15 and the exports of its
five sub-modules)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ising_b200.h")
SHIM = os.path.join(ROOT, "isingmodel.jl_b200", "julia", "IsingModelB200.jl")


def c_prototypes():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b((?:const\s+)?[A-Za-z_0-9]+\s*\**)\s*\b(isb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", txt):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        args = [a.strip() for a in args.replace("\n", " ").split(",")]
        if args == ["void"] or args == [""]:
            args = []
        protos[name] = (" ".join(ret.split()), [c_kind(a) for a in args])
    return protos


def c_kind(decl):
    """ABI class of one C parameter declaration."""
    d = re.sub(r"\bconst\b", "", decl).strip()
    if "*" in d or "[" in d:
        base = d.split("*")[0].split("[")[0].split()
        base = base[0] if base else ""
        if d.count("*") == 2:
            return "ptr:handle_out"
        if base.startswith("isb_"):
            return "ptr:handle"
        return "ptr:" + {"double": "f64", "float": "f32", "int8_t": "i8", "int32_t": "i32", "int64_t": "i64",
                         "uint32_t": "u32", "uint64_t": "u64", "int": "i32", "char": "char", "void": "void"}[base]
    base = d.split()[0]
    return {"int": "i32", "int64_t": "i64", "uint64_t": "u64", "uint32_t": "u32", "double": "f64",
            "float": "f32", "size_t": "u64"}[base]


JL = {
    "Cint": {"i32"}, "Int32": {"i32"}, "Int64": {"i64"}, "UInt64": {"u64"}, "UInt32": {"u32"},
    "Cdouble": {"f64"}, "Float64": {"f64"}, "Csize_t": {"u64"},
    "Ctx": {"ptr:handle"}, "Model": {"ptr:handle"}, "Ens": {"ptr:handle"}, "ShardRun": {"ptr:handle"},
    "Ref{Ctx}": {"ptr:handle_out"}, "Ref{Model}": {"ptr:handle_out"}, "Ref{Ens}": {"ptr:handle_out"},
    "Ref{ShardRun}": {"ptr:handle_out"}, "Ptr{UInt8}": {"ptr:void"},
    "Ptr{Float64}": {"ptr:f64"}, "Ptr{Int8}": {"ptr:i8"}, "Ptr{Int32}": {"ptr:i32"}, "Ptr{Int64}": {"ptr:i64"},
    "Ptr{UInt32}": {"ptr:u32"}, "Ptr{UInt64}": {"ptr:u64"}, "Ref{Cint}": {"ptr:i32"}, "Ref{Int64}": {"ptr:i64"},
    "Ptr{Cvoid}": {"ptr:void", "ptr:handle"}, "Cstring": {"ptr:char"},
}


def split_top(s):
    """Split a Julia tuple body at top-level commas (braces may nest)."""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls():
    txt = open(SHIM).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(isb_[a-z0-9_]+),\s*libisb\),\s*([A-Za-z0-9_{}]+),\s*\(", txt):
        name, ret = m.group(1), m.group(2)
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(txt[i], 0)
            i += 1
        types = split_top(txt[m.end():i - 1])
        # the call's value arguments run up to the parenthesis that closes ccall(
        j, depth = i, 1
        while depth:
            depth += {"(": 1, ")": -1}.get(txt[j], 0)
            j += 1
        values = split_top(txt[i:j - 1].lstrip(", \n"))
        calls.append((name, ret, types, values, txt.count("\n", 0, m.start()) + 1))
    return calls


def test_every_ccall_matches_the_header():
    protos = c_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 15
    for name, ret, types, values, line in calls:
        where = f"IsingModelB200.jl:{line} ccall {name}"
        assert name in protos, f"{where}: not declared in include/ising_b200.h"
        c_ret, c_args = protos[name]
        assert len(types) == len(c_args), f"{where}: {len(types)} argument types, the header has {len(c_args)}"
        assert len(values) == len(types), f"{where}: {len(values)} values for {len(types)} argument types"
        for pos, (jt, ck) in enumerate(zip(types, c_args)):
            assert jt in JL, f"{where}: unknown Julia type {jt}"
            assert ck in JL[jt], f"{where}: argument {pos + 1} is {jt} in the shim but {ck} in the header"
        if c_ret == "int":
            assert ret == "Cint", where
        elif "char" in c_ret:
            assert ret == "Cstring", where
        elif c_ret == "void":
            assert ret == "Cvoid", where


def test_run_entry_points_are_bound():
    """The path's entry points proper: model / ensemble construction, spins in and out, both step
    loops, energy and local fields."""
    bound = {c[0] for c in julia_ccalls()}
    for need in ("isb_create", "isb_model_dense", "isb_model_sparse", "isb_model_bipartite", "isb_ens_create",
                 "isb_ens_set_spins", "isb_ens_get_spins", "isb_ens_set_hidden", "isb_ens_get_hidden",
                 "isb_ens_energy", "isb_ens_local_field", "isb_ens_local_aux_bias", "isb_ssf_run", "isb_bip_run",
                 "isb_last_error"):
        assert need in bound, need


def test_shim_keeps_the_reference_names():
    """Same module, type and function names as the reference (src/IsingModel.jl:3-15; exports at
    src/SpinSystems.jl:3-8, src/SingleSpinFlip.jl:3-4, src/MultiSpinFlip.jl:3, src/OnBipartiteGraph.jl:3-4,
    src/SamplingHelper.jl:3)."""
    txt = open(SHIM).read()
    for mod in ("SpinSystems", "SingleSpinFlip", "MultiSpinFlip", "OnBipartiteGraph", "SamplingHelper"):
        assert re.search(rf"^module {mod}\b", txt, flags=re.M), mod
    for name in ("SpinSystem", "SpinSystemOnBipartiteGraph", "UpdatingAlgorithm", "UpdatingAlgorithmOnBipartiteGraph",
                 "getSpinConfiguration", "getCouplingCoefficients", "getExternalMagneticField", "getHiddenLayer",
                 "getAuxiliaryBias", "calcEnergy", "calcLocalMagneticField", "calcLocalAuxiliaryBias", "heaviside",
                 "AsynchronousHopfieldNetwork", "GlauberDynamics", "MetropolisMethod", "StochasticCellularAutomata",
                 "MomentumAnnealing", "update!", "makeSampler!"):
        assert name in txt, name


def test_ctypes_signatures_match_the_header(pkg):
    """The Python mirror's ctypes table (isingmodel.jl_b200/_lib.py: SIGNATURES) against the same prototypes:
    argument counts, scalar widths / signedness and pointer-vs-value class of every parameter and return value."""
    import ctypes as C

    from isingmodel_jl_b200 import _lib
    protos = c_prototypes()
    scalar = {C.c_int: "i32", C.c_int64: "i64", C.c_uint64: "u64", C.c_double: "f64"}
    assert sorted(_lib.SIGNATURES) == sorted(protos)
    for name, (res, args) in _lib.SIGNATURES.items():
        c_ret, c_args = protos[name]
        assert len(args) == len(c_args), f"{name}: {len(args)} ctypes arguments, the header has {len(c_args)}"
        for pos, (a, ck) in enumerate(zip(args, c_args)):
            if a in scalar:
                assert scalar[a] == ck, f"{name}: argument {pos + 1} is {scalar[a]} in _lib.py but {ck} in the header"
            else:  # c_void_p or POINTER(...)
                assert ck.startswith("ptr:"), f"{name}: argument {pos + 1} is a pointer in _lib.py but {ck} in the header"
        want = {"int": C.c_int, "int64_t": C.c_int64, "void": None, "const char *": C.c_char_p}[c_ret]
        assert res is want, f"{name}: return type"


def test_enum_values_agree_across_the_layers():
    """The enums of include/ising_b200.h against the constants of the Python mirror and of the Julia shim."""
    from isingmodel_jl_b200 import _lib
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    enums = {k: int(v) for k, v in re.findall(r"\b(ISB_[A-Z0-9_]+)\s*=\s*(\d+)", txt)}
    assert len(enums) >= 25
    for name, val in enums.items():
        py = name[len("ISB_"):]
        if hasattr(_lib, py):
            assert getattr(_lib, py) == val, name
    for group in ("PREC_", "RULE_", "BIP_", "ORDER_", "FLUCT_"):
        names = [n for n in enums if n.startswith("ISB_" + group)]
        assert names, group
        for n in names:
            assert hasattr(_lib, n[4:]), f"{n} has no constant in _lib.py"
    jl = open(SHIM).read()
    for m in re.finditer(r"const ([A-Z0-9_, ]+?) = ((?:Cint\(\d+\)(?:, )?)+)", jl):
        names = [n.strip() for n in m.group(1).split(",")]
        vals = [int(v) for v in re.findall(r"Cint\((\d+)\)", m.group(2))]
        assert len(names) == len(vals)
        for n, v in zip(names, vals):
            assert enums.get("ISB_" + n) == v, f"IsingModelB200.jl: {n} = {v}, the header says {enums.get('ISB_' + n)}"


def _strip_julia(txt):
    """Source without comments and string contents (enough for the structural checks below)."""
    out = []
    for line in txt.split("\n"):
        s, q, i = "", False, 0
        while i < len(line):
            ch = line[i]
            if q:
                if ch == "\\":
                    i += 2
                    continue
                if ch == '"':
                    q = False
                    s += '"'
            elif ch == '"':
                q = True
                s += '"'
            elif ch == "#":
                break
            else:
                s += ch
            i += 1
        out.append(s.rstrip())
    return "\n".join(out)


def test_shim_blocks_and_brackets_balance():
    """No julia binary here: at least every block opener has its `end` and every bracket closes."""
    src = _strip_julia(open(SHIM).read())
    depth = {"(": 0, "[": 0, "{": 0}
    pair = {")": "(", "]": "[", "}": "{"}
    for ch in src:
        if ch in depth:
            depth[ch] += 1
        elif ch in pair:
            depth[pair[ch]] -= 1
            assert depth[pair[ch]] >= 0
    assert depth == {"(": 0, "[": 0, "{": 0}, depth
    opens = ends = 0
    stack = []
    for n, line in enumerate(src.split("\n"), 1):
        t = line.strip()
        if not t or t.startswith("abstract type") or t.startswith("@enum"):
            continue
        if re.match(r"(module|function|mutable struct|struct|if|for|while|try|let|begin)\b", t) and not re.search(r"\bend$", t):
            opens += 1
            stack.append((n, t[:40]))
        elif re.search(r"\bdo( \w+)?$", t):
            opens += 1
            stack.append((n, t[:40]))
        if re.match(r"end\b", t):
            ends += 1
            assert stack, f"line {n}: `end` without an opener"
            stack.pop()
    assert opens == ends and not stack, (opens, ends, stack[-3:])


def test_shim_keeps_reference_object_semantics():
    """What the reference's own test needs from the host objects (test/runtests.jl:6-31, src/SpinSystems.jl:61-66,
    128-137, src/SamplingHelper.jl:64-91): independent deep copies, released handles, setters and assignments that reach
    the device, streaming samplers on the snapshot entry points, and the MultiSpinFlip sampling methods."""
    txt = open(SHIM).read()
    bound = {c[0] for c in julia_ccalls()}
    for need in ("isb_model_retain", "isb_model_destroy", "isb_ens_clone", "isb_ens_destroy", "isb_ssf_run_snap",
                 "isb_bip_run_snap", "isb_shard_model_sk", "isb_shard_model_rows_q", "isb_shard_run_create",
                 "isb_shard_run_destroy", "isb_nccl_unique_id", "isb_shard_run_init_nccl", "isb_shard_run_set_nccl_comm",
                 "isb_shard_run_ipc_export", "isb_shard_run_ipc_import", "isb_shard_run_barrier", "isb_shard_run_set_spins",
                 "isb_shard_run_get_spins", "isb_shard_run_steps"):
        assert need in bound, need
    assert len(re.findall(r"Base\.deepcopy_internal\((?:ss)::(SpinSystem|SpinSystemOnBipartiteGraph), dict::IdDict\)", txt)) == 2
    assert txt.count("finalizer(_release!, ss)") == 4            # both constructors of both system types
    assert len(re.findall(r"function Base\.setproperty!\(ss::(SpinSystem|SpinSystemOnBipartiteGraph), name::Symbol, v\)", txt)) == 2
    for setter in ("setSpinConfiguration(ua::UpdatingAlgorithm,", "setCouplingCoefficients(ua::UpdatingAlgorithm,",
                   "setExternalMagneticField(ua::UpdatingAlgorithm,", "setSpinConfiguration(ua::UpdatingAlgorithmOnBipartiteGraph,",
                   "setHiddenLayer(ua::UpdatingAlgorithmOnBipartiteGraph,", "setCouplingCoefficients(ua::UpdatingAlgorithmOnBipartiteGraph,",
                   "setExternalMagneticField(ua::UpdatingAlgorithmOnBipartiteGraph,", "setAuxiliaryBias(ua::UpdatingAlgorithmOnBipartiteGraph,"):
        assert setter in txt, setter
    # every device operation of the host objects is preceded by the host -> device check
    for fn in ("calcEnergy(ss::SpinSystem)", "calcEnergy(ss::SpinSystemOnBipartiteGraph)"):
        line = next(ln for ln in txt.split("\n") if ln.startswith(fn))
        assert "_sync!(ss)" in line, fn
    # the three samplers of src/SamplingHelper.jl (keyword schedule, POSITIONAL schedule for MultiSpinFlip, keyword)
    assert re.search(r"function makeSampler!\(updatingAlgorithm::SingleSpinFlip\.SingleSpinUpdatingAlgorithm, maxMCSteps::Integer;", txt)
    assert re.search(r"function makeSampler!\(updatingAlgorithm::MultiSpinFlip\.MultiSpinUpdatingAlgorithm, maxMCSteps::Integer,\s*annealingSchedule::Function", txt)
    assert re.search(r"function makeSampler!\(updatingAlgorithm::UpdatingAlgorithmOnBipartiteGraph, maxMCSteps::Integer;", txt)
    assert re.search(r"function update!\(ua::MultiSpinFlip\.MultiSpinUpdatingAlgorithm; rng", txt)
    assert txt.count("trace_every = 1") >= 3 and txt.count("put!(channel, ua)") == 6     # n + 1 items from each sampler
