# IsingModelB200.jl — Julia host side of libising_b200.so (UNTESTED in the build environment: no julia binary
# exists in the image; the same C ABI is exercised from Python ctypes in tests/).  Keeps the reference's module,
# type and method names (src/IsingModel.jl:3-15) so that user code written for Wandao123/IsingModel.jl runs
# unchanged; only the arithmetic moves behind `ccall`.
#
#   using IsingModelB200                       # instead of `using IsingModel`
#   ss = SpinSystems.SpinSystem(s, J, h)       # src/SpinSystems.jl:19-51 (same checks, warnings, errors)
#   ua = SingleSpinFlip.GlauberDynamics(ss, T)
#   SingleSpinFlip.update!(ua, node, fluct)    # src/SingleSpinFlip.jl:46-55, one isb_ssf_run call
#   for ua in SamplingHelper.makeSampler!(ua, n; annealingSchedule, rng) ... end
#
# Extension: `spinConfiguration` may be an N x R matrix (R replicas, one per column).
module IsingModelB200

export SpinSystems, SingleSpinFlip, MultiSpinFlip, OnBipartiteGraph, SamplingHelper

const libisb = get(ENV, "ISING_B200_LIB", joinpath(@__DIR__, "..", "libising_b200.so"))

# ------------------------------------------------------------------ raw C ABI (include/ising_b200.h)
module CABI
import ..libisb
const Ctx = Ptr{Cvoid}; const Model = Ptr{Cvoid}; const Ens = Ptr{Cvoid}
const RULE_HOPFIELD, RULE_GLAUBER, RULE_METROPOLIS = Cint(0), Cint(1), Cint(2)
const BIP_SCA, BIP_MA = Cint(0), Cint(1)
const ORDER_SEQUENTIAL, ORDER_LIST, ORDER_RANDOM = Cint(0), Cint(1), Cint(2)
const FLUCT_PHILOX, FLUCT_SHARED, FLUCT_PER_REPLICA = Cint(0), Cint(1), Cint(2)
const PREC_F64, PREC_F32, PREC_AUTO, PREC_BF16X3, PREC_BF16X1, PREC_BF16X2 = Cint(0), Cint(1), Cint(2), Cint(3), Cint(4), Cint(5)
const PREC_FP16X2, PREC_FP16X1 = Cint(6), Cint(7)

lasterror(ctx) = unsafe_string(ccall((:isb_last_error, libisb), Cstring, (Ctx,), ctx))
check(rc, ctx) = rc == 0 ? nothing : error(lasterror(ctx))

function create(device::Integer)
    h = Ref{Ctx}(C_NULL)
    rc = ccall((:isb_create, libisb), Cint, (Cint, Ref{Ctx}), device, h)
    rc == 0 || error(lasterror(C_NULL))
    h[]
end
const _ctx = Ref{Ctx}(C_NULL)
context() = (_ctx[] == C_NULL && (_ctx[] = create(parse(Int, get(ENV, "LOCAL_RANK", "0")))); _ctx[])

function model_dense(J::Matrix{Float64}, h::Vector{Float64}; prec = PREC_AUTO)
    m = Ref{Model}(C_NULL); w = Ref{Cint}(0)
    check(ccall((:isb_model_dense, libisb), Cint,
                (Ctx, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Cint, Ref{Cint}, Ref{Model}),
                context(), size(J, 1), J, stride(J, 2), h, prec, w, m), context())
    m[]
end
# SparseMatrixCSC couplings (what the reference's tests and demo pass) stay sparse: 0-based CSC across the ABI
function model_sparse(n::Integer, colptr::Vector{Int64}, rowval::Vector{Int32}, nzval::Vector{Float64}, h::Vector{Float64})
    m = Ref{Model}(C_NULL); w = Ref{Cint}(0)
    check(ccall((:isb_model_sparse, libisb), Cint,
                (Ctx, Cint, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ref{Cint}, Ref{Model}),
                context(), n, colptr, rowval, nzval, h, w, m), context())
    m[]
end
function model_bipartite(W::Matrix{Float64}, h::Vector{Float64}, b::Vector{Float64}; prec = PREC_F64)
    m = Ref{Model}(C_NULL)
    check(ccall((:isb_model_bipartite, libisb), Cint,
                (Ctx, Cint, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Cint, Ref{Model}),
                context(), size(W, 1), size(W, 2), W, stride(W, 2), h, b, prec, m), context())
    m[]
end
function ensemble(m::Model, R::Integer)
    e = Ref{Ens}(C_NULL)
    check(ccall((:isb_ens_create, libisb), Cint, (Model, Cint, Ref{Ens}), m, R, e), context())
    e[]
end
# spins cross the ABI as Int8, one replica per column of an N x R matrix (== [R][ld] replica-major in C)
set_spins!(e, S::Matrix{Int8}) = check(ccall((:isb_ens_set_spins, libisb), Cint, (Ens, Ptr{Int8}, Int64), e, S, stride(S, 2)), context())
get_spins!(e, S::Matrix{Int8}) = (check(ccall((:isb_ens_get_spins, libisb), Cint, (Ens, Ptr{Int8}, Int64), e, S, stride(S, 2)), context()); S)
set_hidden!(e, S::Matrix{Int8}) = check(ccall((:isb_ens_set_hidden, libisb), Cint, (Ens, Ptr{Int8}, Int64), e, S, stride(S, 2)), context())
get_hidden!(e, S::Matrix{Int8}) = (check(ccall((:isb_ens_get_hidden, libisb), Cint, (Ens, Ptr{Int8}, Int64), e, S, stride(S, 2)), context()); S)
energy(e, R) = (E = Vector{Float64}(undef, R); check(ccall((:isb_ens_energy, libisb), Cint, (Ens, Ptr{Float64}), e, E), context()); E)
function local_field(e, n, R)
    F = Matrix{Float64}(undef, n, R)
    check(ccall((:isb_ens_local_field, libisb), Cint, (Ens, Ptr{Float64}, Int64), e, F, n), context()); F
end
function local_aux_bias(e, n, R)
    F = Matrix{Float64}(undef, n, R)
    check(ccall((:isb_ens_local_aux_bias, libisb), Cint, (Ens, Ptr{Float64}, Int64), e, F, n), context()); F
end
# nodes are 1-based on the Julia side, 0-based across the ABI
function ssf_run!(e, rule, nsteps; nodes = nothing, start = 1, fluct = nothing, per_replica = false, seed = 0,
                  step_offset = 0, T = Float64[], steps_per_T = 1)
    n0 = nodes === nothing ? Ptr{Int32}(C_NULL) : Int32.(nodes .- 1)
    order = nodes === nothing ? ORDER_SEQUENTIAL : ORDER_LIST
    mode = fluct === nothing ? FLUCT_PHILOX : (per_replica ? FLUCT_PER_REPLICA : FLUCT_SHARED)
    f = fluct === nothing ? Ptr{Float64}(C_NULL) : Float64.(fluct)
    check(ccall((:isb_ssf_run, libisb), Cint,
                (Ens, Cint, Int64, Cint, Ptr{Int32}, Cint, Cint, Ptr{Float64}, UInt64, UInt64, Ptr{Float64}, Int64,
                 Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                e, rule, nsteps, order, n0, start - 1, mode, f, seed, step_offset, T, length(T), steps_per_T, 0,
                C_NULL, C_NULL, C_NULL), context())
end
function bip_run!(e, rule, nsteps; Fv = nothing, Fh = nothing, per_replica = false, seed = 0, step_offset = 0,
                  T = Float64[], steps_per_T = 1)
    mode = Fv === nothing ? FLUCT_PHILOX : (per_replica ? FLUCT_PER_REPLICA : FLUCT_SHARED)
    fv = Fv === nothing ? Ptr{Float64}(C_NULL) : Float64.(Fv)   # (units, steps) column-major == [steps][units]
    fh = Fh === nothing ? Ptr{Float64}(C_NULL) : Float64.(Fh)
    check(ccall((:isb_bip_run, libisb), Cint,
                (Ens, Cint, Int64, Cint, Ptr{Float64}, Ptr{Float64}, UInt64, UInt64, Ptr{Float64}, Int64, Int64,
                 Int64, Ptr{Float64}),
                e, rule, nsteps, mode, fv, fh, seed, step_offset, T, length(T), steps_per_T, 0, C_NULL), context())
end
end # module CABI

# ------------------------------------------------------------------ SpinSystems (src/SpinSystems.jl)
module SpinSystems
export SpinSystem, UpdatingAlgorithm, getSpinConfiguration, getCouplingCoefficients, getExternalMagneticField
export calcEnergy, calcLocalMagneticField, SpinSystemOnBipartiteGraph, UpdatingAlgorithmOnBipartiteGraph
export getHiddenLayer, getAuxiliaryBias, calcLocalAuxiliaryBias
using LinearAlgebra
using SparseArrays
import ..CABI

mutable struct SpinSystem
    spinConfiguration::AbstractVecOrMat{<:Number}
    couplingCoefficients::AbstractMatrix{<:AbstractFloat}
    externalMagneticField::AbstractVector{<:AbstractFloat}
    model::CABI.Model
    ens::CABI.Ens
    function SpinSystem(spinConfiguration, couplingCoefficients, externalMagneticField)
        # the reference's checks, verbatim in behaviour (src/SpinSystems.jl:19-49)
        numNodes = size(spinConfiguration, 1)
        (row, column) = size(couplingCoefficients)
        if row != column
            error("The coupling-coefficient matrix is not a square matrix: $(row)rows ≠ $(column)columns.")
        elseif numNodes < row
            @warn "The size of the spin-configuration vector is too smaller than the size of the coupling-coefficient matrix.  The incorresponding components of the coupling-coefficient matrix are ignored."
            couplingCoefficients = couplingCoefficients[1:numNodes, 1:numNodes]
        elseif numNodes > row
            @warn "The size of the spin-configuration vector is too bigger than the size of the coupling-coefficient matrix.  The incorresponding components of the spin-configuration vector are ignored."
            spinConfiguration = spinConfiguration[1:row, :]
        elseif !issymmetric(couplingCoefficients)
            @warn "The coupling-coefficient matrix should be symmetric.  It is symmetrized by its upper-triangular components automatically."
            couplingCoefficients = Symmetric(couplingCoefficients, :U)
        end
        if any(diag(couplingCoefficients) .!= 0)
            @warn "The diagonal components of the coupling-coefficient matrix should be zero.  Their non-zero components are ignored."
            couplingCoefficients -= Diagonal(couplingCoefficients)
        end
        numBias = length(externalMagneticField)
        if size(couplingCoefficients, 1) != numBias
            error("The size of the coupling-coefficient matrix does not match the size of the external-magnetic-field vector: $(row) ≠ $(numBias).")
        end
        h = Vector{Float64}(externalMagneticField)
        S = Matrix{Int8}(reshape(spinConfiguration, size(spinConfiguration, 1), :))
        if couplingCoefficients isa SparseMatrixCSC
            J = SparseMatrixCSC{Float64,Int64}(couplingCoefficients)
            m = CABI.model_sparse(size(J, 1), J.colptr .- 1, Int32.(J.rowval .- 1), J.nzval, h)
        else
            J = Matrix{Float64}(couplingCoefficients)
            m = CABI.model_dense(J, h)
        end
        e = CABI.ensemble(m, size(S, 2)); CABI.set_spins!(e, S)
        new(spinConfiguration, J, h, m, e)
    end
end
abstract type UpdatingAlgorithm end

_pull!(ss::SpinSystem) = (S = CABI.get_spins!(ss.ens, Matrix{Int8}(undef, length(ss.externalMagneticField), size(ss.spinConfiguration, 2)));
                          ss.spinConfiguration = ndims(ss.spinConfiguration) == 1 ? vec(Int.(S)) : Int.(S))
getSpinConfiguration(ua::UpdatingAlgorithm) = ua.spinSystem.spinConfiguration
getCouplingCoefficients(ua::UpdatingAlgorithm) = ua.spinSystem.couplingCoefficients
getExternalMagneticField(ua::UpdatingAlgorithm) = ua.spinSystem.externalMagneticField
_scalar(ss, v) = ndims(ss.spinConfiguration) == 1 ? v[1] : v
calcEnergy(ss::SpinSystem) = _scalar(ss, CABI.energy(ss.ens, size(ss.spinConfiguration, 2)))            # :68-71
calcEnergy(ua::UpdatingAlgorithm) = calcEnergy(ua.spinSystem)
calcLocalMagneticField(ss::SpinSystem) = (F = CABI.local_field(ss.ens, length(ss.externalMagneticField), size(ss.spinConfiguration, 2));
                                          ndims(ss.spinConfiguration) == 1 ? vec(F) : F)                 # :75-78
calcLocalMagneticField(ss::SpinSystem, i::Integer) = _scalar(ss, CABI.local_field(ss.ens, length(ss.externalMagneticField), size(ss.spinConfiguration, 2))[i, :])
calcLocalMagneticField(ua::UpdatingAlgorithm) = calcLocalMagneticField(ua.spinSystem)
calcLocalMagneticField(ua::UpdatingAlgorithm, i::Integer) = calcLocalMagneticField(ua.spinSystem, i)

mutable struct SpinSystemOnBipartiteGraph
    spinConfiguration::AbstractVecOrMat{<:Number}
    hiddenLayer::AbstractVecOrMat{<:Number}
    couplingCoefficients::AbstractMatrix{<:AbstractFloat}
    externalMagneticField::AbstractVector{<:AbstractFloat}
    auxiliaryBias::AbstractVector{<:AbstractFloat}
    model::CABI.Model
    ens::CABI.Ens
    function SpinSystemOnBipartiteGraph(spinConfiguration, hiddenLayer, couplingCoefficients, externalMagneticField, auxiliaryBias; prec = CABI.PREC_F64)
        nv = size(spinConfiguration, 1); nh = size(hiddenLayer, 1)
        (row, column) = size(couplingCoefficients)
        row != nv && error("The size of the coupling-coefficient matrix does not match the number of visible and hidden nodes: $(nv)nodes ≠ $(row)rows.")
        column != nh && error("The size of the coupling-coefficient matrix does not match the number of visible and hidden nodes: $(nh)nodes ≠ $(column)columns.")
        length(externalMagneticField) != nv && error("The size of the external-magnetic-field vector does not match the number of visible nodes: $(length(externalMagneticField)) ≠ $(nv).")
        length(auxiliaryBias) != nh && error("The size of the eauxiliary-bias vector does not match the number of hidden nodes: $(length(auxiliaryBias)) ≠ $(nh).")
        W = Matrix{Float64}(couplingCoefficients); h = Vector{Float64}(externalMagneticField); b = Vector{Float64}(auxiliaryBias)
        S = Matrix{Int8}(reshape(spinConfiguration, nv, :)); T = Matrix{Int8}(reshape(hiddenLayer, nh, :))
        m = CABI.model_bipartite(W, h, b; prec = prec); e = CABI.ensemble(m, size(S, 2))
        CABI.set_spins!(e, S); CABI.set_hidden!(e, T)
        new(spinConfiguration, hiddenLayer, W, h, b, m, e)
    end
end
abstract type UpdatingAlgorithmOnBipartiteGraph end
function _pull!(ss::SpinSystemOnBipartiteGraph)
    R = size(ss.spinConfiguration, 2)
    S = CABI.get_spins!(ss.ens, Matrix{Int8}(undef, length(ss.externalMagneticField), R))
    T = CABI.get_hidden!(ss.ens, Matrix{Int8}(undef, length(ss.auxiliaryBias), R))
    one = ndims(ss.spinConfiguration) == 1
    ss.spinConfiguration = one ? vec(Float64.(S)) : Float64.(S)   # the reference replaces the layers by Vector{Float64}
    ss.hiddenLayer = one ? vec(Float64.(T)) : Float64.(T)         # (src/OnBipartiteGraph.jl:35-42)
end
getSpinConfiguration(ua::UpdatingAlgorithmOnBipartiteGraph) = ua.spinSystem.spinConfiguration
getHiddenLayer(ua::UpdatingAlgorithmOnBipartiteGraph) = ua.spinSystem.hiddenLayer
getCouplingCoefficients(ua::UpdatingAlgorithmOnBipartiteGraph) = ua.spinSystem.couplingCoefficients
getExternalMagneticField(ua::UpdatingAlgorithmOnBipartiteGraph) = ua.spinSystem.externalMagneticField
getAuxiliaryBias(ua::UpdatingAlgorithmOnBipartiteGraph) = ua.spinSystem.auxiliaryBias
calcEnergy(ss::SpinSystemOnBipartiteGraph) = (E = CABI.energy(ss.ens, size(ss.spinConfiguration, 2)); ndims(ss.spinConfiguration) == 1 ? E[1] : E)
calcEnergy(ua::UpdatingAlgorithmOnBipartiteGraph) = calcEnergy(ua.spinSystem)
calcLocalMagneticField(ss::SpinSystemOnBipartiteGraph) = CABI.local_field(ss.ens, length(ss.externalMagneticField), size(ss.spinConfiguration, 2))
calcLocalMagneticField(ua::UpdatingAlgorithmOnBipartiteGraph) = calcLocalMagneticField(ua.spinSystem)
calcLocalAuxiliaryBias(ss::SpinSystemOnBipartiteGraph) = CABI.local_aux_bias(ss.ens, length(ss.auxiliaryBias), size(ss.spinConfiguration, 2))
calcLocalAuxiliaryBias(ua::UpdatingAlgorithmOnBipartiteGraph) = calcLocalAuxiliaryBias(ua.spinSystem)
heaviside(x::T; c::T = one(T)) where {T<:Number} = x > zero(T) ? one(T) : (x < zero(T) ? zero(T) : c)   # :163-171
end # module SpinSystems

# ------------------------------------------------------------------ SingleSpinFlip (src/SingleSpinFlip.jl)
module SingleSpinFlip
export update!, AsynchronousHopfieldNetwork, GlauberDynamics, MetropolisMethod
using Distributions
using ..SpinSystems
import ..CABI
abstract type SingleSpinUpdatingAlgorithm <: UpdatingAlgorithm end
mutable struct AsynchronousHopfieldNetwork <: SingleSpinUpdatingAlgorithm
    spinSystem::SpinSystem
    distribution::ContinuousUnivariateDistribution
    AsynchronousHopfieldNetwork(spinSystem::SpinSystem) = new(spinSystem, Uniform())
end
mutable struct GlauberDynamics <: SingleSpinUpdatingAlgorithm
    spinSystem::SpinSystem
    temperature::AbstractFloat
    distribution::ContinuousUnivariateDistribution
    GlauberDynamics(spinSystem::SpinSystem, temperature::AbstractFloat) = new(spinSystem, temperature, Logistic())
end
mutable struct MetropolisMethod <: SingleSpinUpdatingAlgorithm
    spinSystem::SpinSystem
    temperature::AbstractFloat
    distribution::ContinuousUnivariateDistribution
    MetropolisMethod(spinSystem::SpinSystem, temperature::AbstractFloat) = new(spinSystem, temperature, Exponential())
end
rule(::AsynchronousHopfieldNetwork) = CABI.RULE_HOPFIELD
rule(::GlauberDynamics) = CABI.RULE_GLAUBER
rule(::MetropolisMethod) = CABI.RULE_METROPOLIS
temperature(ua) = hasproperty(ua, :temperature) ? Float64(ua.temperature) : 0.0
# update!(ua, updatedNode, fluctuation): src/SingleSpinFlip.jl:31-36, 46-55, 65-74
function update!(ua::SingleSpinUpdatingAlgorithm, updatedNode::Integer, fluctuation::AbstractFloat = 0.0)
    temperature(ua) < 0 && @warn "$(ua.temperature) is negative."
    CABI.ssf_run!(ua.spinSystem.ens, rule(ua), 1; nodes = [updatedNode], fluct = [fluctuation], T = [temperature(ua)])
    SpinSystems._pull!(ua.spinSystem)
    ua.spinSystem.spinConfiguration[updatedNode, :]
end
end # module SingleSpinFlip

# ------------------------------------------------------------------ OnBipartiteGraph (src/OnBipartiteGraph.jl)
module OnBipartiteGraph
export update!, makeSampler!, StochasticCellularAutomata
using Distributions
using ..SpinSystems
import ..CABI
mutable struct StochasticCellularAutomata <: UpdatingAlgorithmOnBipartiteGraph
    spinSystem::SpinSystemOnBipartiteGraph
    temperature::AbstractFloat
    distribution::ContinuousUnivariateDistribution
    StochasticCellularAutomata(spinSystem::SpinSystemOnBipartiteGraph, temperature::AbstractFloat) = new(spinSystem, temperature, Logistic())
end
mutable struct MomentumAnnealing <: UpdatingAlgorithmOnBipartiteGraph
    spinSystem::SpinSystemOnBipartiteGraph
    temperature::AbstractFloat
    distribution::ContinuousUnivariateDistribution
    MomentumAnnealing(spinSystem::SpinSystemOnBipartiteGraph, temperature::AbstractFloat) = new(spinSystem, temperature, Exponential())
end
rule(::StochasticCellularAutomata) = CABI.BIP_SCA
rule(::MomentumAnnealing) = CABI.BIP_MA
# update!(ua, Fv, Fh): src/OnBipartiteGraph.jl:30-43, 53-66
function update!(ua::UpdatingAlgorithmOnBipartiteGraph, fluctuationForSpinConfiguration::AbstractVector{<:AbstractFloat},
                 fluctuationForHiddenLayer::AbstractVector{<:AbstractFloat})
    ua.temperature < 0 && @warn "$(ua.temperature) is negative."
    CABI.bip_run!(ua.spinSystem.ens, rule(ua), 1; Fv = fluctuationForSpinConfiguration, Fh = fluctuationForHiddenLayer,
                  T = [Float64(ua.temperature)])
    SpinSystems._pull!(ua.spinSystem)
    ua.spinSystem.spinConfiguration
end
end # module OnBipartiteGraph

# ------------------------------------------------------------------ MultiSpinFlip (stub in the reference: src/MultiSpinFlip.jl)
module MultiSpinFlip
export update!, makeSampler!, StochasticCellularAutomata
using LinearAlgebra
using ..SpinSystems
import ..OnBipartiteGraph
abstract type MultiSpinUpdatingAlgorithm <: UpdatingAlgorithm end
# The general-graph SCA as the bipartite embedding of demo.jl:82-90: W = (J + qI)/2, biases h/2, sigma = tau = s.
mutable struct StochasticCellularAutomata <: MultiSpinUpdatingAlgorithm
    spinSystem::SpinSystem
    bipartite::OnBipartiteGraph.StochasticCellularAutomata
    pinningParameter::Float64
    function StochasticCellularAutomata(ss::SpinSystem, temperature::AbstractFloat;
                                        pinningParameter = 0.5 * eigmax(Symmetric(ss.couplingCoefficients)))
        s = ss.spinConfiguration; J = ss.couplingCoefficients; h = ss.externalMagneticField
        b = SpinSystemOnBipartiteGraph(copy(s), copy(s), 0.5 * (J + pinningParameter * I), 0.5 * h, 0.5 * h)
        new(ss, OnBipartiteGraph.StochasticCellularAutomata(b, temperature), pinningParameter)
    end
end
function update!(ua::StochasticCellularAutomata, Fv::AbstractVector{<:AbstractFloat}, Fh::AbstractVector{<:AbstractFloat})
    OnBipartiteGraph.update!(ua.bipartite, Fv, Fh)
    ua.spinSystem.spinConfiguration = Int.(ua.bipartite.spinSystem.spinConfiguration)
end
end # module MultiSpinFlip

# ------------------------------------------------------------------ SamplingHelper (src/SamplingHelper.jl)
module SamplingHelper
export update!, makeSampler!
using Random
using ..SpinSystems
import ..SingleSpinFlip, ..OnBipartiteGraph, ..MultiSpinFlip
import ..CABI

function update!(ua::SingleSpinFlip.SingleSpinUpdatingAlgorithm; rng::AbstractRNG = Random.default_rng())   # :22-26
    updatedNode = rand(rng, axes(getSpinConfiguration(ua), 1))
    fluctuation = rand(rng, ua.distribution)
    SingleSpinFlip.update!(ua, updatedNode, fluctuation)
end
function update!(ua::UpdatingAlgorithmOnBipartiteGraph; rng::AbstractRNG = Random.default_rng())           # :104-108
    Fv = rand(rng, ua.distribution, size(ua.spinSystem.spinConfiguration, 1))
    Fh = rand(rng, ua.distribution, size(ua.spinSystem.hiddenLayer, 1))
    OnBipartiteGraph.update!(ua, Fv, Fh)
end

# makeSampler!: same draw order, schedule timing and n+1-item Channel contract as src/SamplingHelper.jl:28-51;
# `stride` > 1 (extension) runs `stride` steps per ccall and yields after each chunk.
function makeSampler!(ua::SingleSpinFlip.SingleSpinUpdatingAlgorithm, maxMCSteps::Integer;
                      annealingSchedule::Function = n -> ua.temperature, rng::AbstractRNG = Random.default_rng(),
                      stride::Integer = 1)::Channel{SpinSystems.UpdatingAlgorithm}
    maxMCSteps < 0 && @warn "$maxMCSteps is negative."
    hasT = hasproperty(ua, :temperature)
    updatedNodes = rand(rng, axes(getSpinConfiguration(ua), 1), maxMCSteps)
    fluctuations = rand(rng, ua.distribution, maxMCSteps)
    Channel{SpinSystems.UpdatingAlgorithm}() do channel
        hasT && (ua.temperature = annealingSchedule(0))
        put!(channel, ua)
        k = 0
        while k < maxMCSteps
            m = min(stride, maxMCSteps - k)
            T = hasT ? Float64[annealingSchedule(j) for j in k+1:k+m] : Float64[0.0]
            CABI.ssf_run!(ua.spinSystem.ens, SingleSpinFlip.rule(ua), m; nodes = updatedNodes[k+1:k+m],
                          fluct = fluctuations[k+1:k+m], T = T, steps_per_T = hasT ? 1 : m)
            k += m
            hasT && (ua.temperature = annealingSchedule(k))
            SpinSystems._pull!(ua.spinSystem)
            put!(channel, ua)
        end
    end
end
function makeSampler!(ua::UpdatingAlgorithmOnBipartiteGraph, maxMCSteps::Integer;
                      annealingSchedule::Function = n -> ua.temperature, rng::AbstractRNG = Random.default_rng(),
                      stride::Integer = 1)::Channel{SpinSystems.UpdatingAlgorithmOnBipartiteGraph}
    maxMCSteps < 0 && @warn "$maxMCSteps is negative."
    Fv = rand(rng, ua.distribution, (size(getSpinConfiguration(ua), 1), maxMCSteps))   # :121
    Fh = rand(rng, ua.distribution, (size(getHiddenLayer(ua), 1), maxMCSteps))         # :122
    Channel{SpinSystems.UpdatingAlgorithmOnBipartiteGraph}() do channel
        ua.temperature = annealingSchedule(0)
        put!(channel, ua)
        k = 0
        while k < maxMCSteps
            m = min(stride, maxMCSteps - k)
            T = Float64[annealingSchedule(j) for j in k+1:k+m]
            CABI.bip_run!(ua.spinSystem.ens, OnBipartiteGraph.rule(ua), m; Fv = Fv[:, k+1:k+m], Fh = Fh[:, k+1:k+m], T = T)
            k += m
            ua.temperature = annealingSchedule(k)
            SpinSystems._pull!(ua.spinSystem)
            put!(channel, ua)
        end
    end
end
end # module SamplingHelper

end # module IsingModelB200
