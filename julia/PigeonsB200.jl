# PigeonsB200.jl — the reference-side binding of libpigeons_b200.so.
#
# This file is what a Pigeons.jl maintainer adds (as a package extension or a small companion package) so that
#
#     pt = pigeons(target = toy_mvn_target(100), n_chains = 256, explorer = AutoMALA(), on = B200())
#
# runs the inner scan loop (`run_one_round!`, src/pt/pigeons.jl:46-55) on a B200 through ONE `ccall` per round, while
# everything above it — `adapt`, `report!`, `write_checkpoint`, `run_checks`, `stepping_stone`, `sample_array` —
# stays the reference's own Julia code operating on the same `reduced_recorders` NamedTuple.
#
# Plug points (all are dispatch points the reference already has):
#   * `pigeons(pt_arguments, on::Submission)`            src/api.jl:8-19, src/submission/Submission.jl:4-9
#   * the informal `replicas` interface                   src/replicas/replicas.jl:11-40 (swap!, locals, load, communicator, entangler)
#   * `run_one_round!(pt)`                                src/pt/pigeons.jl:46-55
#   * `reduce_recorders!`                                 src/recorders/recorders.jl:88-120
# FFI conventions follow the only FFI the reference has (BridgeStan): `Cint` return codes, caller-owned output
# buffers, a library-owned error string released by the library (ext/PigeonsBridgeStanExt/interface.jl:118-183).
#
# STATUS: Julia is not installable in the build container, so this file has never been executed.  What IS checked
# mechanically: tests/test_abi_layout.py parses the `struct` blocks below and compares every field (name, order,
# C type, offset, size) with `sizeof/offsetof` printed by tests/abi_layout.c for include/pigeons_b200.h, and
# `PGN_ABI_VERSION` below with the header's.  Third-party constructors used by `materialise_recorders`
# (OnlineStatsBase `Mean`, `Sum`, `Variance`, `GroupBy`, `Group`) are written from the field layouts of
# OnlineStatsBase 1.x (SURVEY.md appendix A.4) and marked `# OnlineStatsBase 1.x` where they matter.
module PigeonsB200

using Pigeons
using Pigeons: PT, Inputs, Shared, Replica, Submission, LoadBalance, Entangler, SliceSampler, AutoMALA, MALA, Compose, Mix,
               IdentityPreconditioner, DiagonalPreconditioner, MixDiagonalPreconditioner, ScaledPrecisionNormalPath,
               RoundTripRecorder, LogSum, OnlineStateRecorder, recorder_builders, create_recorders, n_chains,
               n_scans_in_round, single_process_load
using OnlineStatsBase
using OnlineStatsBase: Mean, Sum, Variance, GroupBy, Group, EqualWeight
using OrderedCollections: OrderedDict
using SplittableRandoms: SplittableRandom

export B200, B200Replicas

# --------------------------------------------------------------------------------------------------------------------
# constants of include/pigeons_b200.h
# --------------------------------------------------------------------------------------------------------------------
const PGN_ABI_VERSION = 4

const PGN_OK = 0
const PGN_ERROR_NAMES = Dict(
    1 => "PGN_ERR_INVALID", 2 => "PGN_ERR_NO_DEVICE", 3 => "PGN_ERR_CUDA", 4 => "PGN_ERR_NAN_RATIO",
    5 => "PGN_ERR_BAD_DENSITY", 6 => "PGN_ERR_SLICE_MAX_ITER", 7 => "PGN_ERR_STEP_UNDERFLOW",
    8 => "PGN_ERR_NOT_POSITIVE", 9 => "PGN_ERR_TIMEOUT")

const PGN_TARGET_TOY_MVN = 1
const PGN_TARGET_FUNNEL = 2
const PGN_TARGET_GMM = 3
const PGN_TARGET_ISING = 4
const PGN_TARGET_LOGREG = 5
const PGN_TARGET_TEST_SWAPPER = 6
const PGN_TARGET_MIXED = 7
const PGN_TARGET_UNID = 8

const PGN_EXPLORER_NONE = 0
const PGN_EXPLORER_TOY = 1
const PGN_EXPLORER_SLICE = 2
const PGN_EXPLORER_AUTOMALA = 3
const PGN_EXPLORER_ISING_METROPOLIS = 4
const PGN_EXPLORER_MALA = 5
const PGN_EXPLORER_COMPOSE = 6
const PGN_EXPLORER_MIX = 7
const PGN_MAX_MIX = 4

const PGN_PRECOND_IDENTITY = 0
const PGN_PRECOND_DIAGONAL = 1
const PGN_PRECOND_MIX_DIAGONAL = 2

const PGN_RECORDERS_PER_REPLICA = 0
const PGN_RECORDERS_PER_CHAIN = 1

# --------------------------------------------------------------------------------------------------------------------
# struct mirrors, field for field (checked against the header by tests/test_abi_layout.py)
# --------------------------------------------------------------------------------------------------------------------
struct PgnConfig
    abi_version::Int32
    target_kind::Int32
    dim::Int32
    n_chains::Int32
    seed::Int64
    rank::Int32
    world_size::Int32
    device::Int32
    n_modes::Int32
    p::NTuple{8, Float64}
    means::Ptr{Float64}
    log_weights::Ptr{Float64}
    data_x::Ptr{Float64}
    data_y::Ptr{Float64}
    recorder_order::Int32
    n_chains_variational::Int32
end

struct PgnExplorerParams
    kind::Int32
    slice_w::Float64
    slice_p::Int32
    slice_n_passes::Int32
    slice_max_iter::Int32
    n_refresh::Int32
    step_size::Float64
    precond_kind::Int32
    mix_p0::Float64
    mix_p01::Float64
    std_devs::Ptr{Float64}
    ising_n_steps::Int32
    n_steps::Int32
    step_kind::NTuple{4, Int32}
    n_mix::Int32
    mix_n_refresh::NTuple{4, Int32}
    mix_precond_kind::NTuple{4, Int32}
    mix_step_size::NTuple{4, Float64}
    mix_variant_p0::NTuple{4, Float64}
    mix_variant_p01::NTuple{4, Float64}
end

struct PgnRoundOut
    swap_n::Ptr{Int64}
    swap_mean::Ptr{Float64}
    logsum_fwd::Ptr{Float64}
    logsum_bwd::Ptr{Float64}
    expl_acc_n::Ptr{Int64}
    expl_acc_mean::Ptr{Float64}
    expl_n_steps::Ptr{Int64}
    am_n::Ptr{Int64}
    am_mean::Ptr{Float64}
    rev_n::Ptr{Int64}
    rev_mean::Ptr{Float64}
    n_tempered_restarts::Int64
    n_round_trips::Int64
    online_n::Int64
    online_mean::Ptr{Float64}
    online_var::Ptr{Float64}
    index_process::Ptr{Int32}
    swap_lr::Ptr{Float64}
    swap_u::Ptr{Float64}
    swap_accept::Ptr{UInt8}
    target_trace::Ptr{Float64}
    n_density_points::Int64
    n_ref_equiv_evals::Int64
    kernel_ms::Float64
    gemm_ms::Float64
    batch_steps::Int64
    n_launches::Int64
    active_columns::Int64
    gemm_columns::Int64
end

struct PgnReplicaState
    x::Ptr{Float64}
    replica_index::Ptr{Int32}
    rng_counter::Ptr{UInt64}
    round_trip_state::Ptr{Int32}
end

struct PgnDeviceInfo
    sm_major::Int32
    sm_minor::Int32
    n_sms::Int32
    global_mem_bytes::Int64
    max_resident_chains::Int32
    name::NTuple{128, UInt8}
end

# --------------------------------------------------------------------------------------------------------------------
# library handle and error plumbing
# --------------------------------------------------------------------------------------------------------------------
const libpigeons_b200 = Ref{String}(get(ENV, "PIGEONS_B200_LIB", "libpigeons_b200.so"))

struct B200Error <: Exception
    code::Int
    msg::String
end
Base.showerror(io::IO, e::B200Error) = print(io, get(PGN_ERROR_NAMES, e.code, string(e.code)), ": ", e.msg)

"""
Non-zero return codes become exceptions, as the reference's own errors on this path are
(`log_potentials.jl:47-49`, `SliceSampler.jl:35-37,52-59`, `AutoMALA.jl:240-242`).  The message is owned by the
library (BridgeStan convention, interface.jl:118-183) and released with `pgn_free_string`.
"""
function check(rc::Cint, err::Ref{Cstring})
    rc == PGN_OK && return nothing
    msg = err[] == C_NULL ? "" : unsafe_string(err[])
    err[] == C_NULL || ccall((:pgn_free_string, libpigeons_b200[]), Cvoid, (Cstring,), err[])
    throw(B200Error(Int(rc), msg))
end

# --------------------------------------------------------------------------------------------------------------------
# the submission flag and the device-resident replicas container
# --------------------------------------------------------------------------------------------------------------------
"""
`pigeons(...; on = B200())` — run the scan loop on the GPU(s) of this machine.
`n_gpus > 1`: the chain ladder is split into contiguous blocks (`LoadBalance`, src/mpi_utils/LoadBalance.jl:70-73),
one engine handle per GPU, all driven by this one Julia process (one task per handle); neighbouring handles are
connected with `pgn_peer_attach`, so no MPI is involved.
"""
Base.@kwdef struct B200 <: Submission
    n_gpus::Int = 1
    devices::Vector{Int} = collect(0:(n_gpus - 1))
    recorder_order::Int = PGN_RECORDERS_PER_REPLICA      # the reference's per-replica recorders + tree merge (recorders.jl:88-120)
end

"""
The `replicas` of a run on the device: opaque engine handles plus host-side bookkeeping.  Replica states live in
HBM (SoA, chain order); `locals` materialises `Replica` objects on demand (checkpoints, `run_checks`).
"""
mutable struct B200Replicas
    handles::Vector{Ptr{Cvoid}}          # one per GPU / shard, in chain order
    first_chain::Vector{Int}
    n_local::Vector{Int}
    n_chains::Int
    dim::Int
    seed::Int
    keep_alive::Vector{Any}              # host arrays whose pointers were handed to pgn_create
    recorders_template                   # create_recorders(inputs, shared): key set and empty values
    builders                             # recorder_builders(inputs, shared), for the per-replica (empty) recorders
end

Pigeons.load(r::B200Replicas) = single_process_load(r.n_chains)
Pigeons.communicator(::B200Replicas) = nothing
Pigeons.entangler(r::B200Replicas) = Entangler(r.n_chains; parent_communicator = nothing, verbose = false)
Pigeons.locals(r::B200Replicas) = materialise_replicas(r)
# `swap!` never runs on the host for this container: the DEO swap is part of the device scan
Pigeons.swap!(pair_swapper, ::B200Replicas, swap_graph) =
    error("swap! runs inside pgn_run_round for B200Replicas (no host-side scan loop)")

function destroy!(r::B200Replicas)
    for h in r.handles
        h == C_NULL || ccall((:pgn_destroy, libpigeons_b200[]), Cint, (Ptr{Cvoid},), h)
    end
    empty!(r.handles)
end

# --------------------------------------------------------------------------------------------------------------------
# target -> closed device family (+ POD parameters); anything else raises: there is NO CPU fallback
# --------------------------------------------------------------------------------------------------------------------
normal_ref_params(sigma::Float64) = (sigma, log(sigma), 1.0 / (sigma * sigma))

"""
Returns `(target_kind, dim, p::NTuple{8,Float64}, n_modes, means, log_weights, data_x, data_y)`.
Extend by adding methods for other target types whose densities have a device implementation.
"""
function device_target end

pad8(v...) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, 8)

# toy_mvn_target(dim) (src/targets/toy_mvn_target.jl:8) is a ScaledPrecisionNormalPath
device_target(t::ScaledPrecisionNormalPath, reference) =
    (PGN_TARGET_TOY_MVN, t.dim, pad8(t.precision0, t.precision1), 0, nothing, nothing, nothing, nothing)

# Pigeons.TestSwapper (src/swap/pair_swapper.jl:100-149)
device_target(t::Pigeons.TestSwapper, reference) =
    (PGN_TARGET_TEST_SWAPPER, 0, pad8(t.constant_swap_accept_pr), 0, nothing, nothing, nothing, nothing)

"""
Targets that are not types of Pigeons itself (the funnel of test/supporting/dimensional-analysis.jl:33-47, the
Ising model of examples/ising.jl, mixtures given as `DistributionLogPotential(MixtureModel(...))`, a logistic
regression with an analytic gradient) declare their device mapping with one of these constructors.
"""
struct DeviceFunnel;   dim::Int; sigma_y::Float64; sigma_ref::Float64; end
struct DeviceIsing;    beta::Float64; L::Int; end
struct DeviceMixture;  means::Matrix{Float64}; log_weights::Vector{Float64}; sigma::Float64; sigma_ref::Float64; end   # means: K x d
struct DeviceLogistic; x::Matrix{Float64}; y::Vector{Float64}; prior_sigma::Float64; end                               # x: n x d
# the unidentifiable product of test/test_DistributionLogPotential.jl:7-21 (reference: Uniform(0,1)^2, SliceSampler only)
struct DeviceUnid;     n_trials::Int; n_successes::Int; end

device_target(t::DeviceFunnel, reference) =
    (PGN_TARGET_FUNNEL, t.dim, pad8(t.sigma_y, log(t.sigma_y), 1.0 / t.sigma_y^2, normal_ref_params(t.sigma_ref)...), 0,
     nothing, nothing, nothing, nothing)
device_target(t::DeviceIsing, reference) =
    (PGN_TARGET_ISING, t.L * t.L, pad8(t.beta, t.L), 0, nothing, nothing, nothing, nothing)
function device_target(t::DeviceMixture, reference)
    K, d = size(t.means)
    K <= 8 || error("the device mixture holds at most 8 components (no CPU fallback)")
    cst = d * log(t.sigma) + 0.5 * d * log(2pi)
    means_rowmajor = collect(transpose(t.means))      # C expects [K][d] row-major == Julia d x K column-major
    (PGN_TARGET_GMM, d, pad8(t.sigma, cst, 1.0 / t.sigma^2, normal_ref_params(t.sigma_ref)...), K,
     vec(means_rowmajor), copy(t.log_weights), nothing, nothing)
end
function device_target(t::DeviceLogistic, reference)
    n, d = size(t.x)
    x_rowmajor = collect(transpose(t.x))              # [n][d] row-major
    (PGN_TARGET_LOGREG, d, pad8(n, 0.0, 0.0, normal_ref_params(t.prior_sigma)...), 0, nothing, nothing, vec(x_rowmajor), copy(t.y))
end

device_target(t::DeviceUnid, reference) =
    (PGN_TARGET_UNID, 2, pad8(t.n_trials, t.n_successes), 0, nothing, nothing, nothing, nothing)

device_target(t, reference) =
    error("no device implementation for a target of type $(typeof(t)): the B200 engine supports a closed family of " *
          "targets and has no CPU fallback (run with `on = ThisProcess()` instead)")

# --------------------------------------------------------------------------------------------------------------------
# explorer -> PgnExplorerParams
# --------------------------------------------------------------------------------------------------------------------
precond_code(::IdentityPreconditioner) = (PGN_PRECOND_IDENTITY, 1 / 3, 2 / 3)
precond_code(::DiagonalPreconditioner) = (PGN_PRECOND_DIAGONAL, 1 / 3, 2 / 3)
precond_code(p::MixDiagonalPreconditioner) = (PGN_PRECOND_MIX_DIAGONAL, Float64(p.p0), Float64(p.p0 + p.p1))   # Preconditioner.jl:64-69

n_refresh(e, dim) = e.base_n_refresh * ceil(Int, dim^e.exponent_n_refresh)                                       # AutoMALA.jl:122

zero4(T) = ntuple(_ -> zero(T), 4)

"""
Mirrors the `@kwdef` explorer structs after host-side adaptation (SliceSampler.jl:8-20, AutoMALA.jl:29-68,
MALA.jl, Compose.jl:16-19, Mix.jl:7-21).  `std_devs` is the array whose pointer goes into the struct; the caller
keeps it alive (`GC.@preserve`) across the `ccall`.
Returns `(PgnExplorerParams, std_devs_or_nothing)`.
"""
function explorer_params(explorer, dim::Int)
    kind = PGN_EXPLORER_NONE
    slice = SliceSampler()
    nref, step, pk, p0, p01 = 0, 1.0, PGN_PRECOND_IDENTITY, 1 / 3, 2 / 3
    sd = nothing
    n_mix = 0
    n_steps = 0
    step_kind = zero4(Int32)
    mix_nr, mix_pk, mix_ss, mix_p0, mix_p01 = zero4(Int32), zero4(Int32), zero4(Float64), zero4(Float64), zero4(Float64)
    ising_steps = 3
    if explorer === nothing                           # TestSwapper: step! is a no-op
        kind = PGN_EXPLORER_NONE
    elseif explorer isa Pigeons.ToyExplorer
        kind = PGN_EXPLORER_TOY
    elseif explorer isa SliceSampler
        kind, slice = PGN_EXPLORER_SLICE, explorer
    elseif explorer isa AutoMALA || explorer isa MALA
        kind = explorer isa AutoMALA ? PGN_EXPLORER_AUTOMALA : PGN_EXPLORER_MALA
        nref, step = n_refresh(explorer, dim), explorer.step_size
        pk, p0, p01 = precond_code(explorer.preconditioner)
        sd = explorer.estimated_target_std_deviations
    elseif explorer isa Mix && all(e -> e isa AutoMALA, explorer.explorers) && 2 <= length(explorer.explorers) <= PGN_MAX_MIX
        # a mixture of autoMALA kernels runs on the plain autoMALA kernel, which draws the variant itself
        kind = PGN_EXPLORER_AUTOMALA
        es = explorer.explorers
        n_mix = length(es)
        codes = map(e -> precond_code(e.preconditioner), es)
        fill4(f, T) = ntuple(i -> i <= n_mix ? T(f(i)) : zero(T), 4)
        mix_nr = fill4(i -> n_refresh(es[i], dim), Int32)
        mix_pk = fill4(i -> codes[i][1], Int32)
        mix_ss = fill4(i -> es[i].step_size, Float64)
        mix_p0 = fill4(i -> codes[i][2], Float64)
        mix_p01 = fill4(i -> codes[i][3], Float64)
        nref, step = n_refresh(es[1], dim), es[1].step_size
        pk, p0, p01 = codes[1]
        sd = es[1].estimated_target_std_deviations    # every variant adapts from the same recorders (Mix.jl:14-17)
    elseif (explorer isa Compose || explorer isa Mix) && 1 <= length(explorer.explorers) <= PGN_MAX_MIX
        # general program: every explorer in turn (Compose.jl:16-19) or one drawn uniformly (Mix.jl:20-21)
        kind = explorer isa Compose ? PGN_EXPLORER_COMPOSE : PGN_EXPLORER_MIX
        es = explorer.explorers
        n_steps = length(es)
        step_code(e) = e isa SliceSampler ? PGN_EXPLORER_SLICE : e isa AutoMALA ? PGN_EXPLORER_AUTOMALA :
                       e isa MALA ? PGN_EXPLORER_MALA : e isa Pigeons.ToyExplorer ? PGN_EXPLORER_TOY :
                       error("$(typeof(e)) cannot be part of a device Compose / Mix (no CPU fallback)")
        grad(e) = e isa AutoMALA || e isa MALA
        fills(f, T) = ntuple(i -> i <= n_steps ? T(f(es[i])) : zero(T), 4)
        step_kind = fills(step_code, Int32)
        mix_nr = fills(e -> grad(e) ? n_refresh(e, dim) : 0, Int32)
        mix_pk = fills(e -> grad(e) ? precond_code(e.preconditioner)[1] : PGN_PRECOND_IDENTITY, Int32)
        mix_ss = fills(e -> grad(e) ? e.step_size : 1.0, Float64)
        mix_p0 = fills(e -> grad(e) ? precond_code(e.preconditioner)[2] : 1 / 3, Float64)
        mix_p01 = fills(e -> grad(e) ? precond_code(e.preconditioner)[3] : 2 / 3, Float64)
        for e in es
            e isa SliceSampler && (slice = e)          # the SliceSamplers of one program share their parameters
            grad(e) && sd === nothing && (sd = e.estimated_target_std_deviations)
        end
    elseif hasproperty(explorer, :n_steps) && nameof(typeof(explorer)) == :IsingMetropolis     # examples/ising.jl:91-95
        kind, ising_steps = PGN_EXPLORER_ISING_METROPOLIS, explorer.n_steps
    else
        error("no device implementation for an explorer of type $(typeof(explorer)) (no CPU fallback)")
    end
    sd_vec = sd === nothing ? nothing : Vector{Float64}(sd)
    ep = PgnExplorerParams(kind, slice.w, slice.p, slice.n_passes, slice.max_iter, nref, step, pk, p0, p01,
                           sd_vec === nothing ? Ptr{Float64}(C_NULL) : pointer(sd_vec), ising_steps,
                           n_steps, step_kind, n_mix, mix_nr, mix_pk, mix_ss, mix_p0, mix_p01)
    return ep, sd_vec
end

# --------------------------------------------------------------------------------------------------------------------
# create_replicas (src/replicas/replicas.jl:65-99) for the device
# --------------------------------------------------------------------------------------------------------------------
function create_b200_replicas(inputs::Inputs, shared::Shared, on::B200)
    kind, dim, p, n_modes, means, log_w, data_x, data_y = device_target(inputs.target, inputs.reference)
    N = n_chains(inputs)
    keep = Any[means, log_w, data_x, data_y]
    ptr(a) = a === nothing ? Ptr{Float64}(C_NULL) : pointer(a)
    handles, firsts, counts = Ptr{Cvoid}[], Int[], Int[]
    for (rank, device) in enumerate(on.devices)
        cfg = PgnConfig(PGN_ABI_VERSION, kind, dim, N, inputs.seed, rank - 1, length(on.devices), device, n_modes, p,
                        ptr(means), ptr(log_w), ptr(data_x), ptr(data_y), on.recorder_order,
                        Pigeons.n_chains_var(inputs))      # chains 1..n_var: the variational leg (StabilizedPT.jl:96-116)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve means log_w data_x data_y begin
            err = Ref{Cstring}(C_NULL)
            rc = ccall((:pgn_create, libpigeons_b200[]), Cint, (Ref{PgnConfig}, Ref{Ptr{Cvoid}}, Ref{Cstring}), cfg, h, err)
            check(rc, err)
        end
        fc, nl = Ref{Int32}(0), Ref{Int32}(0)
        ccall((:pgn_local_range, libpigeons_b200[]), Cint, (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}), h[], fc, nl)
        push!(handles, h[]); push!(firsts, fc[]); push!(counts, nl[])
    end
    for i in eachindex(handles)                       # neighbour mailboxes: same process, plain peer pointers
        err = Ref{Cstring}(C_NULL)
        if i > 1
            check(ccall((:pgn_peer_attach, libpigeons_b200[]), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{Cstring}),
                        handles[i], 0, handles[i - 1], err), err)
        end
        if i < length(handles)
            check(ccall((:pgn_peer_attach, libpigeons_b200[]), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{Cstring}),
                        handles[i], 1, handles[i + 1], err), err)
        end
    end
    for h in handles                                  # initialization(target, rng, i) + chain = replica index
        err = Ref{Cstring}(C_NULL)
        check(ccall((:pgn_init_replicas, libpigeons_b200[]), Cint, (Ptr{Cvoid}, Ref{Cstring}), h, err), err)
    end
    replicas = B200Replicas(handles, firsts, counts, N, dim, inputs.seed, keep, create_recorders(inputs, shared),
                            recorder_builders(inputs, shared))
    finalizer(destroy!, replicas)
    return replicas
end

# --------------------------------------------------------------------------------------------------------------------
# entry point: pigeons(inputs, ::B200)
# --------------------------------------------------------------------------------------------------------------------
function Pigeons.pigeons(inputs::Inputs, on::B200)
    shared = Shared(inputs)
    replicas = create_b200_replicas(inputs, shared, on)
    exec_folder = Pigeons.pt_exec_folder(inputs.checkpoint, Pigeons.use_auto_exec_folder)
    pt = PT(inputs, replicas, shared, exec_folder, create_recorders(inputs, shared))
    return pigeons(pt)                                # the reference's own round loop (src/pt/pigeons.jl:12-28)
end

# --------------------------------------------------------------------------------------------------------------------
# run_one_round! — ONE engine call per round and per handle
# --------------------------------------------------------------------------------------------------------------------
struct RoundBuffers
    swap_n::Vector{Int64}; swap_mean::Vector{Float64}; logsum_fwd::Vector{Float64}; logsum_bwd::Vector{Float64}
    expl_acc_n::Vector{Int64}; expl_acc_mean::Vector{Float64}; expl_n_steps::Vector{Int64}
    am_n::Vector{Int64}; am_mean::Vector{Float64}; rev_n::Vector{Int64}; rev_mean::Vector{Float64}
    online_mean::Vector{Float64}; online_var::Vector{Float64}
    index_process::Union{Nothing, Matrix{Int32}}      # [n_local, n_scans] column-major == C [n_scans][n_local]
    swap_lr::Union{Nothing, Matrix{Float64}}; swap_u::Union{Nothing, Matrix{Float64}}; swap_accept::Union{Nothing, Matrix{UInt8}}
    target_trace::Union{Nothing, Matrix{Float64}}     # [d, n_scans]
end

"""
Caller-allocated outputs of `pgn_run_round` for one handle.  Event logs are allocated only for the recorders the
run asked for (`index_process`, `traces`; the SwapStat log is an engine extension used by `materialise_recorders`
when `parity_mode` replays the recorder arithmetic in the reference's order).
"""
function allocate_round_out(n_local::Int, d::Int, n_scans::Int; index_process::Bool, traces::Bool, swap_log::Bool, n_targets::Int = 1)
    z64(n) = zeros(Int64, n); zf(n) = zeros(Float64, n)
    return RoundBuffers(z64(n_local), zf(n_local), zf(n_local), zf(n_local), z64(n_local), zf(n_local), z64(n_local),
                        z64(n_local), zf(n_local), z64(n_local), zf(n_local), zf(max(d, 1)), zf(max(d, 1)),
                        index_process ? zeros(Int32, n_local, n_scans) : nothing,
                        swap_log ? zeros(Float64, n_local, n_scans) : nothing,
                        swap_log ? zeros(Float64, n_local, n_scans) : nothing,
                        swap_log ? zeros(UInt8, n_local, n_scans) : nothing,
                        traces ? zeros(Float64, max(d, 1), n_targets * n_scans) : nothing)   # two legs: C [n_scans][2][d]
end

optr(a, T) = a === nothing ? Ptr{T}(C_NULL) : pointer(a)

round_out_struct(b::RoundBuffers) = PgnRoundOut(
    pointer(b.swap_n), pointer(b.swap_mean), pointer(b.logsum_fwd), pointer(b.logsum_bwd),
    pointer(b.expl_acc_n), pointer(b.expl_acc_mean), pointer(b.expl_n_steps),
    pointer(b.am_n), pointer(b.am_mean), pointer(b.rev_n), pointer(b.rev_mean),
    0, 0, 0, pointer(b.online_mean), pointer(b.online_var),
    optr(b.index_process, Int32), optr(b.swap_lr, Float64), optr(b.swap_u, Float64), optr(b.swap_accept, UInt8),
    optr(b.target_trace, Float64), 0, 0, 0.0, 0.0, 0, 0, 0, 0)

"""
The annealing parameter of every global chain: `schedule.grids`, or for two legs the parameters of
`vcat(variational_leg.log_potentials, reverse(fixed_leg.log_potentials))` (`StabilizedPT.jl:63-65`).
"""
tempering_parameters(t::Pigeons.NonReversiblePT) = Vector{Float64}(t.schedule.grids)
tempering_parameters(t::Pigeons.StabilizedPT) =
    vcat(Vector{Float64}(t.variational_leg.schedule.grids), reverse(Vector{Float64}(t.fixed_leg.schedule.grids)))

"""
Mean and standard deviation of the `GaussianReference` once `update_reference!` has run (`GaussianReference.jl:22-28`;
the path of the variational leg is then `InterpolatingPath(variational, target)`, `variational.jl:36-40`), else `nothing`.
"""
function variational_parameters(pt)
    v = pt.inputs.variational
    (v isa Pigeons.GaussianReference && haskey(v.mean, :singleton_variable)) || return nothing, nothing
    leg = pt.shared.tempering isa Pigeons.StabilizedPT ? pt.shared.tempering.variational_leg : pt.shared.tempering
    leg.path.ref isa Pigeons.GaussianReference || return nothing, nothing            # not activated yet
    return Vector{Float64}(v.mean[:singleton_variable]), Vector{Float64}(v.standard_deviation[:singleton_variable])
end

function Pigeons.run_one_round!(pt::PT{<:Any, B200Replicas})
    r = pt.replicas
    n_scans = n_scans_in_round(pt.shared.iterators)
    grids = tempering_parameters(pt.shared.tempering)
    var_mean, var_sd = variational_parameters(pt)
    ep, sd = explorer_params(pt.shared.explorer, r.dim)
    want_index = haskey(r.recorders_template, :index_process)
    want_traces = haskey(r.recorders_template, :traces)
    two_legs = pt.shared.tempering isa Pigeons.StabilizedPT
    bufs = [allocate_round_out(r.n_local[i], r.dim, n_scans; index_process = want_index, traces = want_traces, swap_log = false,
                               n_targets = two_legs ? 2 : 1)
            for i in eachindex(r.handles)]
    outs = Vector{PgnRoundOut}(undef, length(r.handles))
    timed = @timed begin
        # one task per handle: the calls block until the round is over, the kernels of neighbouring handles talk to each
        # other through their mailboxes while they run
        @sync for i in eachindex(r.handles)
            Threads.@spawn begin
                h = r.handles[i]
                err = Ref{Cstring}(C_NULL)
                GC.@preserve grids sd bufs var_mean var_sd begin
                    check(ccall((:pgn_set_schedule, libpigeons_b200[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Ref{Cstring}),
                                h, grids, length(grids), err), err)
                    if Pigeons.n_chains_var(pt.inputs) > 0
                        vptr(a) = a === nothing ? Ptr{Float64}(C_NULL) : pointer(a)
                        check(ccall((:pgn_set_variational, libpigeons_b200[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Cstring}),
                                    h, vptr(var_mean), vptr(var_sd), err), err)
                    end
                    check(ccall((:pgn_set_explorer, libpigeons_b200[]), Cint, (Ptr{Cvoid}, Ref{PgnExplorerParams}, Ref{Cstring}),
                                h, ep, err), err)
                    out = Ref(round_out_struct(bufs[i]))
                    check(ccall((:pgn_run_round, libpigeons_b200[]), Cint, (Ptr{Cvoid}, Int64, Ref{PgnRoundOut}, Ref{Cstring}),
                                h, n_scans, out, err), err)
                    outs[i] = out[]
                end
            end
        end
    end
    pt.shared.iterators.scan = 0                      # next_scan! resets the scan counter when a round ends (Iterators.jl:37-47)
    reduced = materialise_recorders(pt, bufs, outs, n_scans)
    Pigeons.record_timed_if_requested!(reduced, :round, timed)
    return reduced
end

# --------------------------------------------------------------------------------------------------------------------
# pgn_round_out -> the `reduced_recorders` NamedTuple `adapt` / `report!` / `stepping_stone` expect
# --------------------------------------------------------------------------------------------------------------------
mean_stat(mu::Float64, n::Integer) = Mean(mu, EqualWeight(), Int(n))                    # OnlineStatsBase 1.x: Mean(μ, weight, n)
sum_stat(s::Integer) = Sum(Int(s), 1)                                                   # OnlineStatsBase 1.x: Sum(sum, n)
function variance_stat(mu::Float64, var_bessel::Float64, n::Integer)                    # OnlineStatsBase 1.x: Variance(σ2, μ, weight, n)
    s2 = n > 1 ? var_bessel * (n - 1) / n : 0.0                                         # the struct holds the biased σ2; value() applies Bessel
    return Variance(s2, mu, EqualWeight(), Int(n))
end

function groupby_from(keys_values, T::Type, init)
    g = GroupBy(T, init)
    for (k, stat) in keys_values
        g.value[k] = stat                                                               # OrderedDict, insertion order = chain order
        g.n += OnlineStatsBase.nobs(stat)
    end
    return g
end

"""
Fills a fresh `create_recorders(inputs, shared)` NamedTuple from the engine's fixed-layout arrays:

| recorder (src/recorders/recorder.jl)            | filled from                                            |
|---|---|
| `swap_acceptance_pr = GroupBy((i,i+1) -> Mean)` | `swap_n[i]`, `swap_mean[i]` (pair stored at its lower chain)   |
| `log_sum_ratio = GroupBy((i,j) -> LogSum)`      | `logsum_fwd[i]` -> key `(i,i+1)`, `logsum_bwd[i]` -> key `(i+1,i)` |
| `explorer_acceptance_pr`, `am_factors`, `reversibility_rate` = `GroupBy(chain -> Mean)` | `*_n`, `*_mean` |
| `explorer_n_steps = GroupBy(chain -> Sum)`      | `expl_n_steps`                                         |
| `round_trip::RoundTripRecorder`                 | `n_tempered_restarts`, `n_round_trips` (summed over handles) |
| `_transformed_online`, `online`                 | `online_n`, `online_mean`, `online_var` of the handle owning chain N |
| `index_process::Dict{Int,Vector{Int}}`          | column `r` = chains visited by replica `r`, inverted from the device's chain-major log |
| `traces::Dict{Pair{Int,Int},Any}`               | `target_trace[:, scan]` under key `N => scan`           |

Keys are inserted in chain order.  The reference inserts them in replica-merge order; `GroupBy` lookups
(`value_with_default`, `recorder_values` sorted by key in `adapt`) do not depend on insertion order.
"""
function materialise_recorders(pt, bufs::Vector{RoundBuffers}, outs::Vector{PgnRoundOut}, n_scans::Int)
    r = pt.replicas
    N = r.n_chains
    template = create_recorders(pt.inputs, pt.shared)
    cat(f) = reduce(vcat, (f(b) for b in bufs))
    swap_n, swap_mean = cat(b -> b.swap_n), cat(b -> b.swap_mean)
    ls_f, ls_b = cat(b -> b.logsum_fwd), cat(b -> b.logsum_bwd)
    acc_n, acc_mu, steps = cat(b -> b.expl_acc_n), cat(b -> b.expl_acc_mean), cat(b -> b.expl_n_steps)
    am_n, am_mu, rev_n, rev_mu = cat(b -> b.am_n), cat(b -> b.am_mean), cat(b -> b.rev_n), cat(b -> b.rev_mean)
    filled = Dict{Symbol, Any}()
    if haskey(template, :swap_acceptance_pr)
        filled[:swap_acceptance_pr] = groupby_from(((i, i + 1) => mean_stat(swap_mean[i], swap_n[i]) for i in 1:(N - 1) if swap_n[i] > 0),
                                                   Tuple{Int, Int}, Mean())
    end
    if haskey(template, :log_sum_ratio)
        kv = Pair{Tuple{Int, Int}, LogSum{Float64}}[]
        for i in 1:(N - 1)
            swap_n[i] > 0 || continue
            push!(kv, (i, i + 1) => LogSum(ls_f[i], Int(swap_n[i])))
            push!(kv, (i + 1, i) => LogSum(ls_b[i], Int(swap_n[i])))
        end
        filled[:log_sum_ratio] = groupby_from(kv, Tuple{Int, Int}, LogSum())
    end
    per_chain(n, mu) = (c => mean_stat(mu[c], n[c]) for c in 1:N if n[c] > 0)
    haskey(template, :explorer_acceptance_pr) && (filled[:explorer_acceptance_pr] = groupby_from(per_chain(acc_n, acc_mu), Int, Mean()))
    haskey(template, :am_factors) && (filled[:am_factors] = groupby_from(per_chain(am_n, am_mu), Int, Mean()))
    haskey(template, :reversibility_rate) && (filled[:reversibility_rate] = groupby_from(per_chain(rev_n, rev_mu), Int, Mean()))
    haskey(template, :explorer_n_steps) &&
        (filled[:explorer_n_steps] = groupby_from((c => sum_stat(steps[c]) for c in 1:N if steps[c] != 0), Int, Sum()))
    if haskey(template, :round_trip)
        rt = RoundTripRecorder()
        rt.n_tempered_restarts = sum(o.n_tempered_restarts for o in outs)
        rt.n_round_trips = sum(o.n_round_trips for o in outs)
        filled[:round_trip] = rt
    end
    # the handle owning the target chain(s): chain N, or with two legs chains n_var and n_var + 1 (one handle, engine rule)
    n_var = Pigeons.n_chains_var(pt.inputs)
    two_legs = 0 < n_var < N
    t_chain = two_legs ? n_var : N
    owner = findfirst(i -> r.first_chain[i] <= t_chain < r.first_chain[i] + r.n_local[i], eachindex(r.handles))
    last, olast = bufs[owner], outs[owner]
    for key in (:_transformed_online, :online)
        haskey(template, key) || continue
        rec = OnlineStateRecorder()
        if olast.online_n > 0 && r.dim > 0
            n = olast.online_n
            rec.stats[Pair(:singleton_variable, Mean)] = Group([mean_stat(last.online_mean[c], n) for c in 1:r.dim])
            rec.stats[Pair(:singleton_variable, Variance)] =
                Group([variance_stat(last.online_mean[c], last.online_var[c], n) for c in 1:r.dim])
        end
        filled[key] = rec
    end
    if haskey(template, :index_process)
        ip = Dict{Int, Vector{Int}}()
        chain_major = reduce(vcat, (b.index_process for b in bufs))          # [N, n_scans]: replica sitting at each chain
        for s in 1:n_scans, c in 1:N
            push!(get!(ip, Int(chain_major[c, s]), Int[]), c)                # recorder is keyed by replica, holds its chains
        end
        filled[:index_process] = ip
    end
    if haskey(template, :traces)
        tr = Dict{Pair{Int, Int}, Any}()
        for s in 1:n_scans
            if two_legs                                  # both target chains record (VariationalDEO.jl:20, pigeons.jl:110-131)
                tr[n_var => s] = last.target_trace[1:r.dim, 2 * s - 1]
                tr[(n_var + 1) => s] = last.target_trace[1:r.dim, 2 * s]
            else
                tr[N => s] = last.target_trace[1:r.dim, s]
            end
        end
        filled[:traces] = tr
    end
    ks = keys(template)
    return NamedTuple{ks}(Tuple(get(filled, k, template[k]) for k in ks))
end

# --------------------------------------------------------------------------------------------------------------------
# Replica objects for write_checkpoint / run_checks (src/pt/checkpoint.jl:110-145, src/pt/checks.jl:52-78)
# --------------------------------------------------------------------------------------------------------------------
"""
The device RNG is Philox4x32-10 keyed by (seed, replica_index) with a 64-bit draw counter; `Replica.rng` is typed
`SplittableRandom` (src/replicas/Replica.jl:19), so a checkpoint stores the counter in the `seed` field and the
replica index in `gamma`: enough for `pgn_set_state` to continue bit for bit, and for `check_against_serial` to
compare two device runs field by field.  It is NOT the stream of an unmodified CPU run (SURVEY.md §8 f2).
"""
function materialise_replicas(r::B200Replicas)
    out = Replica[]
    for (i, h) in enumerate(r.handles)
        n = r.n_local[i]
        x = zeros(Float64, max(r.dim, 1), n)                          # C [n_local][d] row-major == Julia d x n
        ri, ctr, rt = zeros(Int32, n), zeros(UInt64, n), zeros(Int32, n)
        st = PgnReplicaState(pointer(x), pointer(ri), pointer(ctr), pointer(rt))
        GC.@preserve x ri ctr rt begin
            err = Ref{Cstring}(C_NULL)
            check(ccall((:pgn_get_state, libpigeons_b200[]), Cint, (Ptr{Cvoid}, Ref{PgnReplicaState}, Ref{Cstring}), h, st, err), err)
        end
        for j in 1:n
            recorders = create_recorders(r.builders)                  # empty: the device reduces them every round
            rng = SplittableRandom(ctr[j], UInt64(ri[j]))
            push!(out, Replica(x[1:r.dim, j], r.first_chain[i] + j - 1, rng, recorders, Int(ri[j])))
        end
    end
    return out
end

"""Load replicas (e.g. deserialised `replica=i.jls` files, sorted by chain) back into the engine handles."""
function load_replicas!(r::B200Replicas, replicas::Vector{<:Replica})
    sorted = sort(replicas, by = rep -> rep.chain)
    for (i, h) in enumerate(r.handles)
        n, fc = r.n_local[i], r.first_chain[i]
        x = zeros(Float64, max(r.dim, 1), n)
        ri, ctr, rt = zeros(Int32, n), zeros(UInt64, n), zeros(Int32, n)
        for j in 1:n
            rep = sorted[fc + j - 1]
            r.dim > 0 && (x[1:r.dim, j] .= rep.state)
            ri[j] = rep.replica_index
            ctr[j] = rep.rng.seed
        end
        st = PgnReplicaState(pointer(x), pointer(ri), pointer(ctr), pointer(rt))
        GC.@preserve x ri ctr rt begin
            err = Ref{Cstring}(C_NULL)
            check(ccall((:pgn_set_state, libpigeons_b200[]), Cint, (Ptr{Cvoid}, Ref{PgnReplicaState}, Ref{Cstring}), h, st, err), err)
        end
    end
end

end # module
