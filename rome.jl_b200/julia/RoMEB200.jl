# RoMEB200.jl -- Julia binding of librome_b200.so (include/rome_b200.h) and the batched entry points a
# RoME/IncrementalInference maintainer would call instead of the per-particle CalcFactor loop.
#
# NOT EXECUTED in the build container (no `julia`; see DESIGN.md "what could not be verified").  The C ABI it
# binds IS tested: tests/ drives the identical entry points through ctypes (rome.jl_b200/_lib.py mirrors this
# file call for call).
#
# Reference interfaces replaced (paths relative to RoME.jl):
#   (cf::CalcFactor{<:Pose2Pose2})(X, p, q)                       src/factors/Pose2D.jl:40-67
#   (cf::CalcFactor{<:PriorPose2})(m, p)                          src/factors/PriorPose2.jl:27-47
#   (cf::CalcFactor{<:Pose2Point2BearingRange})(meas, p, l)       src/factors/BearingRange2D.jl:39-64
#   getSample(cf::CalcFactor{<:Pose2Point2BearingRange})          src/factors/BearingRange2D.jl:17-27
#   (cf::CalcFactor{<:Pose3Pose3})(X, p, q)                       src/factors/Pose3Pose3.jl:17-29
#   (cf::CalcFactor{<:PriorPose3})(m, p)                          src/factors/Pose3D.jl:15-19
module RoMEB200

using RoME, DistributedFactorGraphs, IncrementalInference

const LIB = get(ENV, "ROME_B200_LIB", joinpath(@__DIR__, "..", "librome_b200.so"))

# enums of include/rome_b200.h
const POSE2, POINT2, POSE3, POINT3, ROTATION3 = Cint(0), Cint(1), Cint(2), Cint(3), Cint(4)
const POSE2POSE2, PRIORPOSE2, BEARINGRANGE, POSE3POSE3, PRIORPOSE3 = Cint(0), Cint(1), Cint(2), Cint(3), Cint(4)
# next-row families (SURVEY.md 8f N1)
const PRIORPOINT2, POINT2POINT2, POSE2POINT2, POSE2POINT2RANGE, POINT2POINT2RANGE, POSE2POINT2BEARING =
    Cint(5), Cint(6), Cint(7), Cint(8), Cint(9), Cint(10)
const PRIORPOINT3, POINT3POINT3, POSE3POSE3XYYAW, POSE3POSE3ROTATION, POSE3POSE3UNITTRANS =
    Cint(11), Cint(12), Cint(13), Cint(14), Cint(15)
# families with a third variable (src/factors/Pose3Pose3.jl:57-95)
const POSE3POSE3ROTOFFSET, POSE3POSE3TRANSFORM = Cint(16), Cint(17)
const RESIDUAL, PROPOSAL_FWD, PROPOSAL_BWD, STATS, SAMPLE, WRITE_MEAS, JACOBIAN, INDEPENDENT, DECONV, PRECISE =
    UInt32(1), UInt32(2), UInt32(4), UInt32(8), UInt32(16), UInt32(32), UInt32(64), UInt32(128), UInt32(256), UInt32(512)
const ROUTED_ONLY, BARRIER_WAIT, BARRIER_SIGNAL = UInt32(1024), UInt32(2048), UInt32(4096)
const PRODUCT_REANCHOR = UInt32(1)
const PRODUCT_MANIFOLD = UInt32(2)   # Pose3: rotations multiplied in the tangent space at the anchor rotation

struct Buffers            # struct rome_b200_buffers
    meas::Ptr{Cfloat}
    meas_out::Ptr{Cfloat}
    res::Ptr{Cfloat}
    prop_fwd::Ptr{Cfloat}
    prop_bwd::Ptr{Cfloat}
    stats::Ptr{Cfloat}
    jac::Ptr{Cfloat}
end

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer=0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:rome_b200_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ref)
        rc == 0 || error("rome_b200_create: " * unsafe_string(ccall((:rome_b200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(ref[])
        finalizer(c -> ccall((:rome_b200_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), ctx)
        return ctx
    end
end

check(ctx::Context, rc) = rc == 0 ? nothing :
    error("rome_b200 error $rc: " * unsafe_string(ccall((:rome_b200_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h)))

npad(N) = cld(N, 8) * 8

# ---- variables --------------------------------------------------------------------------------------------
# coords: d x N x nvars Float64 (column-major == the C layout [nvars][N][d]); e.g. for Pose2
#   coords[:, n, v] = getCoordinates(Pose2, getVal(fg, labels[v])[n])
function set_particles!(ctx::Context, vartype::Cint, coords::Array{Float64,3})
    d, N, nvars = size(coords)
    check(ctx, ccall((:rome_b200_set_particles, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cdouble}),
                     ctx.h, vartype, nvars, N, coords))
end

function coordinates(fg::AbstractDFG, labels::Vector{Symbol}, ::Type{T}, N::Int) where {T<:InferenceVariable}
    d = getDimension(T)
    out = Array{Float64,3}(undef, d, N, length(labels))
    M = getManifold(T)
    for (v, l) in enumerate(labels), (n, p) in enumerate(getVal(fg, l)[1:N])
        out[:, n, v] .= vee(M, getPointIdentity(M), log(M, getPointIdentity(M), p))   # == getCoordinates(T, p)
    end
    return out
end

# ---- factors ------------------------------------------------------------------------------------------------
function set_factors!(ctx::Context, ::Type{Pose2Pose2}, ip::Vector{Int32}, iq::Vector{Int32}, Z::Vector{<:MvNormal})
    mu = reduce(hcat, mean.(Z)); cv = reduce(hcat, vec.(Matrix.(cov.(Z))))
    check(ctx, ccall((:rome_b200_set_factors_pose2pose2, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}),
                     ctx.h, length(ip), ip, iq, mu, cv))
end
function set_factors!(ctx::Context, ::Type{PriorPose2}, ip::Vector{Int32}, Z::Vector{<:MvNormal})
    mu = reduce(hcat, mean.(Z)); cv = reduce(hcat, vec.(Matrix.(cov.(Z))))
    check(ctx, ccall((:rome_b200_set_factors_priorpose2, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}), ctx.h, length(ip), ip, mu, cv))
end
function set_factors!(ctx::Context, ::Type{Pose2Point2BearingRange}, ip::Vector{Int32}, il::Vector{Int32},
                      fcts::Vector{<:Pose2Point2BearingRange})
    b = reduce(hcat, [[mean(f.bearing), std(f.bearing)] for f in fcts])
    r = reduce(hcat, [[mean(f.range), std(f.range)] for f in fcts])
    check(ctx, ccall((:rome_b200_set_factors_bearingrange, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}), ctx.h, length(ip), ip, il, b, r))
end
function set_factors!(ctx::Context, ::Type{Pose3Pose3}, ip::Vector{Int32}, iq::Vector{Int32}, Z::Vector{<:MvNormal})
    mu = reduce(hcat, mean.(Z)); cv = reduce(hcat, vec.(Matrix.(cov.(Z))))
    check(ctx, ccall((:rome_b200_set_factors_pose3pose3, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}),
                     ctx.h, length(ip), ip, iq, mu, cv))
end

# every family whose belief is a single MvNormal (Point2Point2, Point3Point3, Pose3Pose3XYYaw, ...): iq = nothing for priors
function set_factors!(ctx::Context, family::Cint, ip::Vector{Int32}, iq::Union{Nothing,Vector{Int32}}, Z::Vector{<:MvNormal})
    mu = reduce(hcat, mean.(Z)); cv = reduce(hcat, vec.(Matrix.(cov.(Z))))
    check(ctx, ccall((:rome_b200_set_factors_gaussian, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}),
                     ctx.h, family, length(ip), ip, iq === nothing ? C_NULL : iq, mu, cv))
end

# families with a third variable: Pose3Pose3RotOffset (ir: Rotation3 variables), Pose3Pose3Transform (ir: Pose3 variables)
function set_factors!(ctx::Context, family::Cint, ip::Vector{Int32}, iq::Vector{Int32}, ir::Vector{Int32}, Z::Vector{<:MvNormal})
    mu = reduce(hcat, mean.(Z)); cv = reduce(hcat, vec.(Matrix.(cov.(Z))))
    check(ctx, ccall((:rome_b200_set_factors_ternary, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}),
                     ctx.h, family, length(ip), ip, iq, ir, mu, cv))
end

# ---- belief update on the device (SURVEY.md 8f N2): product of the proposal densities of every variable ---------
# plan: CSR over variables of `vartype`; source = (index into `bufs`, proposal row = factor index)
function set_product_plan!(ctx::Context, vartype::Cint, var_offsets::Vector{Int32}, src_buf::Vector{Int32}, src_row::Vector{Int32})
    check(ctx, ccall((:rome_b200_set_product_plan, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                     ctx.h, vartype, length(var_offsets) - 1, var_offsets, src_buf, src_row))
end
# bufs: DEVICE pointers of the proposal buffers (rome_b200_malloc_device); replaces manifoldProduct / manikde! on the host
function product!(ctx::Context, vartype::Cint, bufs::Vector{Ptr{Cfloat}}; seed=UInt64(0), sweep=UInt32(0), gibbs_iters=0)
    check(ctx, ccall((:rome_b200_product, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Cint, Ptr{Ptr{Cfloat}}, UInt64, UInt32, Cint, UInt32, Ptr{Cfloat}),
                     ctx.h, vartype, length(bufs), bufs, seed, sweep, gibbs_iters, PRODUCT_REANCHOR | PRODUCT_MANIFOLD, C_NULL))
end
function get_particles(ctx::Context, vartype::Cint, d::Int, N::Int, nvars::Int)
    out = Array{Float64,3}(undef, d, N, nvars)
    check(ctx, ccall((:rome_b200_get_particles, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), ctx.h, vartype, out))
    return out
end

# ---- the hot path --------------------------------------------------------------------------------------------
# Host-buffer evaluation of every factor of `family`: returns (res, prop_fwd, stats) as Float32 arrays in the
# library's particle-major layout (d x Npad x nF column-major).  `meas === nothing` draws the measurement in-kernel.
function eval_host(ctx::Context, family::Cint, nF::Int, N::Int, dm::Int, dr::Int, dfwd::Int, nstats::Int;
                   meas::Union{Nothing,Array{Float32,3}}=nothing, seed::UInt64=UInt64(0), stream_id::UInt32=UInt32(0),
                   proposals::Bool=true)
    Np = npad(N)
    res = zeros(Float32, dr, Np, nF); stats = zeros(Float32, nstats, nF)
    prop = proposals ? zeros(Float32, dfwd, Np, nF) : zeros(Float32, 0, 0, 0)
    flags = RESIDUAL | STATS | (proposals ? PROPOSAL_FWD : UInt32(0)) | (meas === nothing ? SAMPLE : UInt32(0))
    GC.@preserve meas res prop stats begin
        b = Ref(Buffers(meas === nothing ? C_NULL : pointer(meas), C_NULL, pointer(res),
                        proposals ? pointer(prop) : C_NULL, C_NULL, pointer(stats), C_NULL))
        check(ctx, ccall((:rome_b200_eval_host, LIB), Cint,
                         (Ptr{Cvoid}, Cint, UInt32, UInt64, UInt32, Cint, Cint, Ref{Buffers}),
                         ctx.h, family, flags, seed, stream_id, 0, -1, b))
    end
    return res, prop, stats
end

# ---- drop-in: batched replacement of `approxConvBelief` for the five hot families ------------------------------
# [IIF-knowledge, unverified here] IIF reaches the factor functor through
#   approxConvBelief -> evalFactor -> evalPotentialSpecific -> _solveCCWNumeric! (per particle, Optim NelderMead).
# The replacement works one level up: ONE library call per (factor type, sweep) evaluates getSample + the closed-form
# root for every factor of the type x every particle; the N proposal points of a factor are what approxConvBelief returns.
#
# `DeviceGraph` keeps the context, the variable numbering and the factor tables ACROSS calls: tables are uploaded when the
# graph is mirrored (or a factor is added), particles only when they changed on the host (`upload_particles!`).

const FAMILY_OF = Dict{DataType,Cint}(Pose2Pose2 => POSE2POSE2, PriorPose2 => PRIORPOSE2,
                                      Pose2Point2BearingRange => BEARINGRANGE, Pose3Pose3 => POSE3POSE3,
                                      PriorPose3 => PRIORPOSE3)
# family -> (first variable type, last variable type or nothing, dm, dr, d of the forward proposal, nstats)
const FAMILY_DIMS = Dict{Cint,Tuple}(POSE2POSE2 => (Pose2, Pose2, 3, 3, 3, 16), PRIORPOSE2 => (Pose2, nothing, 3, 3, 3, 16),
                                     BEARINGRANGE => (Pose2, Point2, 2, 2, 2, 16), POSE3POSE3 => (Pose3, Pose3, 6, 6, 6, 32),
                                     PRIORPOSE3 => (Pose3, nothing, 6, 6, 6, 32))
const VARTYPE_OF = Dict{DataType,Cint}(Pose2 => POSE2, Point2 => POINT2, Pose3 => POSE3)

mutable struct DeviceGraph
    ctx::Context
    fg::AbstractDFG
    N::Int
    vlabels::Dict{DataType,Vector{Symbol}}          # variable type -> labels in device order
    vindex::Dict{Symbol,Int32}                      # label -> index inside its type's particle store
    flabels::Dict{Cint,Vector{Symbol}}              # family -> factor labels in table order
    sweep::UInt32
end

function DeviceGraph(fg::AbstractDFG; device::Integer=0, N::Int=getSolverParams(fg).N)
    ctx = Context(device)
    vlabels = Dict{DataType,Vector{Symbol}}(T => sortDFG(ls(fg, T)) for T in (Pose2, Point2, Pose3))
    vindex = Dict{Symbol,Int32}()
    for (T, ls_) in vlabels, (i, l) in enumerate(ls_)
        vindex[l] = Int32(i - 1)
    end
    dg = DeviceGraph(ctx, fg, N, vlabels, vindex, Dict{Cint,Vector{Symbol}}(), UInt32(0))
    upload_factors!(dg)
    upload_particles!(dg)
    return dg
end

"upload the particles of every variable type (call after the host changed them)"
function upload_particles!(dg::DeviceGraph)
    for (T, labels) in dg.vlabels
        isempty(labels) && continue
        set_particles!(dg.ctx, VARTYPE_OF[T], coordinates(dg.fg, labels, T, dg.N))
    end
end

"(re)build the factor tables of the five hot families from the graph"
function upload_factors!(dg::DeviceGraph)
    fg = dg.fg
    for (FT, fam) in FAMILY_OF
        labels = Symbol[l for l in lsf(fg) if getFactorType(fg, l) isa FT]
        dg.flabels[fam] = labels
        isempty(labels) && continue
        i0 = Int32[dg.vindex[getVariableOrder(fg, l)[1]] for l in labels]
        fcts = [getFactorType(fg, l) for l in labels]
        if FT === Pose2Pose2 || FT === Pose3Pose3
            i1 = Int32[dg.vindex[getVariableOrder(fg, l)[2]] for l in labels]
            set_factors!(dg.ctx, FT, i0, i1, [f.Z for f in fcts])
        elseif FT === Pose2Point2BearingRange
            i1 = Int32[dg.vindex[getVariableOrder(fg, l)[2]] for l in labels]
            set_factors!(dg.ctx, FT, i0, i1, fcts)
        elseif FT === PriorPose2
            set_factors!(dg.ctx, FT, i0, [f.Z for f in fcts])
        else  # PriorPose3
            set_factors!(dg.ctx, PRIORPOSE3, i0, nothing, [f.Z for f in fcts])
        end
    end
end

"""
    approxConvBatch(dg, F) -> (proposals::Dict{Symbol,Vector}, res, stats)

Convolve EVERY factor of type `F` (Pose2Pose2, PriorPose2, Pose2Point2BearingRange, Pose3Pose3, PriorPose3) toward its
last variable (the prior's variable): fused getSample + closed-form root on the device, N proposal points per factor.
`proposals[factor label]` is what `approxConvBelief(fg, factor, target)` returns for that factor.
"""
function approxConvBatch(dg::DeviceGraph, ::Type{F}; seed::UInt64=UInt64(0)) where {F}
    fam = FAMILY_OF[F]
    T0, T1, dm, dr, dfwd, nstats = FAMILY_DIMS[fam]
    labels = dg.flabels[fam]
    out = Dict{Symbol,Vector}()
    isempty(labels) && return out, zeros(Float32, dr, npad(dg.N), 0), zeros(Float32, nstats, 0)
    res, prop, stats = eval_host(dg.ctx, fam, length(labels), dg.N, dm, dr, dfwd, nstats; seed=seed, stream_id=dg.sweep)
    dg.sweep += UInt32(1)
    TT = T1 === nothing ? T0 : T1                      # type of the target variable
    M = getManifold(TT)
    e = getPointIdentity(M)
    for (k, l) in enumerate(labels)
        tgt = getVariableOrder(dg.fg, l)[end]
        # proposals are offsets from the target variable's anchor == the coordinates of its first particle
        a = vee(M, e, log(M, e, getVal(dg.fg, tgt)[1]))
        out[l] = [getPoint(TT, a .+ Float64.(prop[:, n, k])) for n in 1:dg.N]
    end
    return out, res, stats
end

# one-shot convenience (mirrors the graph, convolves, drops the mirror) -- the cached form above is the one a solver uses
approxConvBatch(fg::AbstractDFG, ::Type{F}; device::Integer=0, N::Int=getSolverParams(fg).N, seed::UInt64=UInt64(0)) where {F} =
    approxConvBatch(DeviceGraph(fg; device=device, N=N), F; seed=seed)

# ---- the method a maintainer adds to IIF / RoME so that canonical graphs drop in unchanged ---------------------------
# [IIF-knowledge, unverified here: IncrementalInference 0.35 is not vendored]  `approxConvBelief(dfg, fc, target, ...)`
# is IIF's per-factor convolution; specialising it on the factor TYPE keeps `generateGraph_Hexagonal`, the g2o loader,
# `solveTree!` etc. untouched while every convolution of a hot family is served from a per-graph cache that is refreshed
# once per (family, sweep):
#
#   const _MIRROR = IdDict{AbstractDFG,DeviceGraph}()                       # one device mirror per graph
#   const _CACHE  = Dict{Tuple{UInt,Cint,UInt32},Dict{Symbol,Vector}}()     # (graph id, family, sweep) -> proposals
#
#   function IncrementalInference.approxConvBelief(dfg::AbstractDFG, fc::DFGFactor{<:CommonConvWrapper{<:F}},
#                                                  target::Symbol, args...; kw...) where
#            {F<:Union{Pose2Pose2,PriorPose2,Pose2Point2BearingRange,Pose3Pose3,PriorPose3}}
#       getVariableOrder(fc)[end] == target || return invoke(approxConvBelief, Tuple{AbstractDFG,DFGFactor,Symbol}, dfg, fc, target, args...; kw...)
#       dg = get!(() -> RoMEB200.DeviceGraph(dfg), _MIRROR, dfg)
#       key = (objectid(dfg), RoMEB200.FAMILY_OF[F], dg.sweep)
#       props = get!(_CACHE, key) do
#           RoMEB200.upload_particles!(dg)                                   # particles changed since the last sweep
#           first(RoMEB200.approxConvBatch(dg, F))
#       end
#       return manikde!(getManifold(getVariable(dfg, target)), props[getLabel(fc)])
#   end
#
# Backward convolutions (toward the first variable) take the same route with ROME_B200_PROPOSAL_BWD; the numeric
# (Nelder-Mead) path remains for every other factor type and for multihypo / nullhypo factors.

end # module
