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
const POSE2, POINT2, POSE3, POINT3 = Cint(0), Cint(1), Cint(2), Cint(3)
const POSE2POSE2, PRIORPOSE2, BEARINGRANGE, POSE3POSE3, PRIORPOSE3 = Cint(0), Cint(1), Cint(2), Cint(3), Cint(4)
# next-row families (SURVEY.md 8f N1)
const PRIORPOINT2, POINT2POINT2, POSE2POINT2, POSE2POINT2RANGE, POINT2POINT2RANGE, POSE2POINT2BEARING =
    Cint(5), Cint(6), Cint(7), Cint(8), Cint(9), Cint(10)
const PRIORPOINT3, POINT3POINT3, POSE3POSE3XYYAW, POSE3POSE3ROTATION, POSE3POSE3UNITTRANS =
    Cint(11), Cint(12), Cint(13), Cint(14), Cint(15)
const RESIDUAL, PROPOSAL_FWD, PROPOSAL_BWD, STATS, SAMPLE, WRITE_MEAS, JACOBIAN, INDEPENDENT, DECONV, PRECISE =
    UInt32(1), UInt32(2), UInt32(4), UInt32(8), UInt32(16), UInt32(32), UInt32(64), UInt32(128), UInt32(256), UInt32(512)
const ROUTED_ONLY, BARRIER_WAIT, BARRIER_SIGNAL = UInt32(1024), UInt32(2048), UInt32(4096)
const PRODUCT_REANCHOR = UInt32(1)

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
                     ctx.h, vartype, length(bufs), bufs, seed, sweep, gibbs_iters, PRODUCT_REANCHOR, C_NULL))
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

# ---- drop-in: batched replacement of `approxConvBelief` for all Pose2Pose2 factors of a graph ----------------
# [IIF-knowledge, unverified here] IIF reaches the factor functor through
#   approxConvBelief -> evalFactor -> evalPotentialSpecific -> _solveCCWNumeric! (per particle, Optim NelderMead).
# A maintainer overrides at `approxConvBelief` level for the factor types below: gather every factor of the type,
# one library call per (type, sweep), then hand the N proposal points per factor back as the convolution result.
function approxConvBatch(ctx::Context, fg::AbstractDFG, ::Type{Pose2Pose2}; N::Int=getSolverParams(fg).N, seed=UInt64(0))
    flabels = [l for l in lsf(fg) if getFactorType(fg, l) isa Pose2Pose2]
    vlabels = ls(fg, Pose2) |> sortDFG
    vidx = Dict(l => Int32(i - 1) for (i, l) in enumerate(vlabels))
    ip = Int32[vidx[getVariableOrder(fg, l)[1]] for l in flabels]
    iq = Int32[vidx[getVariableOrder(fg, l)[2]] for l in flabels]
    set_particles!(ctx, POSE2, coordinates(fg, vlabels, Pose2, N))
    set_factors!(ctx, Pose2Pose2, ip, iq, [getFactorType(fg, l).Z for l in flabels])
    res, prop, stats = eval_host(ctx, POSE2POSE2, length(flabels), N, 3, 3, 3, 16; seed=seed)
    # proposals are offsets from the target variable's anchor == its first particle's coordinates
    out = Dict{Symbol,Vector}()
    M = getManifold(Pose2)
    for (k, l) in enumerate(flabels)
        tgt = getVariableOrder(fg, l)[2]
        a = vee(M, getPointIdentity(M), log(M, getPointIdentity(M), getVal(fg, tgt)[1]))
        out[l] = [getPoint(Pose2, a .+ Float64.(prop[:, n, k])) for n in 1:N]
    end
    return out, res, stats
end

end # module
