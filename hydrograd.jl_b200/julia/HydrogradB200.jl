# HydrogradB200.jl -- thin Julia shim over libhydrograd_b200.so (include/hydrograd_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: neither the build container nor the GPU boxes have a `julia`
# binary (SURVEY.md, environment table).  The file is the reference-side binding a Hydrograd.jl
# maintainer would add; the same ABI is exercised end to end from Python (hydrograd.jl_b200/api.py,
# tests/test_gpu_parity.py).  Everything below only flattens Hydrograd.jl's own structs and forwards
# pointers -- no arithmetic happens on the Julia side.
#
# Drop-in usage (src/applications/solve_swe_2D.jl:281-307): replace the two closures by
#
#     b200 = HydrogradB200.Context(swe_extra_params)           # once, after SWE2D_Extra_Parameters is built
#     ode_f = ODEFunction((u, p, t)      -> HydrogradB200.swe_2d_rhs(similar(u), u, p, t, b200);  jac_prototype = nothing)
#     ode_f = ODEFunction((du, u, p, t)  -> HydrogradB200.swe_2d_rhs(du, u, p, t, b200))        # bInPlaceODE
#
# `swe_2d_rhs` keeps the reference signature (semi_discretize_swe_2D.jl:18-19) with the context in
# place of p_extra.  The ChainRulesCore.rrule below makes Zygote / ZygoteVJP-based SciMLSensitivity
# adjoints use the hand-written CUDA VJP (hg_rhs_vjp) instead of differentiating through the RHS.
# Forward-mode callers (ForwardDiff.Dual state and / or parameters: ForwardDiff.jacobian around the solve in
# swe_2D_sensitivity.jl:80, the ForwardDiffSensitivity / ForwardSensitivity inversion options) are served by the
# Dual methods of `swe_2d_rhs` below: values and partials are split, every partial goes through the device's
# forward mode (hg_rhs_jvp_multi: one upload of the state, all partials of a chunk in one launch) and the Duals are
# reassembled.  A default context runs the fused forward-mode tile kernel; with `strict=true` the calls run on the plain
# tables in the reference's evaluation order.  The rrule's pullback reuses the state its forward call left on the device
# (hg_rhs_vjp with Q = NULL while hg_state_generation is unchanged), so an adjoint stage uploads only the cotangent.
module HydrogradB200

using ChainRulesCore
import ForwardDiff
import ComponentArrays                      # only set_ude_model needs it (offsets of the Lux parameter arrays)

const LIB = get(ENV, "HYDROGRAD_B200_LIB", joinpath(@__DIR__, "..", "libhydrograd_b200.so"))

const HG_PARAM = Dict("" => Int32(0), "zb" => Int32(1), "ManningN" => Int32(2), "Q" => Int32(3), "UDE" => Int32(4))

# ---- mirrors of the C structs (field order and types must match hydrograd_b200.h) ----------------
struct MeshDesc
    n_cells::Int64; n_faces::Int64; n_ghost::Int64; ld::Int64
    index_base::Int32
    cell_nfaces::Ptr{Int64}; cell_faces::Ptr{Int64}; cell_neighbors::Ptr{Int64}
    cell_normals::Ptr{Float64}; face_is_boundary::Ptr{UInt8}
    face_lengths::Ptr{Float64}; cell_areas::Ptr{Float64}; cell_centroids::Ptr{Float64}
end
struct BcDesc
    n_inletq::Int64; n_exith::Int64; n_wall::Int64; n_symm::Int64
    bc_ptr::Ptr{Int64}; ghost_ids::Ptr{Int64}; internal_cells::Ptr{Int64}
    outward_normals::Ptr{Float64}; face_lengths::Ptr{Float64}
    n_halo::Int64; halo_flip::Ptr{UInt8}; halo_area::Ptr{Float64}      # multi-GPU extension: 0 / NULL for one GPU
end
struct FieldsDesc
    g::Float64; k_n::Float64; h_small::Float64
    riemann_solver::Cstring
    hstill::Ptr{Float64}; hstill_ghost::Ptr{Float64}; zb_cells::Ptr{Float64}; zb_ghost::Ptr{Float64}
    S0_cells::Ptr{Float64}; ManningN_cells::Ptr{Float64}; matID_cells::Ptr{Int64}; n_mat::Int64
    inletQ_TotalQ::Ptr{Float64}; exitH_WSE::Ptr{Float64}
end
struct Options
    device::Int32; tile_cells::Int32; reorder::Int32; strict::Int32; path::Int32
    reserved::NTuple{11,Int32}
end

mutable struct Context
    handle::Ptr{Cvoid}
    N::Int
    active::Int32
    function Context(h::Ptr{Cvoid}, N::Int, active::Int32)
        ctx = new(h, N, active)
        finalizer(c -> (c.handle != C_NULL && ccall((:hg_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.handle); c.handle = C_NULL), ctx)
        return ctx
    end
    function Context(p_extra; device::Integer=0, tile_cells::Integer=256, strict::Bool=false)
        h = _create(p_extra, Int32(device), Int32(tile_cells), strict)
        ctx = new(h, p_extra.my_mesh_2D.numOfCells, HG_PARAM[p_extra.active_param_name])
        finalizer(c -> (c.handle != C_NULL && ccall((:hg_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.handle); c.handle = C_NULL), ctx)
        return ctx
    end
end

# a context around an already created handle (multi-GPU rank contexts)
_wrap(h::Ptr{Cvoid}, N::Int, active::Int32) = Context(h, N, active)

_err(h) = unsafe_string(ccall((:hg_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
_check(rc, h) = rc == 0 ? nothing : error("hydrograd_b200 error $rc: " * _err(h))

# Flatten mesh_2D / BoundaryConditions2D / SWE2D_Extra_Parameters (no copies of the big matrices that
# are already dense: cellFacesList, cellNodesCount, face_lengths, cell_areas, cell_centroids, S0_cells).
function _create(px, device::Int32, tile::Int32, strict::Bool)
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    _with_descs(px) do mesh, bcd, fld
        opt = Ref(Options(device, tile, Int32(1), Int32(strict), Int32(0), ntuple(_ -> Int32(0), 11)))
        rc = ccall((:hg_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{MeshDesc}, Ref{BcDesc}, Ref{FieldsDesc}, Ref{Options}),
                   handle, Ref(mesh), Ref(bcd), Ref(fld), opt)
        _check(rc, C_NULL)
    end
    return handle[]
end

# f(mesh::MeshDesc, bc::BcDesc, fields::FieldsDesc) is called while every array the descriptors point at is kept alive
function _with_descs(f, px)
    m  = px.my_mesh_2D
    bc = px.boundary_conditions
    N, ld = m.numOfCells, size(m.cellFacesList, 2)
    neigh   = zeros(Int64, N, ld)                       # cellNeighbors_Dict -> N x ld
    normals = zeros(Float64, N, ld, 2)                  # cell_normals       -> N x ld x 2
    for i in 1:N, j in 1:m.cellNodesCount[i]
        neigh[i, j] = m.cellNeighbors_Dict[i][j]
        normals[i, j, 1] = m.cell_normals[i][j][1]
        normals[i, j, 2] = m.cell_normals[i][j][2]
    end
    isb = UInt8.(m.bFace_is_boundary)
    # boundary entries in the reference's own processing order (bc_2D.jl:279-295)
    ptr = Int64[0]; gh = Int64[]; ic = Int64[]; nx = Float64[]; ny = Float64[]; len = Float64[]
    for k in 1:bc.nInletQ_BCs
        append!(gh, bc.inletQ_ghostCellIDs[k]); append!(ic, bc.inletQ_internalCellIDs[k])
        append!(nx, bc.inletQ_faceOutwardNormals[k][:, 1]); append!(ny, bc.inletQ_faceOutwardNormals[k][:, 2])
        append!(len, px.inletQ_Length[k]); push!(ptr, length(gh))
    end
    for k in 1:bc.nExitH_BCs
        append!(gh, bc.exitH_ghostCellIDs[k]); append!(ic, bc.exitH_internalCellIDs[k])
        append!(nx, bc.exitH_faceOutwardNormals[k][:, 1]); append!(ny, bc.exitH_faceOutwardNormals[k][:, 2])
        append!(len, zeros(length(bc.exitH_ghostCellIDs[k]))); push!(ptr, length(gh))
    end
    for k in 1:bc.nWall_BCs
        append!(gh, bc.wall_ghostCellIDs[k]); append!(ic, bc.wall_internalCellIDs[k])
        append!(nx, bc.wall_outwardNormals[k][:, 1]); append!(ny, bc.wall_outwardNormals[k][:, 2])
        append!(len, zeros(length(bc.wall_ghostCellIDs[k]))); push!(ptr, length(gh))
    end
    for k in 1:bc.nSymm_BCs
        append!(gh, bc.symm_ghostCellIDs[k]); append!(ic, bc.symm_internalCellIDs[k])
        append!(nx, bc.symm_outwardNormals[k][:, 1]); append!(ny, bc.symm_outwardNormals[k][:, 2])
        append!(len, zeros(length(bc.symm_ghostCellIDs[k]))); push!(ptr, length(gh))
    end
    bnorm = vcat(nx, ny)
    cellfaces = Matrix{Int64}(m.cellFacesList); nfaces = Vector{Int64}(m.cellNodesCount)
    flen = Vector{Float64}(m.face_lengths); areas = Vector{Float64}(m.cell_areas)
    cent = Matrix{Float64}(m.cell_centroids); S0 = Matrix{Float64}(px.S0_cells)
    matid = Vector{Int64}(px.srh_all_Dict["matID_cells"])
    nmat = length(px.srh_all_Dict["srhhydro_ManningsN"])
    c = px.swe_2D_constants
    solver = c.RiemannSolver
    GC.@preserve neigh normals isb ptr gh ic bnorm len cellfaces nfaces flen areas cent S0 matid solver px begin
        mesh = MeshDesc(N, m.numOfFaces, m.numOfAllBounaryFaces, ld, Int32(1), pointer(nfaces), pointer(cellfaces),
                        pointer(neigh), pointer(normals), pointer(isb), pointer(flen), pointer(areas), pointer(cent))
        bcd = BcDesc(bc.nInletQ_BCs, bc.nExitH_BCs, bc.nWall_BCs, bc.nSymm_BCs, pointer(ptr), pointer(gh), pointer(ic),
                     pointer(bnorm), pointer(len), 0, Ptr{UInt8}(C_NULL), Ptr{Float64}(C_NULL))
        fld = FieldsDesc(c.g, c.k_n, c.h_small, Base.unsafe_convert(Cstring, solver),
                         pointer(px.hstill), pointer(px.hstill_ghostCells), pointer(px.zb_cells), pointer(px.zb_ghostCells),
                         pointer(S0), pointer(px.ManningN_cells), pointer(matid), nmat,
                         pointer(px.inletQ_TotalQ), pointer(px.exitH_WSE))
        return f(mesh, bcd, fld)
    end
end

# ---------------------------------------------------------------- multi-GPU: one host process, one context per device
# Partition (recursive coordinate bisection, every inlet-q boundary kept on one rank), rank-local meshes with halo
# boundaries, one context per device, contexts connected through the library's own NVLink transport (hg_comm_*): after
# that every resident call (set_state / step_euler / rhs_resident ...) on the rank contexts exchanges its halo on the device.
# No counterpart in the reference (a single serial process); nothing here needs Python or NCCL.
mutable struct MultiContext
    ctxs::Vector{Context}
    own::Vector{Vector{Int64}}          # 1-based global ids of every rank's cells, in the rank's local order
    neighbors::Vector{Vector{Int64}}    # ranks (0-based) behind the halo boundaries of every rank, in order
    counts::Vector{Vector{Int64}}       # entries per halo boundary
end

function _case_array(cs::Ptr{Cvoid}, name::String)
    ptr = Ref{Ptr{Cvoid}}(C_NULL); cnt = Ref{Int64}(0); dt = Ref{Int32}(0)
    rc = ccall((:hg_case_array, LIB), Cint, (Ptr{Cvoid}, Cstring, Ref{Ptr{Cvoid}}, Ref{Int64}, Ref{Int32}), cs, name, ptr, cnt, dt)
    rc == 0 || return nothing
    T = dt[] == 0 ? Float64 : dt[] == 1 ? Int64 : UInt8
    return cnt[] == 0 ? T[] : copy(unsafe_wrap(Array, Ptr{T}(ptr[]), cnt[]))
end

function MultiContext(p_extra; device_ids::Vector{<:Integer}, tile_cells::Integer=256)
    P = length(device_ids)
    m, bc = p_extra.my_mesh_2D, p_extra.boundary_conditions
    N = m.numOfCells
    cx, cy = Vector{Float64}(m.cell_centroids[:, 1]), Vector{Float64}(m.cell_centroids[:, 2])
    groups = [unique(Int64.(bc.inletQ_internalCellIDs[k]) .- 1) for k in 1:bc.nInletQ_BCs]     # 0-based cell ids
    gptr = Int64[0; cumsum(length.(groups))]; gcells = isempty(groups) ? Int64[0] : vcat(groups...)
    part = Vector{Int32}(undef, N)
    _check(ccall((:hg_partition_rcb, LIB), Cint, (Int64, Ptr{Float64}, Ptr{Float64}, Int32, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}),
                 N, cx, cy, Int32(P), length(groups), gptr, gcells, part), C_NULL)
    ctxs = Context[]; own = Vector{Int64}[]; nbs = Vector{Int64}[]; cnts = Vector{Int64}[]
    _with_descs(p_extra) do mesh, bcd, fld
        for r in 0:P-1
            cs = Ref{Ptr{Cvoid}}(C_NULL); err = zeros(UInt8, 512)
            rc = ccall((:hg_partition_extract, LIB), Cint,
                       (Ref{Ptr{Cvoid}}, Ref{MeshDesc}, Ref{BcDesc}, Ref{FieldsDesc}, Ptr{Int32}, Int32, Ptr{Int64}, Ptr{UInt8}, Int64),
                       cs, Ref(mesh), Ref(bcd), Ref(fld), part, Int32(r), C_NULL, err, 512)
            rc == 0 || error("hg_partition_extract: " * unsafe_string(pointer(err)))
            A = Dict(k => _case_array(cs[], k) for k in ("cell_nfaces", "cell_faces", "cell_neighbors", "cell_normals", "face_is_boundary",
                     "face_lengths", "cell_areas", "cell_centroids", "bc_ptr", "bc_ghost_ids", "bc_internal_cells", "bc_normals", "bc_lengths",
                     "halo_flip", "halo_area", "hstill", "hstill_ghost", "zb_cells", "zb_ghost", "S0_cells", "ManningN_cells", "matID_cells",
                     "inletQ_TotalQ", "exitH_WSE", "own", "neighbors", "counts"))
            dims = zeros(Int64, 16)
            ccall((:hg_case_dims, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), cs[], dims)
            ccall((:hg_case_free, LIB), Cvoid, (Ptr{Cvoid},), cs[])
            solver = p_extra.swe_2D_constants.RiemannSolver
            h = Ref{Ptr{Cvoid}}(C_NULL)
            GC.@preserve A solver begin
                pp(k, T) = (A[k] === nothing || isempty(A[k])) ? Ptr{T}(C_NULL) : pointer(A[k])
                lm = MeshDesc(dims[1], dims[2], dims[3], dims[4], Int32(0), pp("cell_nfaces", Int64), pp("cell_faces", Int64),
                              pp("cell_neighbors", Int64), pp("cell_normals", Float64), pp("face_is_boundary", UInt8),
                              pp("face_lengths", Float64), pp("cell_areas", Float64), pp("cell_centroids", Float64))
                lb = BcDesc(dims[6], dims[7], dims[8], dims[9], pp("bc_ptr", Int64), pp("bc_ghost_ids", Int64), pp("bc_internal_cells", Int64),
                            pp("bc_normals", Float64), pp("bc_lengths", Float64), dims[11], pp("halo_flip", UInt8), pp("halo_area", Float64))
                lf = FieldsDesc(fld.g, fld.k_n, fld.h_small, Base.unsafe_convert(Cstring, solver), pp("hstill", Float64),
                                pp("hstill_ghost", Float64), pp("zb_cells", Float64), pp("zb_ghost", Float64), pp("S0_cells", Float64),
                                pp("ManningN_cells", Float64), pp("matID_cells", Int64), dims[10], pp("inletQ_TotalQ", Float64), pp("exitH_WSE", Float64))
                opt = Ref(Options(Int32(device_ids[r + 1]), Int32(tile_cells), Int32(1), Int32(0), Int32(0), ntuple(_ -> Int32(0), 11)))
                _check(ccall((:hg_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{MeshDesc}, Ref{BcDesc}, Ref{FieldsDesc}, Ref{Options}),
                             h, Ref(lm), Ref(lb), Ref(lf), opt), C_NULL)
            end
            push!(ctxs, _wrap(h[], Int(dims[1]), HG_PARAM[p_extra.active_param_name]))
            push!(own, A["own"] .+ 1); push!(nbs, A["neighbors"]); push!(cnts, A["counts"])
        end
    end
    # handles of every rank, then each rank connects to its neighbours: where its block lands in the neighbour's receive
    # buffer (entries before it in the neighbour's list) and which of the neighbour's flag lines is its own
    handles = [(b = zeros(UInt8, 128); isempty(nbs[r]) || _check(ccall((:hg_comm_export, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), ctxs[r].handle, b), ctxs[r].handle); b) for r in 1:P]
    for r in 1:P
        isempty(nbs[r]) && continue
        blob = vcat((handles[q + 1] for q in nbs[r])...)
        off = Int64[]; idx = Int64[]
        for q in nbs[r]
            j = findfirst(==(r - 1), nbs[q + 1])
            push!(off, sum(cnts[q + 1][1:j-1]; init = 0)); push!(idx, j - 1)
        end
        _check(ccall((:hg_comm_connect, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{UInt8}, Ptr{Int64}, Ptr{Int64}),
                     ctxs[r].handle, length(nbs[r]), blob, off, idx), ctxs[r].handle)
    end
    return MultiContext(ctxs, own, nbs, cnts)
end

# state in / out of the rank contexts (global vectors of length 3N, the reference's layout)
function set_state(mc::MultiContext, Q::Vector{Float64})
    N = length(Q) ÷ 3
    for (c, o) in zip(mc.ctxs, mc.own)
        q = vcat(Q[o], Q[N .+ o], Q[2N .+ o])
        _check(ccall((:hg_set_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.handle, q), c.handle)
    end
end
function step_euler(mc::MultiContext, dt::Float64, nsteps::Integer)      # launches are asynchronous: the ranks run side by side
    for c in mc.ctxs
        _check(ccall((:hg_step_euler, LIB), Cint, (Ptr{Cvoid}, Float64, Int64), c.handle, dt, Int64(nsteps)), c.handle)
    end
end
function get_state(mc::MultiContext, N::Integer)
    Q = zeros(3N)
    for (c, o) in zip(mc.ctxs, mc.own)
        q = zeros(3 * c.N)
        _check(ccall((:hg_get_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.handle, q), c.handle)
        n = c.N
        Q[o] = q[1:n]; Q[N .+ o] = q[n+1:2n]; Q[2N .+ o] = q[2n+1:3n]
    end
    return Q
end

"""
    swe_2d_rhs(dQdt, Q, params_vector, t, ctx) -> dQdt

Same contract as Hydrograd.swe_2d_rhs (semi_discretize_swe_2D.jl:18-277): `dQdt` is overwritten and returned.
"""
function swe_2d_rhs(dQdt::Vector{Float64}, Q::Vector{Float64}, params_vector::Vector{Float64}, t::Float64, ctx::Context)
    np = ctx.active == 0 ? 0 : length(params_vector)
    GC.@preserve dQdt Q params_vector begin
        rc = ccall((:hg_rhs, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Float64, Ptr{Float64}),
                   ctx.handle, Q, params_vector, np, ctx.active, t, dQdt)
        _check(rc, ctx.handle)
    end
    return dQdt
end
swe_2d_rhs(Q::Vector{Float64}, p::Vector{Float64}, t::Float64, ctx::Context) = swe_2d_rhs(similar(Q), Q, p, t, ctx)
# Parameters that are not a plain Vector{Float64} (the ComponentVector of the UDE network, a view handed over by an optimiser)
# are flattened first; the cotangent returned by the rrule below is projected back onto the caller's type by ChainRulesCore.
_plain(p::Vector{Float64}) = p
_plain(p::AbstractVector) = collect(Float64, p)
swe_2d_rhs(dQdt::Vector{Float64}, Q::AbstractVector, p::AbstractVector, t::Real, ctx::Context) =
    swe_2d_rhs(dQdt, _plain(Q), _plain(p), Float64(t), ctx)
swe_2d_rhs(Q::AbstractVector, p::AbstractVector, t::Real, ctx::Context) = swe_2d_rhs(Vector{Float64}(undef, length(Q)), Q, p, t, ctx)

"""
    swe_2d_rhs_jvp(Q, params_vector, t, v, pdot, ctx) -> (dQdt, J_Q v + J_p pdot)

One forward-mode pass of the device RHS (hg_rhs_jvp); `ctx` must have been created with `strict=true`.
"""
function swe_2d_rhs_jvp(Q::Vector{Float64}, params_vector::Vector{Float64}, t::Float64, v::Vector{Float64}, pdot::Vector{Float64}, ctx::Context)
    dQdt = similar(Q); jv = similar(Q)
    np = ctx.active == 0 ? 0 : length(params_vector)
    GC.@preserve Q params_vector v pdot dQdt jv begin
        rc = ccall((:hg_rhs_jvp, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   ctx.handle, Q, params_vector, np, ctx.active, t, v, (np == 0 ? C_NULL : pointer(pdot)), dQdt, jv)
        _check(rc, ctx.handle)
    end
    return dQdt, jv
end

# ForwardDiff.Dual state and / or parameters (same tag on both when both are Dual): one device pass per partial
_vals(x::AbstractVector{<:ForwardDiff.Dual}) = collect(Float64, ForwardDiff.value.(x))
_vals(x::AbstractVector) = _plain(x)
_part(x::AbstractVector{<:ForwardDiff.Dual}, k) = collect(Float64, ForwardDiff.partials.(x, k))
_part(x::AbstractVector, k) = zeros(length(x))
function _rhs_dual(::Type{ForwardDiff.Dual{Tg,V,NP}}, Q::AbstractVector, p::AbstractVector, t::Real, ctx::Context) where {Tg,V,NP}
    Qv, pv = _vals(Q), _vals(p)
    n3 = length(Qv); np = ctx.active == 0 ? 0 : length(pv)
    y = Vector{Float64}(undef, n3)
    Vm = Matrix{Float64}(undef, n3, NP); Pm = Matrix{Float64}(undef, max(np, 1), NP); JV = Matrix{Float64}(undef, n3, NP)
    for k in 1:NP                                          # column k = partial k (the C layout [K][3N] is this matrix)
        Vm[:, k] .= _part(Q, k)
        np > 0 && (Pm[1:np, k] .= _part(p, k))
    end
    GC.@preserve Qv pv Vm Pm y JV begin
        rc = ccall((:hg_rhs_jvp_multi, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Float64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   ctx.handle, Qv, pv, np, ctx.active, Float64(t), NP, Vm, (np == 0 ? C_NULL : pointer(Pm)), y, JV)
        _check(rc, ctx.handle)
    end
    return [ForwardDiff.Dual{Tg}(y[i], ForwardDiff.Partials(ntuple(k -> JV[i, k], NP))) for i in eachindex(y)]
end
swe_2d_rhs(Q::AbstractVector{D}, p::AbstractVector, t::Real, ctx::Context) where {D<:ForwardDiff.Dual} = _rhs_dual(D, Q, p, t, ctx)
swe_2d_rhs(Q::AbstractVector{<:AbstractFloat}, p::AbstractVector{D}, t::Real, ctx::Context) where {D<:ForwardDiff.Dual} = _rhs_dual(D, Q, p, t, ctx)
function swe_2d_rhs(dQdt::AbstractVector{D}, Q::AbstractVector, p::AbstractVector, t::Real, ctx::Context) where {D<:ForwardDiff.Dual}
    dQdt .= _rhs_dual(D, Q, p, t, ctx)                     # in-place twin (bInPlaceODE, solve_swe_2D.jl:230-235)
    return dQdt
end

"Vector-Jacobian product (Qbar, pbar) = (dRHS/dQ)' * lambda, (dRHS/dp)' * lambda -- what Zygote.pullback returns (debug_AD.jl:60,75)."
function swe_2d_rhs_vjp(Q::Vector{Float64}, params_vector::Vector{Float64}, t::Float64, lambda::Vector{Float64}, ctx::Context;
                        state_resident::Bool=false)
    Qbar = similar(Q); pbar = zeros(max(length(params_vector), 1))
    np = ctx.active == 0 ? 0 : length(params_vector)
    GC.@preserve Q params_vector lambda Qbar pbar begin
        # state_resident: Q is what the last forward call uploaded and nothing has moved it since -- pass NULL, only lambda is copied
        Qptr = state_resident ? Ptr{Float64}(C_NULL) : pointer(Q)
        rc = ccall((:hg_rhs_vjp, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   ctx.handle, Qptr, params_vector, np, ctx.active, t, lambda, Qbar, pbar, C_NULL)
        _check(rc, ctx.handle)
    end
    return Qbar, pbar[1:length(params_vector)]
end

"Changes whenever the state resident on the device changes (uploads, steppers, solves)."
state_generation(ctx::Context) = ccall((:hg_state_generation, LIB), Int64, (Ptr{Cvoid},), ctx.handle)

function ChainRulesCore.rrule(::typeof(swe_2d_rhs), Q::AbstractVector, p::AbstractVector, t::Real, ctx::Context)
    Qv, pv = _plain(Q), _plain(p)
    y = swe_2d_rhs(Qv, pv, Float64(t), ctx)
    gen = state_generation(ctx)                            # the primal's state stays on the device ...
    project_Q, project_p = ProjectTo(Q), ProjectTo(p)      # e.g. back onto the ComponentVector of the network parameters
    function pullback(ybar)
        # ... so a pullback called before anything else touches the context (ZygoteVJP calls it at once) ships only the cotangent
        Qbar, pbar = swe_2d_rhs_vjp(Qv, pv, Float64(t), collect(Float64, unthunk(ybar)), ctx;
                                    state_resident = state_generation(ctx) == gen)
        return NoTangent(), project_Q(Qbar), (ctx.active == 0 ? ZeroTangent() : project_p(pbar)), NoTangent(), NoTangent()
    end
    return y, pullback
end

"""
    sensitivity_tsit5(ctx, Q0, params_vector, tspan, dt; t_save, adaptive, abstol, reltol, fastpow) -> (Q_T, sensitivity, pred)

The whole sensitivity driver in one call (swe_2D_sensitivity.jl:34-80: `ForwardDiff.jacobian(forward_simulation, params_vector)`
with `solve(prob, Tsit5(), ...)` inside): `sensitivity` is the 3N x length(params_vector) Jacobian of the final state.
`ctx` must have been created with `strict=true`.
"""
function sensitivity_tsit5(ctx::Context, Q0::Vector{Float64}, params_vector::Vector{Float64}, tspan::Tuple{Float64,Float64}, dt::Float64;
                           t_save::Vector{Float64}=Float64[], adaptive::Bool=true, abstol::Float64=1e-6, reltol::Float64=1e-3, fastpow::Bool=true)
    _check(ccall((:hg_set_controller_pow, LIB), Cint, (Ptr{Cvoid}, Int32), ctx.handle, Int32(fastpow)), ctx.handle)
    QT = similar(Q0); S = Matrix{Float64}(undef, length(Q0), length(params_vector)); stats = zeros(Int64, 3)
    pred = Matrix{Float64}(undef, length(Q0), length(t_save))        # Array(pred) of the driver: values at the save times
    GC.@preserve Q0 params_vector t_save pred QT S stats begin
        rc = ccall((:hg_solve_tsit5_sens, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Float64, Float64, Float64, Int32, Float64, Float64,
                    Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                   ctx.handle, Q0, params_vector, length(params_vector), ctx.active, tspan[1], tspan[2], dt, Int32(adaptive),
                   abstol, reltol, t_save, length(t_save), pred, QT, S, stats)
        _check(rc, ctx.handle)
    end
    return QT, S, pred      # column k of S = row k of the C layout: d Q(T) / d p_k
end

"custom_ODE_solve (ode_solvers/custom_ODE_solvers.jl:36-95) on the device; returns the 3N x nSaves matrix."
# solve(prob, Tsit5(), adaptive=adaptive, dt=dt, saveat=t_save; abstol, reltol) with the state resident on the device
# (swe_2D_forward_simulation.jl:38-41).  Returns the saved states as a 3N x length(t_save) matrix, like Array(sol).
# dense=true (default): OrdinaryDiffEq's saveat (steps independent of t_save, Tsit5 dense output); dense=false: save times are stops.
function solve_tsit5(ctx::Context, Q0::Vector{Float64}, tspan::Tuple{Float64,Float64}, dt::Float64, t_save::Vector{Float64};
                     adaptive::Bool=true, abstol::Float64=1e-6, reltol::Float64=1e-3, dense::Bool=true, fastpow::Bool=true)
    # fastpow=true: the PI controller raises EEst / qold to their powers with DiffEqBase.fastpow like OrdinaryDiffEq did
    _check(ccall((:hg_set_controller_pow, LIB), Cint, (Ptr{Cvoid}, Int32), ctx.handle, Int32(fastpow)), ctx.handle)
    n3 = length(Q0)
    out = Matrix{Float64}(undef, n3, length(t_save))
    stats = zeros(Int64, 3)
    GC.@preserve Q0 t_save out stats begin
        rc = ccall((:hg_set_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx.handle, Q0)
        rc == 0 || error(unsafe_string(ccall((:hg_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.handle)))
        rc = dense ?
            ccall((:hg_solve_tsit5_dense, LIB), Cint,
                  (Ptr{Cvoid}, Float64, Float64, Float64, Int32, Float64, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}),
                  ctx.handle, tspan[1], tspan[2], dt, Int32(adaptive), abstol, reltol, t_save, length(t_save), out, stats) :
            ccall((:hg_solve_tsit5, LIB), Cint,
                   (Ptr{Cvoid}, Float64, Float64, Float64, Int32, Float64, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int64}),
                   ctx.handle, tspan[1], tspan[2], dt, Int32(adaptive), abstol, reltol, t_save, length(t_save), out, stats)
        rc == 0 || error(unsafe_string(ccall((:hg_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.handle)))
    end
    return out, (accepted=stats[1], rejected=stats[2], rhs=stats[3])
end

# solve(prob, Euler() / RK4() / AB3(), dt=dt, ...) of swe_2D_forward_simulation.jl:40-47 on the resident state (hg_set_state first)
step_ode_euler(ctx::Context, dt::Float64, nsteps::Integer) =
    _check(ccall((:hg_step_ode_euler, LIB), Cint, (Ptr{Cvoid}, Float64, Int64), ctx.handle, dt, nsteps), ctx.handle)
step_rk4(ctx::Context, dt::Float64, nsteps::Integer) =
    _check(ccall((:hg_step_rk4, LIB), Cint, (Ptr{Cvoid}, Float64, Int64), ctx.handle, dt, nsteps), ctx.handle)
step_ab3(ctx::Context, dt::Float64, nsteps::Integer; restart::Bool=false) =
    _check(ccall((:hg_step_ab3, LIB), Cint, (Ptr{Cvoid}, Float64, Int64, Int32), ctx.handle, dt, nsteps, Int32(restart)), ctx.handle)

# forward_settings.ManningN_option == "variable" (semi_discretize_swe_2D.jl:140-149): select the closure once; every
# following RHS / Euler step evaluates n(h) / n(h, |U|, ks) on the device.  kind = ManningN_function_type of the control file.
const MANNING_TYPES = Dict("constant" => 0, "power_law" => 1, "sigmoid" => 2, "inverse" => 3, "h_Umag_ks" => 4)
function set_manning_function(ctx::Context, kind::String, params::Dict, ks_cells::Union{Vector{Float64},Nothing}=nothing)
    haskey(MANNING_TYPES, kind) || error("Unknown Manning's n function type: $kind. Supported types: constant, power_law, sigmoid, inverse.")
    p = Float64[get(params, "n_lower", 0.0), get(params, "n_upper", 0.0), get(params, "k", 0.0), get(params, "h_mid", 0.0)]
    ks = ks_cells === nothing ? Ptr{Float64}(C_NULL) : pointer(ks_cells)
    GC.@preserve p ks_cells begin
        rc = ccall((:hg_set_manning_function, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}),
                   ctx.handle, Int32(MANNING_TYPES[kind]), p, ks)
        rc == 0 || error(unsafe_string(ccall((:hg_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.handle)))
    end
    return nothing
end

# ---- UDE: Manning's n from the Lux network (settings.bPerform_UDE, UDE_choice "ManningN_h" / "ManningN_h_Umag_ks") ----------------
# mirror of hg_ude_desc (include/hydrograd_b200.h); HG_UDE_MAX_HIDDEN = 3
struct UdeDesc
    choice::Int32; n_hidden::Int32
    width::NTuple{3,Int32}; activation::NTuple{3,Int32}
    layernorm::Int32
    ln_epsilon::Float64
    h_bounds::NTuple{2,Float64}; umag_bounds::NTuple{2,Float64}; ks_bounds::NTuple{2,Float64}; output_bounds::NTuple{2,Float64}
    n_params::Int64
    off_weight::NTuple{4,Int64}; off_bias::NTuple{4,Int64}; off_ln_scale::NTuple{3,Int64}; off_ln_bias::NTuple{3,Int64}
end
const UDE_CHOICE = Dict("ManningN_h" => Int32(1), "ManningN_h_Umag_ks" => Int32(2))
const UDE_ACT = Dict("relu" => Int32(1), "leakyrelu" => Int32(2), "sigmoid" => Int32(3), "tanh" => Int32(4), "softplus" => Int32(5))
_pad(v, n, T) = ntuple(i -> i <= length(v) ? T(v[i]) : T(0), n)

"""
    set_ude_model(ctx, UDE_settings, ps::ComponentVector, ks_cells; layernorm_dims = Colon())

After this call `params_vector` of `swe_2d_rhs` / `swe_2d_rhs_vjp` is the flat vector of `ps` (the ComponentArray of
`Lux.setup(rng, ude_model)[1]`, UDE/process_UDE.jl:42) and Manning's n is evaluated by the network on the device
(semi_discretize_swe_2D.jl:165-178).  Offsets are read from the axes of `ps`: Lux names the layers of
`Chain(Dense, LayerNorm, Dense, LayerNorm, ..., Dense, wrapper)` layer_1, layer_2, ...; a Dense holds (weight, bias), a
LayerNorm (bias, scale).  `layernorm_dims` is the `dims` field of the LayerNorm layers (`Colon()`, Lux's default, normalises
over the whole width x N array).
"""
function set_ude_model(ctx::Context, us, ps, ks_cells::Union{Vector{Float64},Nothing}; layernorm_dims=Colon())
    cfg = us.UDE_NN_config
    hidden = Int.(cfg["hidden_layers"]); acts = String.(cfg["activations"])
    L = length(hidden)
    flat = collect(Float64, ps)                       # the same flattening `getdata(ps)` gives the optimiser
    off(layer, field) = first(ComponentArrays.label2index(ps, "layer_$(layer).$(field)")) - 1     # 0-based start in `flat`
    offw = [off(2l - 1, :weight) for l in 1:L]; offb = [off(2l - 1, :bias) for l in 1:L]
    push!(offw, off(2L + 1, :weight)); push!(offb, off(2L + 1, :bias))
    offg = [off(2l, :scale) for l in 1:L]; offbe = [off(2l, :bias) for l in 1:L]
    b2(key) = haskey(cfg, key) ? (Float64(cfg[key][1]), Float64(cfg[key][2])) : (0.0, 1.0)
    desc = UdeDesc(UDE_CHOICE[us.UDE_choice], Int32(L), _pad(hidden, 3, Int32), _pad([UDE_ACT[a] for a in acts], 3, Int32),
                   layernorm_dims isa Colon ? Int32(2) : Int32(1), Float64(1.0f-5),
                   b2("h_bounds"), b2("Umag_bounds"), b2("ks_bounds"), b2("output_bounds"), length(flat),
                   _pad(offw, 4, Int64), _pad(offb, 4, Int64), _pad(offg, 3, Int64), _pad(offbe, 3, Int64))
    ks = ks_cells === nothing ? Ptr{Float64}(C_NULL) : pointer(ks_cells)
    GC.@preserve ks_cells begin
        rc = ccall((:hg_set_ude_model, LIB), Cint, (Ptr{Cvoid}, Ref{UdeDesc}, Ptr{Float64}), ctx.handle, Ref(desc), ks)
        _check(rc, ctx.handle)
    end
    ctx.active = HG_PARAM["UDE"]
    return nothing
end

# Gradient of  lambda_T . Q(T)  through the adaptive Tsit5 solve that just ran on ctx (solve_tsit5): discrete adjoint over its
# accepted steps (step sizes are constants of the differentiation).  Returns (Q_T, dL/dQ0, dL/dparams).
function tsit5_adjoint(ctx::Context, Q0::Vector{Float64}, params_vector::Vector{Float64}, lambda_T::Vector{Float64})
    n = Ref{Int64}(0)
    _check(ccall((:hg_last_steps, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ref{Int64}), ctx.handle, C_NULL, 0, n), ctx.handle)
    h = Vector{Float64}(undef, n[])
    QT = similar(Q0); Q0bar = similar(Q0); pbar = zeros(max(length(params_vector), 1))
    np = ctx.active == 0 ? 0 : length(params_vector)
    GC.@preserve h Q0 params_vector lambda_T QT Q0bar pbar begin
        _check(ccall((:hg_last_steps, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ref{Int64}), ctx.handle, h, length(h), n), ctx.handle)
        _check(ccall((:hg_rk_adjoint_steps, LIB), Cint,
                     (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                     ctx.handle, Int32(1), Q0, params_vector, np, ctx.active, h, length(h), lambda_T, QT, Q0bar, pbar), ctx.handle)
    end
    return QT, Q0bar, pbar[1:length(params_vector)]
end

function custom_ODE_solve(Q0::Vector{Float64}, params_vector::Vector{Float64}, tspan::Tuple{Float64,Float64}, dt::Float64, ctx::Context)
    nsave = length(tspan[1]:dt:tspan[2])
    sol = Matrix{Float64}(undef, 3 * ctx.N, nsave)
    n = Ref{Int64}(0)
    np = ctx.active == 0 ? 0 : length(params_vector)
    GC.@preserve Q0 params_vector sol begin
        rc = ccall((:hg_custom_ode_solve, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Float64, Float64, Float64, Ptr{Float64}, Int64, Ref{Int64}),
                   ctx.handle, Q0, params_vector, np, ctx.active, tspan[1], tspan[2], dt, sol, nsave, n)
        _check(rc, ctx.handle)
    end
    return sol[:, 1:n[]]
end

end # module
