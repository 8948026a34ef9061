# HDGB200.jl - the reference-side binding of libhdg_b200.so (include/hdg_b200.h).
#
# This is what a maintainer of Paulms/HDiscontinuousGalerkin.jl adds to make examples/poisson2D_HDG.jl run on
# a B200: the script-level functions doassemble / apply! / K\b / get_uσ! / errornorm keep their names and
# argument meaning and become thin `ccall`s.  Mesh, function-space and TrialFunction constructors stay as they
# are (src/mesh.jl, src/generate_mesh.jl, src/*FunctionSpaces.jl, src/DiscreteFunctions.jl): the arrays they
# hold are passed zero-copy.  NOT EXECUTED in this repository's CI (no Julia in the build image); the same ABI
# is exercised through Python ctypes in tests/.
module HDGB200

using HDiscontinuousGalerkin
using SparseArrays, LinearAlgebra
import HDiscontinuousGalerkin: getnbasefunctions, getncells, getnfaces, getfaceset

const lib = get(ENV, "LIBHDG_B200", "libhdg_b200.so")

struct Params            # hdg_params
    order::Int32
    quad_degree::Int32
    tau::Float64
    source_id::Int32
    device::Int32
    local_solver::Int32
    reserved::Int32
end

struct SolveInfo         # hdg_solve_info
    iterations::Int32
    converged::Int32
    relres::Float64
    bnorm::Float64
    solve_ms::Float64
end

mutable struct Context
    h::Ptr{Cvoid}
    ncell::Int; nface::Int; n::Int; nt::Int
end

function check(st::Cint, h::Ptr{Cvoid}=C_NULL)
    st == 0 && return
    msg = unsafe_string(ccall((:hdg_last_error, lib), Cstring, (Ptr{Cvoid},), h))
    st == 2 && throw(ArgumentError(msg))                 # det(J) is not positive  (src/ScalarFunctionSpaces.jl:110)
    st == 3 && throw(ArgumentError(msg))                 # quadrature rule not available (src/quadrature.jl:24)
    st == 4 && throw(LinearAlgebra.SingularException(0)) # factorize(Array(Me)) (examples/poisson2D_HDG.jl:160)
    st == 8 && throw(AssertionError(msg))                # src/boundary.jl:22
    error("libhdg_b200 status $st: $msg")
end

"""
    doassemble(Vh, Wh, Mh, τ = 1.0; f = nothing, quad_degree = order + 1) -> (K, rhs, K_element, b_element)

Drop-in for `doassemble` of examples/poisson2D_HDG.jl:58-186.  `K`, `K_element`, `b_element` are handles on
device-resident data (`SparseMatrixCSC(K)` downloads the CSC matrix; `K_element[i]` downloads one block).
`f === nothing` selects the built-in source 2π² sin(πx) sin(πy); any other function is sampled at the
quadrature points on the host exactly like `function_value` (src/DiscreteFunctions.jl:6-24).
"""
function doassemble(Vh, Wh, Mh, τ = 1.0; f = nothing, order::Int, quad_degree::Int = order + 1)
    mesh = Wh.mesh
    prm = Ref(Params(order, quad_degree, τ, f === nothing ? 1 : 0, -1, 0, 0))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:hdg_create, lib), Cint, (Ref{Params}, Ref{Ptr{Cvoid}}), prm, h))
    ctx = Context(h[], getncells(mesh), getnfaces(mesh), getnbasefunctions(Wh), getnbasefunctions(Mh))
    finalizer(c -> ccall((:hdg_destroy, lib), Cvoid, (Ptr{Cvoid},), c.h), ctx)
    bfaces = sort!(collect(getfaceset(mesh, "boundary")))
    GC.@preserve mesh bfaces begin
        # Vector{Cell{2,3,3}} is ncell x 6 Int64, Vector{Node{2,Float64}} is nnode x 2 Float64,
        # mesh.faces is a column-major Matrix{Int}: exactly the layouts hdg_set_mesh takes.
        check(ccall((:hdg_set_mesh, lib), Cint,
                    (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Float64}, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Int64),
                    ctx.h, pointer(mesh.cells), length(mesh.cells), pointer(mesh.nodes), length(mesh.nodes),
                    pointer(mesh.faces), size(mesh.faces, 1), pointer(bfaces), length(bfaces)), ctx.h)
    end
    if f !== nothing
        nq = getnquadpoints(Wh)
        fq = Matrix{Float64}(undef, nq, getncells(mesh))         # fq[q, cell] == C layout fq[cell*nq + q]
        for (ci, cell) in enumerate(CellIterator(mesh)), q in 1:nq
            fq[q, ci] = function_value(f, Wh, cell, q)
        end
        check(ccall((:hdg_set_source_values, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx.h, fq), ctx.h)
    end
    check(ccall((:hdg_assemble, lib), Cint, (Ptr{Cvoid},), ctx.h), ctx.h)
    return TraceMatrix(ctx), TraceVector(ctx, :rhs), LocalBlocks(ctx, :K), LocalBlocks(ctx, :b)
end

struct TraceMatrix; ctx::Context; end
struct TraceVector; ctx::Context; kind::Symbol; end
struct LocalBlocks; ctx::Context; kind::Symbol; end

function Base.getindex(L::LocalBlocks, cell::Int)
    c = L.ctx; m = 3c.n; t = 3c.nt
    Ke = Matrix{Float64}(undef, m, t); be = Vector{Float64}(undef, m)
    check(ccall((:hdg_get_local, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}), c.h, cell, Ke, be), c.h)
    L.kind == :K ? Ke : be
end

"`SparseMatrixCSC(K)`: colptr/rowval/nzval exactly as `sparse(I,J,V)` of src/assembler.jl:47-49 builds them."
function SparseArrays.SparseMatrixCSC(K::TraceMatrix)
    c = K.ctx; N = c.nface * c.nt
    sz = Ref{NTuple{12,Int64}}()   # hdg_sizes is read through hdg_get_sizes in real code; nnz from colptr below
    colptr = Vector{Int64}(undef, N + 1)
    check(ccall((:hdg_get_pattern, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), c.h, colptr, C_NULL), c.h)
    nnz = colptr[end] - 1
    rowval = Vector{Int64}(undef, nnz); nzval = Vector{Float64}(undef, nnz)
    check(ccall((:hdg_get_pattern, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), c.h, colptr, rowval), c.h)
    check(ccall((:hdg_get_values, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.h, nzval), c.h)
    SparseMatrixCSC(N, N, colptr, rowval, nzval)
end

function Base.Vector(v::TraceVector)
    c = v.ctx; out = Vector{Float64}(undef, c.nface * c.nt)
    fn = v.kind == :rhs ? :hdg_get_rhs : :hdg_get_trace
    check(ccall((fn, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.h, out), c.h)
    out
end

"apply!(K, b, dbc) of src/boundary.jl:121-158.  The Dirichlet faces are read back from dbc.prescribed_dofs (nt consecutive
dofs per face, ascending, src/boundary.jl:19-25), so any named face set works ("left", "bottom", ...)."
function HDiscontinuousGalerkin.apply!(K::TraceMatrix, b::TraceVector, dbc::Dirichlet)
    c = K.ctx
    faces = Int64[(d - 1) ÷ c.nt + 1 for d in dbc.prescribed_dofs[1:c.nt:end]]
    check(ccall((:hdg_set_dirichlet_faces, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64), c.h, faces, length(faces)), c.h)
    vals = any(!iszero, dbc.values) ? dbc.values : C_NULL
    check(ccall((:hdg_apply_dirichlet, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), c.h, vals), c.h)
end

"û = K \\ b of examples/poisson2D_HDG.jl:195 (Jacobi-PCG).  `solve(K, b; precond = :mg)` selects the block-Jacobi + P1-vertex
multigrid preconditioner (rectangle_mesh triangulations: recognised in the arrays passed by doassemble), `:block` the
face-block Jacobi one."
Base.:\(K::TraceMatrix, b::TraceVector) = solve(K, b)

function solve(K::TraceMatrix, b::TraceVector; rtol = 1e-13, maxit = 200_000, precond::Symbol = :jacobi)
    id = Dict(:jacobi => 0, :block => 1, :mg => 2)[precond]
    check(ccall((:hdg_set_preconditioner, lib), Cint, (Ptr{Cvoid}, Int32), K.ctx.h, id), K.ctx.h)
    info = Ref(SolveInfo(0, 0, 0.0, 0.0, 0.0))
    check(ccall((:hdg_solve, lib), Cint, (Ptr{Cvoid}, Float64, Int32, Ref{SolveInfo}), K.ctx.h, rtol, maxit, info), K.ctx.h)
    TraceVector(K.ctx, :trace)
end

"get_uσ!(σ_h, u_h, û_h, û, K_e, b_e, mesh) of examples/poisson2D_HDG.jl:197-212: fills the m_values arrays."
function get_uσ!(σ_h, u_h, û_h, û::TraceVector, K_e::LocalBlocks, b_e::LocalBlocks, mesh)
    c = û.ctx
    check(ccall((:hdg_recover, lib), Cint, (Ptr{Cvoid},), c.h), c.h)
    check(ccall((:hdg_get_mvalues, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                c.h, σ_h.m_values, u_h.m_values, û_h.m_values), c.h)     # column-major ncell x nb: same layout
end

"errornorm(u_h, u_ex) (squared L2 error, src/DiscreteFunctions.jl:97-120) for u_ex = sin(πx) sin(πy)."
function errornorm_b200(ctx::Context)
    e = Ref{Float64}(0.0)
    check(ccall((:hdg_errornorm, lib), Cint, (Ptr{Cvoid}, Int32, Ref{Float64}), ctx.h, 1, e), ctx.h)
    e[]
end

"errornorm(u_h, u_ex) for any u_ex: sampled at the cell quadrature points like `function_value` (src/DiscreteFunctions.jl:108)."
function errornorm_b200(ctx::Context, Wh, mesh, u_ex::Function)
    nq = getnquadpoints(Wh)
    uq = Matrix{Float64}(undef, nq, getncells(mesh))              # uq[q, cell] == C layout uex_q[cell*nq + q]
    for (ci, cell) in enumerate(CellIterator(mesh)), q in 1:nq
        uq[q, ci] = function_value(u_ex, Wh, cell, q)
    end
    e = Ref{Float64}(0.0)
    check(ccall((:hdg_errornorm_values, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), ctx.h, uq, e), ctx.h)
    e[]
end

# ---- mesh pre-processing for several GPUs: Morton-curve order of the cells (hdg_order_cells), then the reference's own
# first-encounter numbering of the permuted element list, on the device (hdg_number_faces == parse_cells!, src/triangle_mesh.jl:48-108)
function order_cells(ctx::Context, tri::Matrix{Int64}, nodes::Matrix{Float64})     # tri: 3 x ncell, nodes: 2 x nnode (Julia column-major)
    perm = Vector{Int64}(undef, size(tri, 2))
    check(ccall((:hdg_order_cells, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Float64}, Int64, Ptr{Int64}),
                ctx.h, tri, size(tri, 2), nodes, size(nodes, 2), perm), ctx.h)
    return perm                                                                   # 1-based: new cell i is old cell perm[i]
end

# ---- CG side (examples/poisson2D_CG.jl): DofHandler numbering + sparsity pattern, assembly, apply!, solve, errornorm
function cg_poisson(ctx::Context, order::Int; rtol = 1e-12, maxit = 100_000)
    nd = Ref{Int64}(0)
    check(ccall((:hdg_cg_setup, lib), Cint, (Ptr{Cvoid}, Int32, Ref{Int64}), ctx.h, order, nd), ctx.h)
    check(ccall((:hdg_cg_assemble, lib), Cint, (Ptr{Cvoid},), ctx.h), ctx.h)
    check(ccall((:hdg_cg_apply_dirichlet, lib), Cint, (Ptr{Cvoid},), ctx.h), ctx.h)
    info = Ref(SolveInfo())
    check(ccall((:hdg_cg_solve, lib), Cint, (Ptr{Cvoid}, Float64, Int32, Ref{SolveInfo}), ctx.h, rtol, maxit, info), ctx.h)
    e = Ref{Float64}(0.0)
    check(ccall((:hdg_cg_errornorm, lib), Cint, (Ptr{Cvoid}, Ref{Float64}), ctx.h, e), ctx.h)
    return nd[], info[], e[]
end

end
