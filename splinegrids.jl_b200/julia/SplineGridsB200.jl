# SplineGridsB200.jl -- thin `ccall` shim that routes SplineGrids.jl's hot path to libsplinegrids_b200.so
# for arrays living on a CUDA device (CuArray).  Marshalling only: every method keeps the reference's
# signature and semantics (in place, returns `nothing`, synchronous on return) and replaces exactly one
# KernelAbstractions launch.  NOT EXECUTED in the build image (Julia is not installed there); the Python
# ctypes binding `splinegrids.jl_b200/_lib.py` marshals the identical argument lists and is what the tests run.
#
# Usage:  using SplineGrids, CUDA;  include("SplineGridsB200.jl");  using .SplineGridsB200
#         (set ENV["SPLINEGRIDS_B200_LIB"] to the path of libsplinegrids_b200.so)
module SplineGridsB200

using SplineGrids
using SplineGrids: AbstractSplineGrid, SplineDimension, RefinementMatrix, LocallyRefinedControlPoints,
                   obtain, validate_partial_derivatives, validate_mult_input, get_n_basis_functions
using CUDA

const LIB = get(ENV, "SPLINEGRIDS_B200_LIB", "libsplinegrids_b200.so")

suffix(::Type{Float32}) = "f32"
suffix(::Type{Float64}) = "f64"
check(status::Cint, what) = status == 0 || error("$what failed: status $status: " *
    unsafe_string(ccall((:sg_status_string, LIB), Cstring, (Cint,), status)))
stream_ptr() = Base.unsafe_convert(Ptr{Cvoid}, CUDA.stream().handle)
devptr(a::CuArray) = reinterpret(Ptr{Cvoid}, pointer(a))
devptr(::Nothing) = C_NULL

int32_device(a::CuArray{Int32}) = a
int32_device(a::CuArray{<:Integer}) = Int32.(a)            # a NEW device array: the caller keeps it rooted

mutable struct AdjointPlan
    handle::Ptr{Cvoid}
end

# The symbol must be a constant for ccall: generate one method per float type.
for (Tv, suf) in ((Float32, "f32"), (Float64, "f64"))
    sym(name) = QuoteNode(Symbol(name, "_", suf))

    # ---- K1  set_sample_indices!   (src/utils.jl:19-29, kernel src/util_kernels.jl:22-49)
    @eval function SplineGrids.set_sample_indices!(sd::SplineDimension{$Tv, Int32, <:Any, <:CuArray})::Nothing
        ka = sd.knot_vector.knots_all
        check(ccall(($(sym("sg_span_indices")), LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}),
                    devptr(sd.sample_indices), devptr(sd.sample_points), length(sd.sample_points),
                    devptr(ka), length(ka), sd.degree, stream_ptr()), "sg_span_indices")
        CUDA.synchronize()
        return nothing
    end

    # ---- K2  evaluate!(::SplineDimension)   (src/spline_dimension.jl:231-242)
    @eval function SplineGrids.evaluate!(sd::SplineDimension{$Tv, Int32, <:Any, <:CuArray})::Nothing
        ka = sd.knot_vector.knots_all
        check(ccall(($(sym("sg_basis_tables")), LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Cint, Ptr{Cvoid}),
                    devptr(sd.eval), devptr(ka), length(ka), devptr(sd.sample_points), devptr(sd.sample_indices),
                    length(sd.sample_points), sd.degree, sd.max_derivative_order, stream_ptr()), "sg_basis_tables")
        CUDA.synchronize()
        return nothing
    end

    # ---- K3  evaluate!(::AbstractSplineGrid)   (src/spline_grid.jl:200-230)
    @eval function SplineGrids.evaluate!(grid::AbstractSplineGrid{Nin, Nout, HasWeights, $Tv};
            derivative_order::NTuple{Nin, <:Integer} = ntuple(_ -> 0, Nin),
            control_points = grid.control_points,
            eval::CuArray = grid.eval)::Nothing where {Nin, Nout, HasWeights}
        validate_partial_derivatives(grid, derivative_order)
        @assert size(control_points) == size(grid.control_points)
        @assert size(eval) == size(grid.eval)
        sds = grid.spline_dimensions
        cp = obtain(control_points)
        GC.@preserve sds cp eval begin
            check(ccall(($(sym("sg_evaluate")), LIB), Cint,
                        (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}},
                         Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                        devptr(eval), Nin,
                        Int64[length(sd.sample_points) for sd in sds],
                        Int64[get_n_basis_functions(sd) for sd in sds], Nout,
                        Ptr{Cvoid}[devptr(sd.eval) for sd in sds],
                        Ptr{Cvoid}[devptr(sd.sample_indices) for sd in sds],
                        Cint[sd.degree for sd in sds], Cint[sd.max_derivative_order for sd in sds],
                        Cint[derivative_order...], devptr(cp), devptr(grid.weights), stream_ptr()), "sg_evaluate")
        end
        CUDA.synchronize()
        return nothing
    end

    # ---- K4  evaluate_adjoint!(::AbstractSplineGrid{…,false})   (src/adjoint.jl:52-83)
    @eval function SplineGrids.evaluate_adjoint!(grid::AbstractSplineGrid{Nin, Nout, false, $Tv};
            derivative_order::NTuple{Nin, <:Integer} = ntuple(_ -> 0, Nin),
            control_points = grid.control_points,
            eval::CuArray = grid.eval)::Nothing where {Nin, Nout}
        validate_partial_derivatives(grid.spline_dimensions, derivative_order)
        cp = obtain(control_points)
        @assert size(cp) == size(grid.control_points)
        @assert size(eval) == size(grid.eval)
        sds = grid.spline_dimensions
        GC.@preserve sds cp eval begin
            check(ccall(($(sym("sg_evaluate_adjoint")), LIB), Cint,
                        (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}},
                         Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                        devptr(cp), Nin,
                        Int64[length(sd.sample_points) for sd in sds],
                        Int64[get_n_basis_functions(sd) for sd in sds], Nout,
                        Ptr{Cvoid}[devptr(sd.eval) for sd in sds],
                        Ptr{Cvoid}[devptr(sd.sample_indices) for sd in sds],
                        Cint[sd.degree for sd in sds], Cint[sd.max_derivative_order for sd in sds],
                        Cint[derivative_order...], devptr(eval), C_NULL,
                        C_NULL, 0,            # workspace: NULL -> cudaMallocAsync on the stream
                        stream_ptr()), "sg_evaluate_adjoint")
        end
        CUDA.synchronize()
        return nothing
    end

    # ---- K5 / K6  mult! / mult_adjoint!   (src/refinement_matrix.jl:421-445, src/adjoint.jl:127-152)
    for (jlname, cname, out, inp) in ((:mult!, "sg_refmat_mul", :Y, :B), (:mult_adjoint!, "sg_refmat_mul_adjoint", :B, :Y))
        args = jlname == :mult! ? :(Y::CuArray{$Tv}, As::NTuple{N, <:RefinementMatrix}, B::CuArray{$Tv}) :
                                  :(B::CuArray{$Tv}, As::NTuple{N, <:RefinementMatrix}, Y::CuArray{$Tv})
        @eval function SplineGrids.$jlname($(args.args...), dims_refinement::NTuple{N, <:Integer})::Nothing where {N}
            validate_mult_input(Y, As, B, dims_refinement)
            # Int32 index arrays as NAMED locals (a RefinementMatrix(::Matrix) carries Int64, src/refinement_matrix.jl:476-479):
            # the converted copies must stay rooted until the kernel has run, so they are part of the GC.@preserve list
            # (a temporary built inside the argument comprehension could be finalised before the launch).
            rps = map(A -> int32_device(A.row_pointer), As)
            css = map(A -> int32_device(A.column_start), As)
            nzs = map(A -> A.nzval, As)
            GC.@preserve As Y B rps css nzs begin
                check(ccall(($(sym(cname)), LIB), Cint,
                            (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{Cint},
                             Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Cvoid}),
                            devptr($out), devptr($inp), ndims(Y), Int64[size(Y)...], Int64[size(B)...], N,
                            Cint[dims_refinement...],
                            Ptr{Cvoid}[devptr(a) for a in rps], Ptr{Cvoid}[devptr(a) for a in css],
                            Ptr{Cvoid}[devptr(a) for a in nzs], Int64[length(a) for a in nzs],
                            stream_ptr()), $cname)
                CUDA.synchronize()                     # inside the preserve block: the kernel has consumed the arrays
            end
            return nothing
        end
    end

    # ---- K7  local_refinement_kernel launch of evaluate!(::LocallyRefinedControlPoints)   (src/control_points.jl:339-347)
    @eval function scatter_active!(cp_new::CuArray{$Tv}, refinement_indices::CuMatrix{Int32}, refinement_values::CuMatrix{$Tv})::Nothing
        n_active = size(refinement_indices, 1)
        n_active == 0 && return nothing
        nin = ndims(cp_new) - 1
        GC.@preserve cp_new refinement_indices refinement_values begin
            check(ccall(($(sym("sg_scatter_active")), LIB), Cint,
                        (Ptr{Cvoid}, Cint, Ptr{Int64}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
                        devptr(cp_new), nin, Int64[size(cp_new)[1:nin]...], size(cp_new, nin + 1),
                        devptr(refinement_indices), devptr(refinement_values), n_active, stream_ptr()), "sg_scatter_active")
            CUDA.synchronize()
        end
        return nothing
    end

    # ---- K8  local_refinement_adjoint_kernel launch of evaluate_adjoint!(::LocallyRefinedControlPoints)   (src/adjoint.jl:185-193)
    @eval function gather_zero_active!(refinement_values::CuMatrix{$Tv}, cp_new::CuArray{$Tv}, refinement_indices::CuMatrix{Int32})::Nothing
        n_active = size(refinement_indices, 1)
        n_active == 0 && return nothing
        nin = ndims(cp_new) - 1
        GC.@preserve cp_new refinement_indices refinement_values begin
            check(ccall(($(sym("sg_gather_zero_active")), LIB), Cint,
                        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int64}, Cint, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
                        devptr(refinement_values), devptr(cp_new), nin, Int64[size(cp_new)[1:nin]...], size(cp_new, nin + 1),
                        devptr(refinement_indices), n_active, stream_ptr()), "sg_gather_zero_active")
            CUDA.synchronize()
        end
        return nothing
    end

    # ---- adjoint plan: evaluate_adjoint! without the per-call preparation (include/splinegrids_b200.h, "adjoint plan")
    @eval function plan_create(grid::AbstractSplineGrid{Nin, Nout, false, $Tv}, derivative_order::NTuple{Nin, <:Integer}) where {Nin, Nout}
        sds = grid.spline_dimensions
        handle = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve sds begin
            check(ccall(($(sym("sg_adjoint_plan_create")), LIB), Cint,
                        (Ptr{Ptr{Cvoid}}, Cint, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Cint}, Ptr{Cint},
                         Ptr{Cint}, Cint, Ptr{Cvoid}),
                        handle, Nin, Int64[length(sd.sample_points) for sd in sds], Int64[get_n_basis_functions(sd) for sd in sds],
                        Nout, Ptr{Cvoid}[devptr(sd.eval) for sd in sds], Ptr{Cvoid}[devptr(sd.sample_indices) for sd in sds],
                        Cint[sd.degree for sd in sds], Cint[sd.max_derivative_order for sd in sds], Cint[derivative_order...], 0,
                        stream_ptr()), "sg_adjoint_plan_create")
        end
        plan = AdjointPlan(handle[])
        finalizer(p -> ccall((:sg_adjoint_plan_destroy, LIB), Cint, (Ptr{Cvoid},), p.handle), plan)
        return plan
    end
    @eval function evaluate_adjoint_planned!(plan::AdjointPlan, cp::CuArray{$Tv}, eval::CuArray{$Tv}, workspace::CuVector{UInt8})::Nothing
        GC.@preserve plan cp eval workspace begin
            check(ccall(($(sym("sg_evaluate_adjoint_planned")), LIB), Cint,
                        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64,
                         Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                        plan.handle, devptr(cp), devptr(eval), C_NULL, devptr(workspace), length(workspace), C_NULL, 0, 0, 0, 0, 0, 1,
                        C_NULL, stream_ptr()), "sg_evaluate_adjoint_planned")
            CUDA.synchronize()
        end
        return nothing
    end

    # ---- value + partial derivatives in one call (docs/src/examples_optics.md:189-191, examples_pde.md:69-72)
    @eval function evaluate_multi!(grid::AbstractSplineGrid{Nin, Nout, false, $Tv}, derivative_orders::Vector{<:NTuple{Nin, <:Integer}},
            evals::Vector{<:CuArray{$Tv}}; control_points = grid.control_points)::Nothing where {Nin, Nout}
        foreach(d -> validate_partial_derivatives(grid, d), derivative_orders)
        sds = grid.spline_dimensions
        cp = obtain(control_points)
        GC.@preserve sds cp evals begin
            check(ccall(($(sym("sg_evaluate_multi")), LIB), Cint,
                        (Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}, Cint, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Cint},
                         Ptr{Cint}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                        Ptr{Cvoid}[devptr(e) for e in evals], length(evals), Cint[d for der in derivative_orders for d in der], Nin,
                        Int64[length(sd.sample_points) for sd in sds], Int64[get_n_basis_functions(sd) for sd in sds], Nout,
                        Ptr{Cvoid}[devptr(sd.eval) for sd in sds], Ptr{Cvoid}[devptr(sd.sample_indices) for sd in sds],
                        Cint[sd.degree for sd in sds], Cint[sd.max_derivative_order for sd in sds], devptr(cp), C_NULL, stream_ptr()),
                  "sg_evaluate_multi")
            CUDA.synchronize()
        end
        return nothing
    end
end


# ---- hierarchy loops with the two kernel launches replaced (the loops themselves are the reference's) -------------------
function SplineGrids.evaluate!(control_points::LocallyRefinedControlPoints{Nin, Nout, Tv, Int32, <:CuArray}) where {Nin, Nout, Tv}
    (; control_points_refined, local_refinements) = control_points
    for (i, lr) in enumerate(local_refinements)                                   # src/control_points.jl:324-348
        cp_new = control_points_refined[i]
        i > 1 && SplineGrids.mult!(cp_new, Tuple(lr.refinement_matrices), control_points_refined[i - 1], Tuple(lr.dims_refinement))
        scatter_active!(cp_new, lr.refinement_indices, lr.refinement_values)      # K7
    end
end
function SplineGrids.evaluate_adjoint!(control_points::LocallyRefinedControlPoints{Nin, Nout, Tv, Int32, <:CuArray})::Nothing where {Nin, Nout, Tv}
    (; control_points_refined, local_refinements) = control_points
    for (i, lr) in Iterators.reverse(enumerate(local_refinements))                 # src/adjoint.jl:179-203
        cp_new = control_points_refined[i]
        gather_zero_active!(lr.refinement_values, cp_new, lr.refinement_indices)  # K8
        i > 1 && SplineGrids.mult_adjoint!(control_points_refined[i - 1], Tuple(lr.refinement_matrices), cp_new, Tuple(lr.dims_refinement))
    end
    return nothing
end

# ---- multi-GPU gradient exchange (no counterpart in the reference, SURVEY.md 8e): one task per GPU ---------------------------
# peer_stage / peer_flags: the ranks' staging buffers and flag arrays mapped into this process (CUDA IPC or CUDA.jl peer access),
# my_flags / local_sync: this rank's own flag array (world x UInt64, zeroed once) and 64 bytes of local device memory (zeroed once).
function adjoint_push!(plan::AdjointPlan, cp::CuArray{Float64}, eval::CuArray{Float64}, workspace::CuVector{UInt8},
        peer_stage::Vector{Ptr{Cvoid}}, my_rank::Integer, k0::Integer, np::Integer, max_planes::Integer; keep_local::Bool = false,
        multicast_stage::Ptr{Cvoid} = C_NULL)   # NVLS multicast address of the staging buffers (CUDA.jl: cuMulticast* mapping)
    GC.@preserve plan cp eval workspace peer_stage begin
        check(ccall((:sg_evaluate_adjoint_planned_f64, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Ptr{Cvoid}}, Cint, Cint, Int64, Int64, Int64,
                     Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                    plan.handle, devptr(cp), devptr(eval), C_NULL, devptr(workspace), length(workspace), peer_stage, length(peer_stage),
                    my_rank, k0, np, max_planes, keep_local, multicast_stage, stream_ptr()), "sg_evaluate_adjoint_planned (push)")
    end
end
function exchange_wait_reduce!(grad::CuArray{Float64}, my_stage::CuArray{Float64}, my_flags::CuVector{UInt64}, local_sync::CuVector{UInt8},
        peer_flags::Vector{Ptr{Cvoid}}, my_rank::Integer, k0s::Vector{Int64}, nps::Vector{Int64}, max_planes::Integer)
    world = length(peer_flags)
    plane_elems = prod(size(grad)[1:(ndims(grad) - 2)])
    GC.@preserve grad my_stage my_flags local_sync peer_flags k0s nps begin
        check(ccall((:sg_exchange_wait_reduce_f64, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Cint,
                     Int64, Ptr{Cvoid}),
                    devptr(grad), devptr(my_stage), devptr(my_flags), devptr(local_sync), peer_flags, world, my_rank, k0s, nps,
                    plane_elems, size(grad, ndims(grad) - 1), size(grad, ndims(grad)), max_planes, stream_ptr()),
              "sg_exchange_wait_reduce")
    end
end
# Support-plane (halo) variant: planes go to the ranks whose slabs read them, the barrier involves those neighbours only, and
# grad is written on this rank's support planes k0s[my_rank+1] .. k0s[my_rank+1] + nps[my_rank+1] - 1 (0-based) only.
function adjoint_push_support!(plan::AdjointPlan, cp::CuArray{Float64}, eval::CuArray{Float64}, workspace::CuVector{UInt8},
        peer_stage::Vector{Ptr{Cvoid}}, my_rank::Integer, k0s::Vector{Int64}, nps::Vector{Int64}, max_planes::Integer;
        keep_local::Bool = false)
    GC.@preserve plan cp eval workspace peer_stage k0s nps begin
        check(ccall((:sg_evaluate_adjoint_planned_support_f64, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{Int64},
                     Ptr{Int64}, Int64, Cint, Ptr{Cvoid}),
                    plan.handle, devptr(cp), devptr(eval), C_NULL, devptr(workspace), length(workspace), peer_stage, length(peer_stage),
                    my_rank, k0s, nps, max_planes, keep_local, stream_ptr()), "sg_evaluate_adjoint_planned_support")
    end
end
function exchange_wait_reduce_support!(grad::CuArray{Float64}, my_stage::CuArray{Float64}, my_flags::CuVector{UInt64},
        local_sync::CuVector{UInt8}, peer_flags::Vector{Ptr{Cvoid}}, my_rank::Integer, k0s::Vector{Int64}, nps::Vector{Int64},
        max_planes::Integer)
    world = length(peer_flags)
    plane_elems = prod(size(grad)[1:(ndims(grad) - 2)])
    GC.@preserve grad my_stage my_flags local_sync peer_flags k0s nps begin
        check(ccall((:sg_exchange_wait_reduce_support_f64, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Cint,
                     Int64, Ptr{Cvoid}),
                    devptr(grad), devptr(my_stage), devptr(my_flags), devptr(local_sync), peer_flags, world, my_rank, k0s, nps,
                    plane_elems, size(grad, ndims(grad) - 1), size(grad, ndims(grad)), max_planes, stream_ptr()),
              "sg_exchange_wait_reduce_support")
    end
end
# A whole iteration (evaluate!, adjoint_push!, exchange_wait_reduce!) can be wrapped in CUDA.@captured: the library only
# enqueues kernels, the barrier between the ranks is a device-side flag wait.

# ---- AD rules for the opaque boundary (scope row f2) ----------------------------------------------------------------------
# The reference differentiates THROUGH its KernelAbstractions kernels with Enzyme (ext/SplineGridsEnzymeExt.jl,
# test/test_EnzymeExt.jl:24-49); a `ccall` is opaque to Enzyme, so the rules are given explicitly.  evaluate! is LINEAR in
# the control points (src/spline_grid.jl:130-182): the tangent of eval is evaluate! applied to the tangent control points
# and the cotangent of the control points is evaluate_adjoint! applied to the cotangent of eval (src/adjoint.jl:52-83).
# Loaded only when Enzyme is available (package extension style).
"""
    evaluate_cp!(eval, grid, control_points, derivative_order)

Positional form of `evaluate!(grid; control_points, eval, derivative_order)` that the AD rules above attach to (Enzyme
custom rules dispatch on positional arguments).  A loss written as in test/test_EnzymeExt.jl:24-29 calls this instead of
the keyword form.
"""
function evaluate_cp!(eval::CuArray, grid::AbstractSplineGrid, control_points::CuArray, derivative_order)::Nothing
    SplineGrids.evaluate!(grid; control_points, eval, derivative_order)
end

@static if isdefined(Base, :get_extension) && Base.find_package("Enzyme") !== nothing
    import Enzyme
    import Enzyme: EnzymeRules
    using Enzyme: Const, Duplicated, Annotation

    # forward mode: d(eval) = evaluate!(grid; control_points = d(control_points))
    function EnzymeRules.forward(config, func::Const{typeof(evaluate_cp!)}, ::Type{<:Const}, eval::Duplicated{<:CuArray},
            grid::Const, cp::Duplicated{<:CuArray}, derivative_order::Const)
        SplineGrids.evaluate!(grid.val; control_points = cp.val, eval = eval.val, derivative_order = derivative_order.val)
        SplineGrids.evaluate!(grid.val; control_points = cp.dval, eval = eval.dval, derivative_order = derivative_order.val)
        return nothing
    end
    # reverse mode: nothing to remember (the map is linear); d(control_points) += evaluate_adjoint!(d(eval)), d(eval) = 0
    function EnzymeRules.augmented_primal(config, func::Const{typeof(evaluate_cp!)}, ::Type{<:Const}, eval::Duplicated{<:CuArray},
            grid::Const, cp::Duplicated{<:CuArray}, derivative_order::Const)
        SplineGrids.evaluate!(grid.val; control_points = cp.val, eval = eval.val, derivative_order = derivative_order.val)
        return EnzymeRules.AugmentedReturn(nothing, nothing, nothing)
    end
    function EnzymeRules.reverse(config, func::Const{typeof(evaluate_cp!)}, ::Type{<:Const}, tape, eval::Duplicated{<:CuArray},
            grid::Const, cp::Duplicated{<:CuArray}, derivative_order::Const)
        g = similar(cp.dval)
        SplineGrids.evaluate_adjoint!(grid.val; control_points = g, eval = eval.dval, derivative_order = derivative_order.val)
        cp.dval .+= g                       # accumulate into the shadow
        fill!(eval.dval, 0)                 # eval is overwritten by the primal: its incoming cotangent is consumed
        return (nothing, nothing, nothing, nothing)
    end
    # the reference's only Enzyme method: zero the shadow of a grid (ext/SplineGridsEnzymeExt.jl:5-11)
    function Enzyme.make_zero!(grid::SplineGrids.SplineGrid{<:Any, <:Any, <:Any, <:Any, <:Any, <:CuArray})::Nothing
        fill!(grid.eval, 0)
        foreach(sd -> fill!(sd.eval, 0), grid.spline_dimensions)
        return nothing
    end
end
end # module
