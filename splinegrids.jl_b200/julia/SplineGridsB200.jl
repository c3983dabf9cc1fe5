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
            GC.@preserve As Y B begin
                check(ccall(($(sym(cname)), LIB), Cint,
                            (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{Cint},
                             Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Cvoid}),
                            devptr($out), devptr($inp), ndims(Y), Int64[size(Y)...], Int64[size(B)...], N,
                            Cint[dims_refinement...],
                            Ptr{Cvoid}[devptr(Int32.(A.row_pointer)) for A in As],     # Int64 -> Int32 if needed
                            Ptr{Cvoid}[devptr(Int32.(A.column_start)) for A in As],
                            Ptr{Cvoid}[devptr(A.nzval) for A in As], Int64[length(A.nzval) for A in As],
                            stream_ptr()), $cname)
            end
            CUDA.synchronize()
            return nothing
        end
    end
end

# Slab-sharded multi-GPU adjoint (no counterpart in the reference, SURVEY.md 8e): one task per GPU holds the slab
#   (sliced last-dimension arrays); `peer_stage` = the ranks' staging buffers mapped into this process (CUDA IPC /
#   CUDA.jl peer access), `k0`/`np` = first control plane / number of planes of this slab's support.
#   sg_evaluate_adjoint_push_f64(cp, <same arguments as sg_evaluate_adjoint_f64>, peer_stage::Ptr{Ptr{Cvoid}},
#                                world::Cint, my_rank::Cint, k0::Int64, np::Int64, max_planes::Int64, stream)
#       adjoint whose last kernel stores the finished control planes into every peer's staging slot (NVLink P2P);
#   <barrier across the ranks on the stream>;
#   sg_exchange_reduce_f64(cp, my_stage, world, k0s::Ptr{Int64}, nps::Ptr{Int64}, plane_elems, c_last, Nout,
#                          max_planes, stream)   -> every rank holds the full gradient (rank-order sum, deterministic).
# A whole iteration (evaluate!, evaluate_adjoint!) can be wrapped in CUDA.@captured: the library only enqueues work.

# K7 / K8 (src/control_points.jl:296-349, src/adjoint.jl:154-205) are reached through the unchanged Julia level loops
# of evaluate!(::LocallyRefinedControlPoints) / evaluate_adjoint!(…): override the two kernel launches the same
# way with sg_scatter_active_* / sg_gather_zero_active_* (argument lists in include/splinegrids_b200.h).

end # module
