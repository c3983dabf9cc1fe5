# Dormant: times the REAL reference package on the host CPU if a `julia` binary is ever on PATH
# (it is not in the build image; bench.py --impl reference uses the C/OpenMP restatement instead).
#   julia --threads=auto --project=/root/reference baseline/run_reference.jl
using SplineGrids, KernelAbstractions
n_cp, deg, n_s, Nout = (128, 128, 128), (3, 3, 3), (512, 512, 32), 1     # 32-plane slab of config C3
dims = SplineDimension.(n_cp, deg, n_s; float_type = Float64)
grid = SplineGrid(dims, Nout)
copyto!(grid.control_points, rand(n_cp..., Nout))
e = rand(size(grid.eval)...)
g = zero(SplineGrids.obtain(grid.control_points))
step() = (evaluate!(grid); evaluate_adjoint!(grid; eval = e, control_points = g))
step()
t = @elapsed for _ in 1:3 step() end
println("{\"impl\": \"reference-julia\", \"value\": $(3 * 2 * prod(n_s) * Nout / t), \"unit\": \"sample-values/s\", \"threads\": $(Threads.nthreads())}")
