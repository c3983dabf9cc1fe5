/* CPU oracle, C restatement -- body template, included once per floating type by oracle.c.
 *
 * TEST INFRASTRUCTURE ONLY (checker + timed CPU baseline); never linked into the product.
 * Every function restates one KernelAbstractions kernel of the reference with the
 * reference's own algorithm: one work item per ndrange element, the full prod(p_d+1)
 * window loop, accumulation straight into the output array, atomics in the adjoints.
 * OpenMP `parallel for schedule(static)` over the ndrange stands in for KernelAbstractions'
 * CPU() backend, which splits the ndrange statically over Threads.nthreads().
 * Citations are path:line in the reference checkout.  Index arrays are 1-based.
 *
 * Expects: T (float|double), FN(name) -> name_f32|name_f64.
 */

/* K1 set_sample_indices_kernel -- src/util_kernels.jl:22-49 */
void FN(orc_span_indices)(const T *sample_points, int64_t n, const T *knots_all, int64_t n_knots,
                          int degree, int32_t *sample_indices)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        T t = sample_points[i];
        int64_t idx = 0;
        for (int64_t k = 0; k < n_knots; ++k) {
            if (t < knots_all[k]) break;
            idx++;
        }
        int64_t lo = degree + 1, hi = n_knots - degree - 1;
        if (idx < lo) idx = lo;
        if (idx > hi) idx = hi;
        sample_indices[i] = (int32_t)idx;
    }
}

/* K2 spline_dimension_kernel -- src/spline_dimension.jl:160-215.
 * eval is (n, p+1, mdo+1) column-major.  eval_prev of the reference is a per-sample local here
 * (same arithmetic, same order). */
void FN(orc_basis_tables)(T *eval, const T *knots_all, const T *sample_points,
                          const int32_t *sample_indices, int64_t n, int degree, int mdo)
{
    const int p = degree, w = p + 1, nd = mdo + 1;
#pragma omp parallel for schedule(static)
    for (int64_t l = 0; l < n; ++l) {
        T cur[ORC_MAXW * ORC_MAXW], prev[ORC_MAXW * ORC_MAXW]; /* [k_ + w*d] */
        T t = sample_points[l];
        int64_t i = sample_indices[l]; /* 1-based */
        for (int q = 0; q < w * nd; ++q) { cur[q] = 0; prev[q] = 0; }
        cur[0] = 1; prev[0] = 1;
        for (int k = 1; k <= p; ++k) {
            for (int q = 0; q < w * nd; ++q) cur[q] = 0;
            for (int k_ = 1; k_ <= k; ++k_) {
                T t_min = knots_all[i + k_ - k - 1];
                T t_max = knots_all[i + k_ - 1];
                T dt = t_max - t_min;
                T frac = prev[k_ - 1] / dt;
                cur[k_ - 1] += frac * (t_max - t);
                cur[k_] = frac * (t - t_min);
                for (int d = 1; d <= mdo + k - p; ++d) {
                    T c = prev[(k_ - 1) + w * (d - 1)] * (T)k / dt;
                    cur[(k_ - 1) + w * d] -= c;
                    cur[k_ + w * d] = c;
                }
            }
            if (k != p) for (int q = 0; q < w * nd; ++q) prev[q] = cur[q];
        }
        for (int d = 0; d < nd; ++d)
            for (int j = 0; j < w; ++j) eval[l + n * (j + (int64_t)w * d)] = cur[j + w * d];
    }
}

/* K3 spline_eval_kernel -- src/spline_grid.jl:119-183.
 * tables[d] is (n_d, p_d+1, mdo_d+1) column-major; cp is (c_1..c_D, nout); eval (n_1..n_D, nout). */
void FN(orc_evaluate)(T *eval, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                      const T *const *tables, const int32_t *const *indices, const int *degree,
                      const int *der, const T *cp, const T *weights)
{
    int64_t n_total = 1, cp_total = 1, nwin = 1;
    int64_t cp_stride[ORC_MAXD];
    for (int d = 0; d < nin; ++d) {
        cp_stride[d] = cp_total;
        n_total *= n_samples[d];
        cp_total *= n_cp[d];
        nwin *= degree[d] + 1;
    }
#pragma omp parallel for schedule(static)
    for (int64_t lin = 0; lin < n_total; ++lin) {
        int64_t J[ORC_MAXD], base = 0, r = lin;
        for (int d = 0; d < nin; ++d) {
            J[d] = r % n_samples[d];
            r /= n_samples[d];
            base += (int64_t)(indices[d][J[d]] - degree[d] - 1) * cp_stride[d];
        }
        for (int o = 0; o < nout; ++o) eval[lin + n_total * o] = 0;          /* :136-138 */
        T denom = 0;
        int I[ORC_MAXD] = {0};
        for (int64_t wv = 0; wv < nwin; ++wv) {                               /* :146 */
            T prod = 1;
            int64_t off = base;
            for (int d = 0; d < nin; ++d) {                                   /* :153-156 */
                prod *= tables[d][J[d] + n_samples[d] * (I[d] + (int64_t)(degree[d] + 1) * der[d])];
                off += I[d] * cp_stride[d];
            }
            if (weights) { prod *= weights[off]; denom += prod; }             /* :159-166 */
            for (int o = 0; o < nout; ++o)                                    /* :171-174 */
                eval[lin + n_total * o] += prod * cp[off + cp_total * o];
            for (int d = 0; d < nin; ++d) {                                   /* next offset, dim 1 fastest */
                if (++I[d] <= degree[d]) break;
                I[d] = 0;
            }
        }
        if (weights) for (int o = 0; o < nout; ++o) eval[lin + n_total * o] /= denom;   /* :178-182 */
    }
}

/* K4 spline_eval_adjoint_kernel + zero fill -- src/adjoint.jl:1-40, :61.
 * weights != NULL is the (unpinned) NURBS extension: b -> b*w/denom. */
void FN(orc_evaluate_adjoint)(T *cp, int nin, const int64_t *n_samples, const int64_t *n_cp, int nout,
                              const T *const *tables, const int32_t *const *indices,
                              const int *degree, const int *der, const T *eval, const T *weights)
{
    int64_t n_total = 1, cp_total = 1, nwin = 1;
    int64_t cp_stride[ORC_MAXD];
    for (int d = 0; d < nin; ++d) {
        cp_stride[d] = cp_total;
        n_total *= n_samples[d];
        cp_total *= n_cp[d];
        nwin *= degree[d] + 1;
    }
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < cp_total * nout; ++q) cp[q] = 0;                  /* :61 */
#pragma omp parallel for schedule(static)
    for (int64_t lin = 0; lin < n_total; ++lin) {
        int64_t J[ORC_MAXD], base = 0, r = lin;
        for (int d = 0; d < nin; ++d) {
            J[d] = r % n_samples[d];
            r /= n_samples[d];
            base += (int64_t)(indices[d][J[d]] - degree[d] - 1) * cp_stride[d];
        }
        T denom = 0;
        if (weights) {
            int I[ORC_MAXD] = {0};
            for (int64_t wv = 0; wv < nwin; ++wv) {
                T prod = 1;
                int64_t off = base;
                for (int d = 0; d < nin; ++d) {
                    prod *= tables[d][J[d] + n_samples[d] * (I[d] + (int64_t)(degree[d] + 1) * der[d])];
                    off += I[d] * cp_stride[d];
                }
                denom += prod * weights[off];
                for (int d = 0; d < nin; ++d) { if (++I[d] <= degree[d]) break; I[d] = 0; }
            }
        }
        int I[ORC_MAXD] = {0};
        for (int64_t wv = 0; wv < nwin; ++wv) {                               /* :21 */
            T prod = 1;
            int64_t off = base;
            for (int d = 0; d < nin; ++d) {
                prod *= tables[d][J[d] + n_samples[d] * (I[d] + (int64_t)(degree[d] + 1) * der[d])];
                off += I[d] * cp_stride[d];
            }
            if (weights) prod = prod * weights[off] / denom;
            for (int o = 0; o < nout; ++o) {                                  /* :33-38 */
                T v = prod * eval[lin + n_total * o];
#pragma omp atomic
                cp[off + cp_total * o] += v;
            }
            for (int d = 0; d < nin; ++d) { if (++I[d] <= degree[d]) break; I[d] = 0; }
        }
    }
}

/* get_column_range -- src/refinement_matrix.jl:103-125 (row i 1-based) */
static inline void FN(orc_colrange)(const int32_t *rp, const int32_t *cs, int64_t m, int64_t nnz,
                                    int64_t i, int64_t *c0, int64_t *nc)
{
    int64_t next = (i == m) ? nnz + 1 : rp[i];
    *c0 = cs[i - 1];
    *nc = next - rp[i - 1];
}

/* K5 refinement_matrix_array_mul_kernel -- src/refinement_matrix.jl:365-403; get_row_extends
 * src/utils.jl:204-235.  refmat_of_dim[d] = index into the matrix arrays or -1 (unrefined). */
void FN(orc_refmat_mul)(T *Y, const T *B, int ndims, const int64_t *sizeY, const int64_t *sizeB,
                        const int *refmat_of_dim, const int32_t *const *rp, const int32_t *const *cs,
                        const T *const *nz, const int64_t *nnz)
{
    int64_t total = 1, strideB[ORC_MAXD], sb = 1;
    for (int d = 0; d < ndims; ++d) { total *= sizeY[d]; strideB[d] = sb; sb *= sizeB[d]; }
#pragma omp parallel for schedule(static)
    for (int64_t lin = 0; lin < total; ++lin) {
        int64_t I[ORC_MAXD], c0[ORC_MAXD], nc[ORC_MAXD], r = lin, nterm = 1;
        for (int d = 0; d < ndims; ++d) {
            I[d] = r % sizeY[d] + 1;
            r /= sizeY[d];
            int a = refmat_of_dim[d];
            if (a < 0) { c0[d] = I[d]; nc[d] = 1; }
            else FN(orc_colrange)(rp[a], cs[a], sizeY[d], nnz[a], I[d], &c0[d], &nc[d]);
            nterm *= nc[d];
        }
        T out = 0;
        int64_t Jb[ORC_MAXD] = {0};
        for (int64_t q = 0; q < nterm; ++q) {
            int64_t off = 0;
            for (int d = 0; d < ndims; ++d) off += (Jb[d] + c0[d] - 1) * strideB[d];
            T contrib = B[off];
            for (int d = 0; d < ndims; ++d) {
                int a = refmat_of_dim[d];
                if (a >= 0) contrib *= nz[a][rp[a][I[d] - 1] + Jb[d] - 1];
            }
            out += contrib;
            for (int d = 0; d < ndims; ++d) { if (++Jb[d] < nc[d]) break; Jb[d] = 0; }
        }
        Y[lin] = out;
    }
}

/* K6 refinement_matrix_array_mul_adjoint_kernel + B .= 0 -- src/adjoint.jl:85-152 (float path) */
void FN(orc_refmat_mul_adjoint)(T *B, const T *Y, int ndims, const int64_t *sizeY, const int64_t *sizeB,
                                const int *refmat_of_dim, const int32_t *const *rp,
                                const int32_t *const *cs, const T *const *nz, const int64_t *nnz)
{
    int64_t total = 1, totalB = 1, strideB[ORC_MAXD];
    for (int d = 0; d < ndims; ++d) { total *= sizeY[d]; strideB[d] = totalB; totalB *= sizeB[d]; }
    for (int64_t q = 0; q < totalB; ++q) B[q] = 0;
#pragma omp parallel for schedule(static)
    for (int64_t lin = 0; lin < total; ++lin) {
        int64_t I[ORC_MAXD], c0[ORC_MAXD], nc[ORC_MAXD], r = lin, nterm = 1;
        for (int d = 0; d < ndims; ++d) {
            I[d] = r % sizeY[d] + 1;
            r /= sizeY[d];
            int a = refmat_of_dim[d];
            if (a < 0) { c0[d] = I[d]; nc[d] = 1; }
            else FN(orc_colrange)(rp[a], cs[a], sizeY[d], nnz[a], I[d], &c0[d], &nc[d]);
            nterm *= nc[d];
        }
        int64_t Jb[ORC_MAXD] = {0};
        for (int64_t q = 0; q < nterm; ++q) {
            int64_t off = 0;
            for (int d = 0; d < ndims; ++d) off += (Jb[d] + c0[d] - 1) * strideB[d];
            T contrib = Y[lin];
            for (int d = 0; d < ndims; ++d) {
                int a = refmat_of_dim[d];
                if (a >= 0) contrib *= nz[a][rp[a][I[d] - 1] + Jb[d] - 1];
            }
#pragma omp atomic
            B[off] += contrib;
            for (int d = 0; d < ndims; ++d) { if (++Jb[d] < nc[d]) break; Jb[d] = 0; }
        }
    }
}

/* K7 local_refinement_kernel -- src/control_points.jl:296-311.
 * cp (c_1..c_D, nout); idx (n_active, nin) column-major 1-based; vals (n_active, nout). */
void FN(orc_scatter_active)(T *cp, int nin, const int64_t *n_cp, int nout, const int32_t *idx,
                            const T *vals, int64_t n_active)
{
    int64_t cp_total = 1;
    for (int d = 0; d < nin; ++d) cp_total *= n_cp[d];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_active; ++i) {
        int64_t off = 0, s = 1;
        for (int d = 0; d < nin; ++d) { off += (int64_t)(idx[i + n_active * d] - 1) * s; s *= n_cp[d]; }
        for (int o = 0; o < nout; ++o) cp[off + cp_total * o] = vals[i + n_active * o];
    }
}

/* K8 local_refinement_adjoint_kernel -- src/adjoint.jl:154-170 */
void FN(orc_gather_zero_active)(T *vals, T *cp, int nin, const int64_t *n_cp, int nout,
                                const int32_t *idx, int64_t n_active)
{
    int64_t cp_total = 1;
    for (int d = 0; d < nin; ++d) cp_total *= n_cp[d];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_active; ++i) {
        int64_t off = 0, s = 1;
        for (int d = 0; d < nin; ++d) { off += (int64_t)(idx[i + n_active * d] - 1) * s; s *= n_cp[d]; }
        for (int o = 0; o < nout; ++o) {
            vals[i + n_active * o] = cp[off + cp_total * o];
            cp[off + cp_total * o] = 0;
        }
    }
}
