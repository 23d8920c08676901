/* blo_api.c -- CPU oracle (test infrastructure): the BLU object wrapper, get_factors,
 * maxvolume and scalar getters.  Follows /root/reference/src/{blu,get_factors,maxvolume}.rs. */
#include "blo_int.h"

/* blu.rs:61-70 */
blo *blo_new(lint m, lint b_nz) {
    if (m < 0 || b_nz < 0) return NULL;
    blo *o = calloc(1, sizeof *o);
    if (!o) return NULL;
    blo_lu_init(&o->lu, m, b_nz);
    o->lhs = calloc((size_t)m + 1, sizeof(double));
    o->ilhs = calloc((size_t)m + 1, sizeof(lint));
    o->nzlhs = 0;
    o->realloc_factor = 1.5;
    return o;
}

void blo_free(blo *o) {
    if (!o) return;
    blo_lu_release(&o->lu);
    free(o->lhs);
    free(o->ilhs);
    free(o);
}

/* blu.rs:337-377 */
static void realloc_ix(lint old, lint nz, lint **a_i, double **a_x) {
    *a_i = realloc(*a_i, (size_t)nz * sizeof(lint));
    *a_x = realloc(*a_x, (size_t)nz * sizeof(double));
    if (!*a_i || !*a_x) abort();
    for (lint k = old; k < nz; k++) { (*a_i)[k] = 0; (*a_x)[k] = 0.0; } /* Vec::resize zero-fills */
}

static void realloc_obj(blo *o) {
    blo_lu *lu = &o->lu;
    double f = fmax(1.0, o->realloc_factor);
    if (lu->addmem_l > 0) {
        lint nelem = (lint)((double)(lu->l_mem + lu->addmem_l) * f);
        realloc_ix(lu->l_mem, nelem, &lu->l_index, &lu->l_value);
        lu->l_mem = nelem;
    }
    if (lu->addmem_u > 0) {
        lint nelem = (lint)((double)(lu->u_mem + lu->addmem_u) * f);
        realloc_ix(lu->u_mem, nelem, &lu->u_index, &lu->u_value);
        lu->u_mem = nelem;
    }
    if (lu->addmem_w > 0) {
        lint nelem = (lint)((double)(lu->w_mem + lu->addmem_w) * f);
        realloc_ix(lu->w_mem, nelem, &lu->w_index, &lu->w_value);
        lu->w_mem = nelem;
    }
}

/* blu.rs:380-395 */
static void clear_lhs(blo *o) {
    lint m = o->lu.m;
    lint nzsparse = (lint)(o->lu.sparse_thres * (double)m);
    lint nz = o->nzlhs;
    if (nz) {
        if (nz <= nzsparse) {
            for (lint p = 0; p < nz; p++) o->lhs[o->ilhs[p]] = 0.0;
        } else {
            memset(o->lhs, 0, (size_t)m * sizeof(double));
        }
        o->nzlhs = 0;
    }
}

/* blu.rs:95-118 */
int blo_factorize(blo *o, const lint *b_begin, const lint *b_end, const lint *b_i, const double *b_x) {
    int c0ntinue = 0, st;
    for (;;) {
        st = blo_lu_factorize(&o->lu, b_begin, b_end, b_i, b_x, c0ntinue);
        if (st != BLO_REALLOCATE) break;
        realloc_obj(o);
        c0ntinue = 1;
    }
    return st;
}

/* blu.rs:182 */
int blo_solve_dense(blo *o, const double *rhs, double *lhs, char trans) {
    return blo_lu_solve_dense(&o->lu, rhs, lhs, trans);
}

/* blu.rs:207-226 */
int blo_solve_sparse(blo *o, lint nzrhs, const lint *irhs, const double *xrhs, char trans) {
    clear_lhs(o);
    return blo_lu_solve_sparse(&o->lu, nzrhs, irhs, xrhs, &o->nzlhs, o->ilhs, o->lhs, trans);
}

/* blu.rs:257-294 */
int blo_solve_for_update(blo *o, lint nzrhs, const lint *irhs, const double *xrhs,
                         char trans, lint want_solution) {
    int st;
    clear_lhs(o);
    for (;;) {
        lint nzlhs = 0;
        /* D13 repaired: blu.rs:268-283 always passes Some(ilhs)/Some(lhs), so with
         * want_solution == 0 the solution is still scattered into lhs while nzlhs is not
         * recorded and lu_clear_lhs (blu.rs:380-395) later clears the wrong entries.
         * BASICLU semantics: no solution unless it was asked for. */
#if BLO_REPAIR_D13
        if (want_solution) {
            st = blo_lu_solve_for_update(&o->lu, nzrhs, irhs, xrhs, &nzlhs, o->ilhs, o->lhs, trans);
            o->nzlhs = nzlhs;
        } else {
            st = blo_lu_solve_for_update(&o->lu, nzrhs, irhs, xrhs, NULL, NULL, NULL, trans);
        }
#else
        st = blo_lu_solve_for_update(&o->lu, nzrhs, irhs, xrhs, &nzlhs, o->ilhs, o->lhs, trans);   /* blu.rs:268-283 as written */
        if (want_solution) o->nzlhs = nzlhs;
#endif
        if (st != BLO_REALLOCATE) break;
        realloc_obj(o);
    }
    return st;
}

/* blu.rs:319-334 */
int blo_update(blo *o, double xtbl) {
    int st;
    for (;;) {
        st = blo_lu_update(&o->lu, xtbl);
        if (st != BLO_REALLOCATE) break;
        realloc_obj(o);
    }
    return st;
}

/* get_factors.rs:48-180 */
int blo_lu_get_factors(blo_lu *lu, lint *rowperm, lint *colperm,
                       lint *l_colptr, lint *l_rowidx, double *l_value_,
                       lint *u_colptr, lint *u_rowidx, double *u_value_) {
#if !BLO_REPAIR_D9
    if (lu->nupdate < 0) BLO_DEFECT_TRAP("D9", "get_factors.rs:59 unwraps nupdate == None (panic)");
#endif
    if (lu->nupdate != 0) return BLO_ERROR_INVALID_CALL; /* D9 repaired: also covers "never factorized" */
    const lint m = lu->m;
    if (rowperm) memcpy(rowperm, lu->pivotrow, (size_t)m * sizeof(lint));
    if (colperm) memcpy(colperm, lu->pivotcol, (size_t)m * sizeof(lint));

    if (l_colptr && l_rowidx && l_value_) {
        lint *colptr = lu->iwork1;
        lint put = 0;
        for (lint k = 0; k < m; k++) {
            l_colptr[k] = put;
            l_rowidx[put] = k;
            l_value_[put++] = 1.0;
            colptr[lu->p[k]] = put;
            put += lu->l_begin_p[k + 1] - lu->l_begin_p[k] - 1;
        }
        l_colptr[m] = put;
        assert(put == lu->l_nz + m);
        for (lint k = 0; k < m; k++)
            for (lint pos = lu->lt_begin_p[k]; lu->l_index[pos] >= 0; pos++) {
                lint dst = colptr[lu->l_index[pos]]++;
                l_rowidx[dst] = k;
                l_value_[dst] = lu->l_value[pos];
            }
    }
    if (u_colptr && u_rowidx && u_value_) {
        lint *colptr = lu->iwork1;
        memset(colptr, 0, (size_t)m * sizeof(lint));
        for (lint j = 0; j < m; j++)
            for (lint pos = lu->w_begin[j]; pos < lu->w_end[j]; pos++) colptr[lu->w_index[pos]]++;
        lint put = 0;
        for (lint k = 0; k < m; k++) {
            lint j = lu->pivotcol[k];
            u_colptr[k] = put;
            put += colptr[j];
            colptr[j] = u_colptr[k];
            u_rowidx[put] = k;
            u_value_[put++] = lu->col_pivot[j];
        }
        u_colptr[m] = put;
        assert(put == lu->u_nz + m);
        for (lint k = 0; k < m; k++) {
            lint j = lu->pivotcol[k];
            for (lint pos = lu->w_begin[j]; pos < lu->w_end[j]; pos++) {
                lint dst = colptr[lu->w_index[pos]]++;
                u_rowidx[dst] = k;
                u_value_[dst] = lu->w_value[pos];
            }
        }
    }
    return BLO_OK;
}

int blo_get_factors(blo *o, lint *rowperm, lint *colperm,
                    lint *l_colptr, lint *l_rowidx, double *l_value,
                    lint *u_colptr, lint *u_rowidx, double *u_value) {
    return blo_lu_get_factors(&o->lu, rowperm, colperm, l_colptr, l_rowidx, l_value,
                              u_colptr, u_rowidx, u_value);
}

/* maxvolume.rs:180-224 */
static int mv_factorize(blo *o, const lint *a_p, const lint *a_i, const double *a_x, const lint *basis) {
    lint m = o->lu.m;
    lint *begin = malloc((size_t)(m + 1) * sizeof(lint)), *end = malloc((size_t)(m + 1) * sizeof(lint));
    for (lint i = 0; i < m; i++) {
        begin[i] = a_p[basis[i]];
        end[i] = a_p[basis[i] + 1];
    }
    int st = blo_factorize(o, begin, end, a_i, a_x);
    free(begin);
    free(end);
    return st;
}

/* maxvolume.rs:64-177 */
int blo_maxvolume(blo *o, lint ncol, const lint *a_p, const lint *a_i, const double *a_x,
                  lint *basis, lint *isbasic, double volumetol, lint *p_nupdate) {
    lint nupdate = 0;
    int st = BLO_OK;
    if (volumetol < 1.0) { st = BLO_ERROR_INVALID_ARGUMENT; goto cleanup; }
    st = mv_factorize(o, a_p, a_i, a_x, basis);
    if (st != BLO_OK) goto cleanup;
    for (lint j = 0; j < ncol; j++) {
        if (isbasic[j]) continue;
        lint nzrhs = a_p[j + 1] - a_p[j];
        st = blo_solve_for_update(o, nzrhs, a_i + a_p[j], a_x + a_p[j], 'N', 1);
        if (st != BLO_OK) goto cleanup;
        double xmax = 0.0, xtbl = 0.0;
        lint imax = 0;
        for (lint k = 0; k < o->nzlhs; k++) {
            lint i = o->ilhs[k];
            if (fabs(o->lhs[i]) > xmax) {
                xtbl = o->lhs[i];
                xmax = fabs(xtbl);
                imax = i;
            }
        }
        if (xmax <= volumetol) continue;
        isbasic[basis[imax]] = 0;
        isbasic[j] = 1;
        basis[imax] = j;
        nupdate++;
        st = blo_solve_for_update(o, 0, &imax, NULL, 'T', 0);
        if (st != BLO_OK) goto cleanup;
        st = blo_update(o, xtbl);
        if (st != BLO_OK) goto cleanup;
        if (o->lu.nforrest == o->lu.m || o->lu.pivot_error > 1e-8 || blo_lu_update_cost(&o->lu) > 1.0) {
            st = mv_factorize(o, a_p, a_i, a_x, basis);
            if (st != BLO_OK) goto cleanup;
        }
    }
cleanup:
    if (p_nupdate) *p_nupdate = nupdate;
    return st;
}

/* ---- oracle-only conveniences ---- */
void blo_trace_enable(blo *o, int on) { o->lu.trace_on = on; }
lint blo_trace_len(const blo *o) { return o->lu.trace_len; }
const blo_trace *blo_trace_data(const blo *o) { return o->lu.trace; }
const double *blo_lhs(const blo *o) { return o->lhs; }
const lint *blo_ilhs(const blo *o) { return o->ilhs; }

void blo_set_param(blo *o, int what, double v) {
    blo_lu *lu = &o->lu;
    switch (what) {
    case BLO_P_DROPTOL: lu->droptol = v; break;
    case BLO_P_ABSTOL: lu->abstol = v; break;
    case BLO_P_RELTOL: lu->reltol = v; break;
    case BLO_P_NZBIAS: lu->nzbias = (lint)v; break;
    case BLO_P_MAXSEARCH: lu->maxsearch = (lint)v; break;
    case BLO_P_PAD: lu->pad = (lint)v; break;
    case BLO_P_STRETCH: lu->stretch = v; break;
    case BLO_P_COMPRESS_THRES: lu->compress_thres = v; break;
    case BLO_P_SPARSE_THRES: lu->sparse_thres = v; break;
    case BLO_P_SEARCH_ROWS: lu->search_rows = (lint)v; break;
    case BLO_P_CHECK_FILE_DIFF: lu->check_file_diff = (int)v; break;
    case BLO_P_REALLOC_FACTOR: o->realloc_factor = v; break;
    default: break;
    }
}

double blo_get_info(const blo *o, int what) {
    const blo_lu *lu = &o->lu;
    switch (what) {
    case BLO_P_DROPTOL: return lu->droptol;
    case BLO_P_ABSTOL: return lu->abstol;
    case BLO_P_RELTOL: return lu->reltol;
    case BLO_P_NZBIAS: return (double)lu->nzbias;
    case BLO_P_MAXSEARCH: return (double)lu->maxsearch;
    case BLO_P_PAD: return (double)lu->pad;
    case BLO_P_STRETCH: return lu->stretch;
    case BLO_P_COMPRESS_THRES: return lu->compress_thres;
    case BLO_P_SPARSE_THRES: return lu->sparse_thres;
    case BLO_P_SEARCH_ROWS: return (double)lu->search_rows;
    case BLO_P_CHECK_FILE_DIFF: return (double)lu->check_file_diff;
    case BLO_P_REALLOC_FACTOR: return o->realloc_factor;
    case BLO_I_M: return (double)lu->m;
    case BLO_I_RANK: return (double)lu->rank;
    case BLO_I_BUMP_SIZE: return (double)lu->bump_size;
    case BLO_I_BUMP_NZ: return (double)lu->bump_nz;
    case BLO_I_MATRIX_NZ: return (double)lu->matrix_nz;
    case BLO_I_L_NZ: return (double)lu->l_nz;
    case BLO_I_U_NZ: return (double)lu->u_nz;
    case BLO_I_R_NZ: return (double)lu->r_nz;
    case BLO_I_NSEARCH_PIVOT: return (double)lu->nsearch_pivot;
    case BLO_I_NEXPAND: return (double)lu->nexpand;
    case BLO_I_NGARBAGE: return (double)lu->ngarbage;
    case BLO_I_FACTOR_FLOPS: return (double)lu->factor_flops;
    case BLO_I_MIN_PIVOT: return lu->min_pivot;
    case BLO_I_MAX_PIVOT: return lu->max_pivot;
    case BLO_I_MAX_ETA: return lu->max_eta;
    case BLO_I_NUPDATE: return (double)lu->nupdate;
    case BLO_I_NFORREST: return (double)lu->nforrest;
    case BLO_I_NFACTORIZE: return (double)lu->nfactorize;
    case BLO_I_NUPDATE_TOTAL: return (double)lu->nupdate_total;
    case BLO_I_NFORREST_TOTAL: return (double)lu->nforrest_total;
    case BLO_I_NSYMPERM_TOTAL: return (double)lu->nsymperm_total;
    case BLO_I_L_FLOPS: return (double)lu->l_flops;
    case BLO_I_U_FLOPS: return (double)lu->u_flops;
    case BLO_I_R_FLOPS: return (double)lu->r_flops;
    case BLO_I_CONDEST_L: return lu->condest_l;
    case BLO_I_CONDEST_U: return lu->condest_u;
    case BLO_I_NORM_L: return lu->norm_l;
    case BLO_I_NORM_U: return lu->norm_u;
    case BLO_I_NORMEST_L_INV: return lu->normest_l_inv;
    case BLO_I_NORMEST_U_INV: return lu->normest_u_inv;
    case BLO_I_ONENORM: return lu->onenorm;
    case BLO_I_INFNORM: return lu->infnorm;
    case BLO_I_RESIDUAL_TEST: return lu->residual_test;
    case BLO_I_PIVOT_ERROR: return lu->pivot_error;
    case BLO_I_UPDATE_COST: return blo_lu_update_cost(lu);
    case BLO_I_TIME_FACTORIZE: return lu->time_factorize;
    case BLO_I_TIME_SOLVE: return lu->time_solve;
    case BLO_I_TIME_UPDATE: return lu->time_update;
    case BLO_I_TIME_SINGLETONS: return lu->time_singletons;
    case BLO_I_TIME_SEARCH_PIVOT: return lu->time_search_pivot;
    case BLO_I_TIME_ELIM_PIVOT: return lu->time_elim_pivot;
    case BLO_I_L_MEM: return (double)lu->l_mem;
    case BLO_I_U_MEM: return (double)lu->u_mem;
    case BLO_I_W_MEM: return (double)lu->w_mem;
    case BLO_I_NZLHS: return (double)o->nzlhs;
    case BLO_I_ELIM_BYTES: return lu->elim_bytes;
    case BLO_I_NELIM_DIV: return (double)lu->nelim_div;
    case BLO_I_PIVOTLEN: return (double)lu->pivotlen;
    case BLO_I_RANKDEF: return (double)lu->rankdef;
    default: return NAN;
    }
}
