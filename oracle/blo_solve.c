/* blo_solve.c -- CPU oracle (test infrastructure): dense and sparse solves.
 * Follows /root/reference/src/lu/{garbage_perm,solve_dense,dfs,solve_symbolic,
 * solve_triangular,solve_sparse,solve_for_update}.rs. */
#include "blo_int.h"

static int is_trans(char t) { return t == 't' || t == 'T'; }

/* garbage_perm.rs:16-48 */
void blo_garbage_perm(blo_lu *lu) {
    const lint m = lu->m, pivotlen = lu->pivotlen;
    lint *pivotcol = lu->pivotcol, *pivotrow = lu->pivotrow, *marked = lu->iwork0;
    if (pivotlen > m) {
        lint marker = ++lu->marker;
        lint put = pivotlen;
        for (lint get = pivotlen - 1; get >= 0; get--) {
            lint j = pivotcol[get];
            if (marked[j] != marker) {
                marked[j] = marker;
                --put;
                pivotcol[put] = j;
                pivotrow[put] = pivotrow[get];
            }
        }
        assert(put + m == pivotlen);
        memmove(pivotcol, pivotcol + put, (size_t)m * sizeof(lint));
        memmove(pivotrow, pivotrow + put, (size_t)m * sizeof(lint));
        lu->pivotlen = m;
    }
}

/* lu/solve_dense.rs:7-120 */
void blo_k_solve_dense(blo_lu *lu, const double *rhs, double *lhs, char trans) {
    blo_garbage_perm(lu);
    assert(lu->pivotlen == lu->m);
    const lint m = lu->m, nforrest = lu->nforrest;
    const lint *p = lu->p, *eta_row = lu->eta_row, *pivotcol = lu->pivotcol, *pivotrow = lu->pivotrow;
    const lint *l_begin_p = lu->l_begin_p, *lt_begin_p = lu->lt_begin_p, *u_begin = lu->u_begin;
    const lint *r_begin = lu->r_begin, *w_begin = lu->w_begin, *w_end = lu->w_end;
    const double *col_pivot = lu->col_pivot, *row_pivot = lu->row_pivot;
    const lint *l_index = lu->l_index, *u_index = lu->u_index, *w_index = lu->w_index;
    const double *l_value = lu->l_value, *u_value = lu->u_value, *w_value = lu->w_value;
    double *work1 = lu->work1;

    if (is_trans(trans)) {
        memcpy(work1, rhs, (size_t)m * sizeof(double));
        /* U' */
        for (lint k = 0; k < m; k++) {
            lint jpivot = pivotcol[k], ipivot = pivotrow[k];
            double x = work1[jpivot] / col_pivot[jpivot];
            for (lint pos = w_begin[jpivot]; pos < w_end[jpivot]; pos++)
                work1[w_index[pos]] -= x * w_value[pos];
            lhs[ipivot] = x;
        }
        /* etas backwards */
        for (lint t = nforrest - 1; t >= 0; t--) {
            double x = lhs[eta_row[t]];
            for (lint pos = r_begin[t]; pos < r_begin[t + 1]; pos++)
                lhs[l_index[pos]] -= x * l_value[pos];
        }
        /* L' */
        for (lint k = m - 1; k >= 0; k--) {
            double x = 0.0;
            for (lint pos = l_begin_p[k]; l_index[pos] >= 0; pos++)
                x += lhs[l_index[pos]] * l_value[pos];
            lhs[p[k]] -= x;
        }
    } else {
        memcpy(work1, rhs, (size_t)m * sizeof(double));
        /* L */
        for (lint k = 0; k < m; k++) {
            double x = 0.0;
            for (lint pos = lt_begin_p[k]; l_index[pos] >= 0; pos++)
                x += work1[l_index[pos]] * l_value[pos];
            work1[p[k]] -= x;
        }
        /* etas */
        lint pos = r_begin[0];
        for (lint t = 0; t < nforrest; t++) {
            double x = 0.0;
            for (; pos < r_begin[t + 1]; pos++) x += work1[l_index[pos]] * l_value[pos];
            work1[eta_row[t]] -= x;
        }
        /* U */
        for (lint k = m - 1; k >= 0; k--) {
            lint jpivot = pivotcol[k], ipivot = pivotrow[k];
            double x = work1[ipivot] / row_pivot[ipivot];
            for (lint q = u_begin[ipivot]; u_index[q] >= 0; q++)
                work1[u_index[q]] -= x * u_value[q];
            lhs[jpivot] = x;
        }
    }
}

/* dfs.rs:25-145.  end == NULL: neighbour lists are terminated by a negative index. */
lint blo_dfs(lint i, const lint *begin, const lint *end, const lint *index, lint top,
             lint *xi, lint *pstack, lint *marked, lint marker) {
    if (marked[i] == marker) return top;
    lint head = 0;
    xi[0] = i;
    while (head >= 0) {
        i = xi[head];
        if (marked[i] != marker) {
            marked[i] = marker;
            pstack[head] = begin[i];
        }
        int done = 1;
        if (end) {
            for (lint p = pstack[head]; p < end[i]; p++) {
                lint inext = index[p];
                if (marked[inext] == marker) continue;
                pstack[head] = p + 1;
                xi[++head] = inext;
                done = 0;
                break;
            }
        } else {
            lint inext;
            for (lint p = pstack[head]; (inext = index[p]) >= 0; p++) {
                if (marked[inext] == marker) continue;
                pstack[head] = p + 1;
                xi[++head] = inext;
                done = 0;
                break;
            }
        }
        if (done) {
            head--;
            xi[--top] = i;
        }
    }
    return top;
}

/* solve_symbolic.rs:19-40 */
lint blo_solve_symbolic(lint m, const lint *begin, const lint *end, const lint *index,
                        lint nrhs, const lint *irhs, lint *ilhs, lint *pstack,
                        lint *marked, lint marker) {
    lint top = m;
    for (lint n = 0; n < nrhs; n++)
        if (marked[irhs[n]] != marker)
            top = blo_dfs(irhs[n], begin, end, index, top, ilhs, pstack, marked, marker);
    return top;
}

/* solve_triangular.rs:27-136 (the four specialisations folded into one loop) */
lint blo_solve_triangular(lint nz_symb, const lint *pattern_symb, const lint *begin, const lint *end,
                          const lint *index, const double *value, const double *pivot,
                          double droptol, double *lhs, lint *pattern, lint *flops) {
    lint nz = 0, flop_count = 0;
    for (lint n = 0; n < nz_symb; n++) {
        lint ipivot = pattern_symb[n];
        if (lhs[ipivot] != 0.0) {
            double x;
            if (pivot) {
                lhs[ipivot] /= pivot[ipivot];
                flop_count++;
            }
            x = lhs[ipivot];
            if (end) {
                for (lint pos = begin[ipivot]; pos < end[ipivot]; pos++) {
                    lhs[index[pos]] -= x * value[pos];
                    flop_count++;
                }
            } else {
                for (lint pos = begin[ipivot]; index[pos] >= 0; pos++) {
                    lhs[index[pos]] -= x * value[pos];
                    flop_count++;
                }
            }
            if (fabs(x) > droptol) pattern[nz++] = ipivot;
            else lhs[ipivot] = 0.0;
        }
    }
    *flops += flop_count;
    return nz;
}

/* lu/solve_sparse.rs:242-258 and lu/solve_for_update.rs:303-319 */
static void unmark_cancellation(lint m, lint top, lint nz, lint nz_symb,
                                const lint *pattern_symb, const lint *pattern, lint *marked) {
    if (nz < nz_symb) {
        lint t = top, n = 0;
        while (n < nz) {
            lint i = pattern_symb[t];
            if (i == pattern[n]) n++;
            else marked[i]--;
            t++;
        }
        for (; t < m; t++) marked[pattern_symb[t]]--;
    }
}

/* second half of a transposed solve: etas backwards then L'.
 * lu/solve_sparse.rs:113-192, lu/solve_for_update.rs:166-245 */
static lint btran_tail(blo_lu *lu, lint nz, lint *pattern, lint *pattern_symb, lint marker,
                       lint *ilhs, double *xlhs, lint *l_flops, lint *r_flops) {
    const lint m = lu->m, nforrest = lu->nforrest;
    const lint nz_sparse = (lint)(lu->sparse_thres * (double)m);
    lint *marked = lu->iwork0;
    const lint *l_index = lu->l_index;
    const double *l_value = lu->l_value;

    for (lint t = nforrest - 1; t >= 0; t--) {
        lint ipivot = lu->eta_row[t];
        if (xlhs[ipivot] != 0.0) {
            double x = xlhs[ipivot];
            for (lint pos = lu->r_begin[t]; pos < lu->r_begin[t + 1]; pos++) {
                lint i = l_index[pos];
                if (marked[i] != marker) {
                    marked[i] = marker;
                    pattern[nz++] = i;
                }
                xlhs[i] -= x * l_value[pos];
                (*r_flops)++;
            }
        }
    }
    if (nz <= nz_sparse) {
        lint mk = ++lu->marker;
        lint top = blo_solve_symbolic(m, lu->lt_begin, NULL, l_index, nz, pattern, pattern_symb,
                                      lu->pstack, marked, mk);
        nz = blo_solve_triangular(m - top, pattern_symb + top, lu->lt_begin, NULL, l_index, l_value,
                                  NULL, lu->droptol, xlhs, ilhs, l_flops);
    } else {
        nz = 0;
        for (lint k = m - 1; k >= 0; k--) {
            lint ipivot = lu->p[k];
            if (xlhs[ipivot] != 0.0) {
                double x = xlhs[ipivot];
                for (lint pos = lu->lt_begin_p[k]; l_index[pos] >= 0; pos++) {
                    xlhs[l_index[pos]] -= x * l_value[pos];
                    (*l_flops)++;
                }
                if (fabs(x) > lu->droptol) ilhs[nz++] = ipivot;
                else xlhs[ipivot] = 0.0;
            }
        }
    }
    return nz;
}

/* second half of a forward solve: U.  lu/solve_sparse.rs:279-349, lu/solve_for_update.rs:363-433 */
static lint ftran_tail(blo_lu *lu, lint nz, lint *pattern, lint *pattern_symb,
                       lint *ilhs, double *xlhs, lint *u_flops) {
    const lint m = lu->m;
    const lint nz_sparse = (lint)(lu->sparse_thres * (double)m);
    lint *marked = lu->iwork0;
    double *work = lu->work0;
    const lint *u_index = lu->u_index;
    const double *u_value = lu->u_value;

    if (nz <= nz_sparse) {
        lint mk = ++lu->marker;
        lint top = blo_solve_symbolic(m, lu->u_begin, NULL, u_index, nz, pattern, pattern_symb,
                                      lu->pstack, marked, mk);
        nz = blo_solve_triangular(m - top, pattern_symb + top, lu->u_begin, NULL, u_index, u_value,
                                  lu->row_pivot, lu->droptol, work, ilhs, u_flops);
        for (lint n = 0; n < nz; n++) {
            lint i = ilhs[n], j = lu->qinv[i]; /* qmap */
            ilhs[n] = j;
            xlhs[j] = work[i];
            work[i] = 0.0;
        }
    } else {
        nz = 0;
        for (lint k = lu->pivotlen - 1; k >= 0; k--) {
            lint ipivot = lu->pivotrow[k], jpivot = lu->pivotcol[k];
            if (work[ipivot] != 0.0) {
                double x = work[ipivot] / lu->row_pivot[ipivot];
                work[ipivot] = 0.0;
                for (lint pos = lu->u_begin[ipivot]; u_index[pos] >= 0; pos++) {
                    work[u_index[pos]] -= x * u_value[pos];
                    (*u_flops)++;
                }
                if (fabs(x) > lu->droptol) {
                    ilhs[nz++] = jpivot;
                    xlhs[jpivot] = x;
                }
            }
        }
    }
    return nz;
}

/* first half of a forward solve: L then etas.  lu/solve_sparse.rs:196-277,
 * lu/solve_for_update.rs:246-326.  Returns nz; result scattered in work0, indices in pattern. */
static lint ftran_head(blo_lu *lu, lint nrhs, const lint *irhs, const double *xrhs,
                       lint *pattern, lint *pattern_symb, lint *l_flops, lint *r_flops) {
    const lint m = lu->m, nforrest = lu->nforrest;
    lint *marked = lu->iwork0;
    double *work = lu->work0;
    const lint *l_index = lu->l_index;
    const double *l_value = lu->l_value;

    lint marker = ++lu->marker;
    lint top = blo_solve_symbolic(m, lu->l_begin, NULL, l_index, nrhs, irhs, pattern_symb,
                                  lu->pstack, marked, marker);
    lint nz_symb = m - top;
    for (lint n = 0; n < nrhs; n++) work[irhs[n]] = xrhs[n];
    lint nz = blo_solve_triangular(nz_symb, pattern_symb + top, lu->l_begin, NULL, l_index, l_value,
                                   NULL, lu->droptol, work, pattern, l_flops);
    unmark_cancellation(m, top, nz, nz_symb, pattern_symb, pattern, marked);

    lint pos = lu->r_begin[0];
    for (lint t = 0; t < nforrest; t++) {
        lint ipivot = lu->eta_row[t];
        double x = 0.0;
        for (; pos < lu->r_begin[t + 1]; pos++) x += work[l_index[pos]] * l_value[pos];
        work[ipivot] -= x;
        if (x != 0.0 && marked[ipivot] != marker) {
            marked[ipivot] = marker;
            pattern[nz++] = ipivot;
        }
    }
    *r_flops += lu->r_begin[nforrest] - lu->r_begin[0];
    return nz;
}

static void solve_done(blo_lu *lu, double tic, lint l_flops, lint u_flops, lint r_flops) {
    double el = blo_now() - tic;
    lu->time_solve += el;
    lu->time_solve_total += el;
    lu->l_flops += l_flops;
    lu->u_flops += u_flops;
    lu->r_flops += r_flops;
    lu->update_cost_numer += (double)r_flops;
}

/* lu/solve_sparse.rs:11-360 */
void blo_k_solve_sparse(blo_lu *lu, lint nrhs, const lint *irhs, const double *xrhs,
                        lint *p_nlhs, lint *ilhs, double *xlhs, char trans) {
    const lint m = lu->m;
    lint *pattern_symb = lu->iwork1, *pattern = lu->iwork1 + m;
    lint *marked = lu->iwork0;
    double *work = lu->work0;
    lint l_flops = 0, u_flops = 0, r_flops = 0;
    double tic = blo_now();

    if (is_trans(trans)) {
        /* U' sparse, lu/solve_sparse.rs:68-111 */
        lint marker = ++lu->marker;
        lint top = blo_solve_symbolic(m, lu->w_begin, lu->w_end, lu->w_index, nrhs, irhs,
                                      pattern_symb, lu->pstack, marked, marker);
        for (lint n = 0; n < nrhs; n++) work[irhs[n]] = xrhs[n];
        lint nz = blo_solve_triangular(m - top, pattern_symb + top, lu->w_begin, lu->w_end,
                                       lu->w_index, lu->w_value, lu->col_pivot, lu->droptol,
                                       work, pattern, &u_flops);
        marker = ++lu->marker;
        for (lint n = 0; n < nz; n++) {
            lint j = pattern[n], i = lu->pinv[j]; /* pmap */
            pattern[n] = i;
            xlhs[i] = work[j];
            work[j] = 0.0;
            marked[i] = marker;
        }
        *p_nlhs = btran_tail(lu, nz, pattern, pattern_symb, marker, ilhs, xlhs, &l_flops, &r_flops);
    } else {
        lint nz = ftran_head(lu, nrhs, irhs, xrhs, pattern, pattern_symb, &l_flops, &r_flops);
        *p_nlhs = ftran_tail(lu, nz, pattern, pattern_symb, ilhs, xlhs, &u_flops);
    }
    solve_done(lu, tic, l_flops, u_flops, r_flops);
}

/* lu/solve_for_update.rs:12-455 */
int blo_k_solve_for_update(blo_lu *lu, lint nrhs, const lint *irhs, const double *xrhs,
                           lint *p_nlhs, lint *ilhs, double *xlhs, char trans) {
    const lint m = lu->m, nforrest = lu->nforrest;
    lint *pattern_symb = lu->iwork1, *pattern = lu->iwork1 + m;
    lint *marked = lu->iwork0;
    double *work = lu->work0;
    const int want_solution = p_nlhs && ilhs && xlhs;
    lint l_flops = 0, u_flops = 0, r_flops = 0;
    double tic = blo_now();

    if (is_trans(trans)) {
        const lint jpivot = irhs[0];
        const lint ipivot = lu->pinv[jpivot]; /* pmap */
        const lint jbegin = lu->w_begin[jpivot], jend = lu->w_end[jpivot];

        /* row eta: symbolic + numeric U' solve seeded with row ipivot of U, no dropping.
         * lu/solve_for_update.rs:70-120 */
        lint marker = ++lu->marker;
        lint top = blo_solve_symbolic(m, lu->w_begin, lu->w_end, lu->w_index, jend - jbegin,
                                      lu->w_index + jbegin, pattern_symb, lu->pstack, marked, marker);
        lint nz_symb = m - top;
        lint room = lu->l_mem - lu->r_begin[nforrest];
        if (room < nz_symb) { lu->addmem_l = nz_symb - room; return BLO_REALLOCATE; }
        for (lint pos = jbegin; pos < jend; pos++) work[lu->w_index[pos]] = lu->w_value[pos];
        blo_solve_triangular(nz_symb, pattern_symb + top, lu->w_begin, lu->w_end, lu->w_index,
                             lu->w_value, lu->col_pivot, 0.0, work, pattern, &u_flops);

        /* store the symbolic pattern with values as the row eta, :124-135 */
        lint put = lu->r_begin[nforrest];
        for (lint t = top; t < m; t++) {
            lint j = pattern_symb[t];
            lu->l_index[put] = lu->pinv[j];
            lu->l_value[put] = work[j];
            put++;
            work[j] = 0.0;
        }
        lu->r_begin[nforrest + 1] = put;
#if BLO_REPAIR_D1
        lu->eta_row[nforrest] = ipivot; /* D1 repaired: own array */
#else
        lu->r_begin[nforrest] = ipivot; /* lu.rs:184-193 as written: eta_row! expands to the r_begin storage */
#endif
        lu->btran_for_update = jpivot;

        if (!want_solution) { solve_done(lu, tic, l_flops, u_flops, r_flops); return BLO_OK; }

        /* scale to U^{-T} e_j, :144-165 */
        marker = ++lu->marker;
        pattern[0] = ipivot;
        marked[ipivot] = marker;
        double pivot = lu->col_pivot[jpivot];
        xlhs[ipivot] = 1.0 / pivot;
        double xdrop = lu->droptol * fabs(pivot);
        lint nz = 1;
        for (lint pos = lu->r_begin[nforrest]; pos < lu->r_begin[nforrest + 1]; pos++) {
            if (fabs(lu->l_value[pos]) > xdrop) {
                lint i = lu->l_index[pos];
                pattern[nz++] = i;
                marked[i] = marker;
                xlhs[i] = -lu->l_value[pos] / pivot;
            }
        }
        *p_nlhs = btran_tail(lu, nz, pattern, pattern_symb, marker, ilhs, xlhs, &l_flops, &r_flops);
    } else {
        lint nz = ftran_head(lu, nrhs, irhs, xrhs, pattern, pattern_symb, &l_flops, &r_flops);

        /* spike into U at u_begin[m], :328-355 */
        lint room = lu->u_mem - lu->u_begin[m];
        lint need = nz + 1;
        if (room < need) {
            for (lint n = 0; n < nz; n++) work[pattern[n]] = 0.0;
            lu->addmem_u = need - room;
            return BLO_REALLOCATE;
        }
        lint put = lu->u_begin[m];
        for (lint n = 0; n < nz; n++) {
            lint i = pattern[n];
            lu->u_index[put] = i;
            lu->u_value[put] = work[i];
            put++;
            if (!want_solution) work[i] = 0.0;
        }
        lu->u_index[put] = -1;
        lu->ftran_for_update = 0;

        if (!want_solution) { solve_done(lu, tic, l_flops, u_flops, r_flops); return BLO_OK; }
        *p_nlhs = ftran_tail(lu, nz, pattern, pattern_symb, ilhs, xlhs, &u_flops);
    }
    solve_done(lu, tic, l_flops, u_flops, r_flops);
    return BLO_OK;
}

/* ---- L2 wrappers with the argument checks ---- */

/* solve_dense.rs:24-32 */
int blo_lu_solve_dense(blo_lu *lu, const double *rhs, double *lhs, char trans) {
    if (lu->nupdate < 0) return BLO_ERROR_INVALID_CALL;
    blo_k_solve_dense(lu, rhs, lhs, trans);
    return BLO_OK;
}

/* solve_sparse.rs:35-73 */
int blo_lu_solve_sparse(blo_lu *lu, lint nzrhs, const lint *irhs, const double *xrhs,
                        lint *p_nzlhs, lint *ilhs, double *lhs, char trans) {
    if (lu->nupdate < 0) return BLO_ERROR_INVALID_CALL;
    int ok = nzrhs >= 0 && nzrhs <= lu->m;
    for (lint n = 0; n < nzrhs && ok; n++) ok = ok && irhs[n] >= 0 && irhs[n] < lu->m;
    if (!ok) return BLO_ERROR_INVALID_ARGUMENT;
    blo_k_solve_sparse(lu, nzrhs, irhs, xrhs, p_nzlhs, ilhs, lhs, trans);
    return BLO_OK;
}

/* solve_for_update.rs:72-119 */
int blo_lu_solve_for_update(blo_lu *lu, lint nzrhs, const lint *irhs, const double *xrhs,
                            lint *p_nzlhs, lint *ilhs, double *lhs, char trans) {
    if (!is_trans(trans) && !xrhs) return BLO_ERROR_ARGUMENT_MISSING;
    if (lu->nupdate < 0) return BLO_ERROR_INVALID_CALL;
    if (lu->nforrest == lu->m) return BLO_ERROR_MAXIMUM_UPDATES;
    int ok;
    if (is_trans(trans)) {
        ok = irhs[0] >= 0 && irhs[0] < lu->m;
    } else {
        ok = nzrhs >= 0 && nzrhs <= lu->m;
        for (lint n = 0; n < nzrhs && ok; n++) ok = ok && irhs[n] >= 0 && irhs[n] < lu->m;
    }
    if (!ok) return BLO_ERROR_INVALID_ARGUMENT;
#if BLO_REPAIR_D7
    /* D7 repair (lu_load semantics) */
    lu->addmem_l = lu->addmem_u = lu->addmem_w = 0;
#endif
    return blo_k_solve_for_update(lu, nzrhs, irhs, xrhs, p_nzlhs, ilhs, lhs, trans);
}
