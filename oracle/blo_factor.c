/* blo_factor.c -- CPU oracle (test infrastructure): the four factorization phases.
 * Follows /root/reference/src/lu/{singletons,setup_bump,markowitz,factorize_bump,
 * build_factors}.rs and src/factorize.rs.  The elimination step (pivot.rs) is in
 * blo_pivot.c. */
#include "blo_int.h"

/* ------------------------------------------------------------------ */
/* Phase 1: singletons.rs                                              */
/* ------------------------------------------------------------------ */

/* singletons.rs:287-393.  iset[j] = XOR of the active row indices of column j
 * (Gilbert's trick); qinv[j] = -(count)-1 doubles as the active-entry counter. */
static lint singleton_cols(lint m, const lint *b_begin, const lint *b_end, const lint *b_i,
                           const lint *b_tp, const lint *b_ti, const double *b_tx,
                           lint *u_p, lint *u_i, double *u_x, lint *l_p, lint *l_i,
                           double *col_pivot, lint *pinv, lint *qinv,
                           lint *iset, lint *queue, lint rank, double abstol, blo_lu *lu) {
    lint rk = rank, tail = 0;
    for (lint j = 0; j < m; j++) {
        if (qinv[j] < 0) {
            lint nz = b_end[j] - b_begin[j];
            lint x = 0;
            for (lint pos = b_begin[j]; pos < b_end[j]; pos++) x ^= b_i[pos];
            iset[j] = x;
            qinv[j] = -nz - 1;
            if (nz == 1) queue[tail++] = j;
        }
    }
    lint put = u_p[rank];
    for (lint front = 0; front < tail; front++) {
        lint j = queue[front];
        assert(qinv[j] == -2 || qinv[j] == -1);
        if (qinv[j] == -1) continue; /* column became empty meanwhile */
        lint i = iset[j];
        assert(i >= 0 && i < m);
        assert(pinv[i] < 0);
        lint end = b_tp[i + 1];
        lint pos = b_tp[i];
        while (b_ti[pos] != j) { assert(pos < end - 1); pos++; }
        double piv = b_tx[pos];
        if (piv == 0.0 || fabs(piv) < abstol) continue; /* leave to the bump */
        qinv[j] = rank;
        pinv[i] = rank;
        lint put0 = put;
        for (pos = b_tp[i]; pos < end; pos++) {
            lint j2 = b_ti[pos];
            if (qinv[j2] < 0) {
                u_i[put] = j2;
                u_x[put] = b_tx[pos];
                put++;
                iset[j2] ^= i;
                if (++qinv[j2] == -2) queue[tail++] = j2;
            }
        }
        u_p[rank + 1] = put;
        col_pivot[j] = piv;
        blo_trace_push(lu, i, j, piv, 0, put - put0 + 1, 1);
        rank++;
    }
    /* empty L columns, singletons.rs:385-391 */
    lint pos = l_p[rk];
    for (; rk < rank; rk++) {
        l_i[pos++] = -1;
        l_p[rk + 1] = pos;
    }
    return rank;
}

/* singletons.rs:398-503 */
static lint singleton_rows(lint m, const lint *b_begin, const lint *b_end, const lint *b_i,
                           const double *b_x, const lint *b_tp, const lint *b_ti,
                           lint *u_p, lint *l_p, lint *l_i, double *l_x,
                           double *col_pivot, lint *pinv, lint *qinv,
                           lint *iset, lint *queue, lint rank, double abstol, blo_lu *lu) {
    lint rk = rank, tail = 0;
    for (lint i = 0; i < m; i++) {
        if (pinv[i] < 0) {
            lint nz = b_tp[i + 1] - b_tp[i];
            lint x = 0;
            for (lint pos = b_tp[i]; pos < b_tp[i + 1]; pos++) x ^= b_ti[pos];
            iset[i] = x;
            pinv[i] = -nz - 1;
            if (nz == 1) queue[tail++] = i;
        }
    }
    lint put = l_p[rank];
    for (lint front = 0; front < tail; front++) {
        lint i = queue[front];
        assert(pinv[i] == -2 || pinv[i] == -1);
        if (pinv[i] == -1) continue;
        lint j = iset[i];
        assert(j >= 0 && j < m);
        assert(qinv[j] < 0);
        lint end = b_end[j];
        lint pos = b_begin[j];
        while (b_i[pos] != i) { assert(pos < end - 1); pos++; }
        double piv = b_x[pos];
        if (piv == 0.0 || fabs(piv) < abstol) continue;
        qinv[j] = rank;
        pinv[i] = rank;
        lint put0 = put;
        for (pos = b_begin[j]; pos < end; pos++) {
            lint i2 = b_i[pos];
            if (pinv[i2] < 0) {
                l_i[put] = i2;
                l_x[put] = b_x[pos] / piv;
                put++;
                iset[i2] ^= j;
                if (++pinv[i2] == -2) queue[tail++] = i2;
            }
        }
        l_i[put++] = -1;
        l_p[rank + 1] = put;
        col_pivot[j] = piv;
        blo_trace_push(lu, i, j, piv, 1, 1, put - put0);
        rank++;
    }
    /* empty U rows, singletons.rs:495-500 */
    lint pos = u_p[rk];
    for (; rk < rank; rk++) u_p[rk + 1] = pos;
    return rank;
}

/* singletons.rs:81-264 */
int blo_singletons(blo_lu *lu, const lint *b_begin, const lint *b_end, const lint *b_i, const double *b_x) {
    const lint m = lu->m;
    lint *iwork1 = lu->iwork1, *iwork2 = lu->iwork1 + m;
    lint *b_tp = lu->w_begin, *b_ti = lu->w_index;
    double *b_tx = lu->w_value;
    double tic = blo_now();

    /* pointers and nnz, singletons.rs:119-132 */
    lint b_nz = 0;
    for (lint j = 0; j < m; j++) {
        if (b_end[j] < b_begin[j]) return BLO_ERROR_INVALID_ARGUMENT;
        b_nz += b_end[j] - b_begin[j];
    }
    /* memory, singletons.rs:135-150 */
    int ok = 1;
    if (lu->l_mem < b_nz) { lu->addmem_l = b_nz - lu->l_mem; ok = 0; }
    if (lu->u_mem < b_nz) { lu->addmem_u = b_nz - lu->u_mem; ok = 0; }
    if (lu->w_mem < b_nz) { lu->addmem_w = b_nz - lu->w_mem; ok = 0; }
    if (!ok) return BLO_REALLOCATE;

    /* row counts + index range, singletons.rs:154-173 */
    memset(iwork1, 0, (size_t)m * sizeof(lint));
    for (lint j = 0; j < m; j++)
        for (lint pos = b_begin[j]; pos < b_end[j]; pos++) {
            lint i = b_i[pos];
            if (i < 0 || i >= m) return BLO_ERROR_INVALID_ARGUMENT;
            iwork1[i]++;
        }
    /* row-wise copy, adjacent-duplicate check, singletons.rs:176-201 */
    lint put = 0;
    for (lint i = 0; i < m; i++) {
        b_tp[i] = put;
        put += iwork1[i];
        iwork1[i] = b_tp[i];
    }
    b_tp[m] = put;
    assert(put == b_nz);
    ok = 1;
    for (lint j = 0; j < m; j++)
        for (lint pos = b_begin[j]; pos < b_end[j]; pos++) {
            lint i = b_i[pos];
            put = iwork1[i]++;
            b_ti[put] = j;
            b_tx[put] = b_x[pos];
            if (put > b_tp[i] && b_ti[put - 1] == j) ok = 0;
        }
    if (!ok) return BLO_ERROR_INVALID_ARGUMENT;

    for (lint i = 0; i < m; i++) lu->pinv[i] = -1;
    for (lint j = 0; j < m; j++) lu->qinv[j] = -1;

    lu->l_begin_p[0] = 0;
    lu->u_begin[0] = 0;
    lint rank = 0;
    if (lu->nzbias >= 0) { /* singletons.rs:213-228 */
        rank = singleton_cols(m, b_begin, b_end, b_i, b_tp, b_ti, b_tx, lu->u_begin, lu->u_index,
                              lu->u_value, lu->l_begin_p, lu->l_index, lu->col_pivot, lu->pinv,
                              lu->qinv, iwork1, iwork2, rank, lu->abstol, lu);
        rank = singleton_rows(m, b_begin, b_end, b_i, b_x, b_tp, b_ti, lu->u_begin, lu->l_begin_p,
                              lu->l_index, lu->l_value, lu->col_pivot, lu->pinv, lu->qinv,
                              iwork1, iwork2, rank, lu->abstol, lu);
    } else {               /* singletons.rs:229-245 */
        rank = singleton_rows(m, b_begin, b_end, b_i, b_x, b_tp, b_ti, lu->u_begin, lu->l_begin_p,
                              lu->l_index, lu->l_value, lu->col_pivot, lu->pinv, lu->qinv,
                              iwork1, iwork2, rank, lu->abstol, lu);
        rank = singleton_cols(m, b_begin, b_end, b_i, b_tp, b_ti, b_tx, lu->u_begin, lu->u_index,
                              lu->u_value, lu->l_begin_p, lu->l_index, lu->col_pivot, lu->pinv,
                              lu->qinv, iwork1, iwork2, rank, lu->abstol, lu);
    }
    /* counters back to -1, singletons.rs:248-257 */
    for (lint i = 0; i < m; i++) if (lu->pinv[i] < 0) lu->pinv[i] = -1;
    for (lint j = 0; j < m; j++) if (lu->qinv[j] < 0) lu->qinv[j] = -1;

    lu->matrix_nz = b_nz;
    lu->rank = rank;
    lu->time_singletons = blo_now() - tic;
    return BLO_OK;
}

/* ------------------------------------------------------------------ */
/* Phase 2: setup_bump.rs:55-264                                       */
/* ------------------------------------------------------------------ */
int blo_setup_bump(blo_lu *lu, const lint *b_begin, const lint *b_end, const lint *b_i, const double *b_x) {
    const lint m = lu->m, rank = lu->rank;
    const lint b_nz = lu->matrix_nz;
    const lint l_nz = lu->l_begin_p[rank] - rank;
    const lint u_nz = lu->u_begin[rank];
    const double abstol = lu->abstol, stretch = lu->stretch;
    const lint pad = lu->pad;
    const lint *pinv = lu->pinv, *qinv = lu->qinv;
    lint *w_begin = lu->w_begin, *w_end = lu->w_end;
    lint *w_begin2 = w_begin + m, *w_end2 = w_end + m;
    lint *w_index = lu->w_index;
    double *w_value = lu->w_value, *colmax = lu->col_pivot;
    lint *iwork0 = lu->iwork0;
    lint bump_nz = b_nz - l_nz - u_nz - rank;
    lint min_rownz = 0, min_colnz = 0;

    assert(l_nz >= 0 && u_nz >= 0 && bump_nz >= 0);

    /* setup_bump.rs:107-112 */
    lint need = bump_nz + (lint)(stretch * (double)bump_nz) + (m - rank) * pad;
    need *= 2;
    if (need > lu->w_mem) { lu->addmem_w = need - lu->w_mem; return BLO_REALLOCATE; }

    blo_file_empty(2 * m, w_begin, w_end, lu->w_flink, lu->w_blink, lu->w_mem);

    /* column file + colmax + row counts, setup_bump.rs:124-185 */
    blo_list_init(lu->colcount_flink, lu->colcount_blink, m, m + 2, &min_colnz);
    lint put = 0;
    for (lint j = 0; j < m; j++) {
        if (qinv[j] >= 0) continue;
        lint cnz = 0;
        double cmx = 0.0;
        for (lint pos = b_begin[j]; pos < b_end[j]; pos++) {
            lint i = b_i[pos];
            if (pinv[i] >= 0) continue;
            cmx = fmax(cmx, fabs(b_x[pos]));
            cnz++;
        }
        if (cmx == 0.0 || cmx < abstol) {
            colmax[j] = 0.0; /* column stays empty, bucket 0 */
            blo_list_add(j, 0, lu->colcount_flink, lu->colcount_blink, m, &min_colnz);
            bump_nz -= cnz;
        } else {
            colmax[j] = cmx;
            blo_list_add(j, cnz, lu->colcount_flink, lu->colcount_blink, m, &min_colnz);
            w_begin[j] = put;
            for (lint pos = b_begin[j]; pos < b_end[j]; pos++) {
                lint i = b_i[pos];
                if (pinv[i] >= 0) continue;
                w_index[put] = i;
                w_value[put] = b_x[pos];
                put++;
                iwork0[i]++;
            }
            w_end[j] = put;
            put += (lint)(stretch * (double)cnz) + pad;
            blo_list_move(j, 0, lu->w_flink, lu->w_blink, 2 * m, NULL);
        }
    }

    /* row file (pattern only), setup_bump.rs:188-224 */
    blo_list_init(lu->rowcount_flink, lu->rowcount_blink, m, m + 2, &min_rownz);
    for (lint i = 0; i < m; i++) {
        if (pinv[i] >= 0) continue;
        lint rnz = iwork0[i];
        iwork0[i] = 0;
        blo_list_add(i, rnz, lu->rowcount_flink, lu->rowcount_blink, m, &min_rownz);
        w_begin2[i] = w_end2[i] = put;
        put += rnz;
        blo_list_move(m + i, 0, lu->w_flink, lu->w_blink, 2 * m, NULL);
        put += (lint)(stretch * (double)rnz) + pad;
    }
    for (lint j = 0; j < m; j++)
        for (lint pos = w_begin[j]; pos < w_end[j]; pos++) {
            lint i = w_index[pos];
            w_index[w_end2[i]++] = j;
        }
    w_begin[2 * m] = put;
    assert(w_begin[2 * m] <= w_end[2 * m]);

    /* D11: the two release-mode consistency asserts, setup_bump.rs:228-251 */
    if (lu->check_file_diff) {
        assert(blo_file_diff(m, w_begin, w_end, w_begin2, w_end2, w_index, NULL) == 0);
        assert(blo_file_diff(m, w_begin2, w_end2, w_begin, w_end, w_index, NULL) == 0);
    }

    lu->bump_nz = bump_nz;
    lu->bump_size = m - rank;
    lu->min_colnz = min_colnz;
    lu->min_rownz = min_rownz;
    return BLO_OK;
}

/* ------------------------------------------------------------------ */
/* Markowitz search: markowitz.rs:34-219                               */
/* ------------------------------------------------------------------ */
static int mk_done(blo_lu *lu, lint pivot_row, lint pivot_col, lint nsearch,
                   lint min_colnz, lint min_rownz, double tic) {
    lu->pivot_row = pivot_row;
    lu->pivot_col = pivot_col;
    lu->nsearch_pivot += nsearch;
    if (min_colnz >= 0) lu->min_colnz = min_colnz;
    if (min_rownz >= 0) lu->min_rownz = min_rownz;
    lu->time_search_pivot += blo_now() - tic;
    return BLO_OK;
}

int blo_markowitz(blo_lu *lu) {
    const lint m = lu->m;
    const lint *w_begin = lu->w_begin, *w_end = lu->w_end, *w_index = lu->w_index;
    const double *w_value = lu->w_value, *colmax = lu->col_pivot;
    const lint *colcount_flink = lu->colcount_flink;
    lint *rowcount_flink = lu->rowcount_flink, *rowcount_blink = lu->rowcount_blink;
    const double abstol = lu->abstol, reltol = lu->reltol;
    const lint maxsearch = lu->maxsearch, search_rows = lu->search_rows;
    const lint nz_start = search_rows ? (lu->min_colnz < lu->min_rownz ? lu->min_colnz : lu->min_rownz)
                                      : lu->min_colnz;
    const int64_t m64 = m;
    double tic = blo_now();
    lint pivot_row = -1, pivot_col = -1;
    int64_t mc64 = m64 * m64;
    lint nsearch = 0, min_colnz = -1, min_rownz = -1;
    assert(nz_start >= 1);

    /* empty column => rank-deficiency step, markowitz.rs:73-78 */
    if (colcount_flink[m] != m) {
        pivot_col = colcount_flink[m];
        assert(pivot_col >= 0 && pivot_col < m);
        assert(w_end[pivot_col] == w_begin[pivot_col]);
        return mk_done(lu, pivot_row, pivot_col, nsearch, min_colnz, min_rownz, tic);
    }

    for (lint nz = nz_start; nz <= m; nz++) {
        /* columns with nz entries, markowitz.rs:82-123 */
        lint j;
        for (j = colcount_flink[m + nz]; j < m; j = colcount_flink[j]) {
            if (min_colnz == -1) min_colnz = nz;
            assert(w_end[j] - w_begin[j] == nz);
            double cmx = colmax[j];
            assert(cmx >= 0.0);
#if !BLO_REPAIR_D6
            if (cmx == 0.0 || cmx < abstol) BLO_DEFECT_TRAP("D6", "markowitz.rs:90-92 continues without advancing j (endless loop)");
#endif
            if (cmx == 0.0 || cmx < abstol) continue; /* D6 repaired: advance j.  Reached when a column keeps entries but its pivot-row entry was dropped from U, so pivot.rs:96-106 never emptied it */
            double tol = fmax(abstol, reltol * cmx);
            for (lint pos = w_begin[j]; pos < w_end[j]; pos++) {
                double x = fabs(w_value[pos]);
                if (x == 0.0 || x < tol) continue;
                lint i = w_index[pos];
                assert(i >= 0 && i < m);
                int64_t nz1 = nz, nz2 = w_end[m + i] - w_begin[m + i];
                assert(nz2 >= 1);
                int64_t mc = (nz1 - 1) * (nz2 - 1);
                if (mc < mc64) {
                    mc64 = mc;
                    pivot_row = i;
                    pivot_col = j;
                    if (search_rows && mc64 <= (nz1 - 1) * (nz1 - 1))
                        return mk_done(lu, pivot_row, pivot_col, nsearch, min_colnz, min_rownz, tic);
                }
            }
            assert(mc64 < m64 * m64);
            if (++nsearch >= maxsearch)
                return mk_done(lu, pivot_row, pivot_col, nsearch, min_colnz, min_rownz, tic);
        }
        assert(j == m + nz);

        if (!search_rows) continue;

        /* rows with nz entries, markowitz.rs:130-190 */
        lint i, inext;
        for (i = rowcount_flink[m + nz]; i < m; i = inext) {
            if (min_rownz == -1) min_rownz = nz;
            inext = rowcount_flink[i];
            assert(w_end[m + i] - w_begin[m + i] == nz);
            int cheap = 0, found = 0;
            for (lint pos = w_begin[m + i]; pos < w_end[m + i]; pos++) {
                lint jj = w_index[pos];
                assert(jj >= 0 && jj < m);
                int64_t nz1 = nz, nz2 = w_end[jj] - w_begin[jj];
                assert(nz2 >= 1);
                int64_t mc = (nz1 - 1) * (nz2 - 1);
                if (mc >= mc64) continue;
                cheap = 1;
                double cmx = colmax[jj];
                assert(cmx >= 0.0);
                if (cmx == 0.0 || cmx < abstol) continue;
                lint where = w_begin[jj];
                while (w_index[where] != i) { assert(where < w_end[jj] - 1); where++; }
                double x = fabs(w_value[where]);
                if (x >= abstol && x >= reltol * cmx) {
                    found = 1;
                    mc64 = mc;
                    pivot_row = i;
                    pivot_col = jj;
                    if (mc64 <= nz1 * (nz1 - 1))
                        return mk_done(lu, pivot_row, pivot_col, nsearch, min_colnz, min_rownz, tic);
                }
            }
            if (cheap && !found) {
                /* park the row until it is updated, markowitz.rs:178-179 */
                blo_list_move(i, m + 1, rowcount_flink, rowcount_blink, m, NULL);
            } else {
                assert(mc64 < m64 * m64);
                if (++nsearch >= maxsearch)
                    return mk_done(lu, pivot_row, pivot_col, nsearch, min_colnz, min_rownz, tic);
            }
        }
        assert(i == m + nz);
    }
    return mk_done(lu, pivot_row, pivot_col, nsearch, min_colnz, min_rownz, tic);
}

/* ------------------------------------------------------------------ */
/* Phase 3 driver: factorize_bump.rs:12-49                             */
/* ------------------------------------------------------------------ */
int blo_factorize_bump(blo_lu *lu) {
    const lint m = lu->m;
    while (lu->rank + lu->rankdef < m) {
        if (lu->pivot_col < 0) {
            int st = blo_markowitz(lu);
            if (st != BLO_OK) return st;
        }
        assert(lu->pivot_col >= 0);
        if (lu->pivot_row < 0) {
            /* empty column: drop it, no pivot */
            blo_trace_push(lu, -1, lu->pivot_col, 0.0, 7, 0, 0);
            blo_list_remove(lu->colcount_flink, lu->colcount_blink, lu->pivot_col);
            lu->pivot_col = -1;
            lu->rankdef++;
        } else {
            assert(lu->pinv[lu->pivot_row] == -1);
            assert(lu->qinv[lu->pivot_col] == -1);
            int st = blo_pivot(lu);
            if (st != BLO_OK) return st;
            lu->pinv[lu->pivot_row] = lu->rank;
            lu->qinv[lu->pivot_col] = lu->rank;
            lu->pivot_col = lu->pivot_row = -1;
            lu->rank++;
        }
    }
    return BLO_OK;
}

/* ------------------------------------------------------------------ */
/* Phase 4: build_factors.rs:113-423                                   */
/* ------------------------------------------------------------------ */
int blo_build_factors(blo_lu *lu) {
    const lint m = lu->m, rank = lu->rank, pad = lu->pad;
    const double stretch = lu->stretch;
    lint *pivotcol = lu->pivotcol, *pivotrow = lu->pivotrow;
    lint *l_index = lu->l_index, *u_index = lu->u_index, *w_index = lu->w_index;
    double *l_value = lu->l_value, *u_value = lu->u_value, *w_value = lu->w_value;
    lint *iwork1 = lu->iwork1;

    lint l_nz = lu->l_begin_p[rank] - rank;
    lint u_nz = lu->u_begin[rank];

    /* memory, build_factors.rs:163-177 */
    lint need = 2 * (l_nz + m);
    if (lu->l_mem < need) { lu->addmem_l = need - lu->l_mem; return BLO_REALLOCATE; }
    need = u_nz + m + 1;
    if (lu->u_mem < need) { lu->addmem_u = need - lu->u_mem; return BLO_REALLOCATE; }
    need = u_nz + (lint)(stretch * (double)u_nz) + m * pad;
    if (lu->w_mem < need) { lu->addmem_w = need - lu->w_mem; return BLO_REALLOCATE; }

    /* permutations, build_factors.rs:192-209 */
    lint lrank = rank;
    for (lint i = 0; i < m; i++) {
        if (lu->pinv[i] < 0) lu->pinv[i] = lrank++;
        pivotrow[lu->pinv[i]] = i;
    }
    assert(lrank == m);
    lrank = rank;
    for (lint j = 0; j < m; j++) {
        if (lu->qinv[j] < 0) lu->qinv[j] = lrank++;
        pivotcol[lu->qinv[j]] = j;
    }
    assert(lrank == m);

    /* unit pivots for dependent columns, build_factors.rs:221-223 */
    for (lint k = rank; k < m; k++) lu->col_pivot[pivotcol[k]] = 1.0;

    /* L column-wise completion, build_factors.rs:229-238 */
    lint put = lu->l_begin_p[rank];
    for (lint k = rank; k < m; k++) {
        l_index[put++] = -1;
        lu->l_begin_p[k + 1] = put;
    }
    assert(lu->l_begin_p[m] == l_nz + m);
    for (lint i = 0; i < m; i++) lu->l_begin[i] = lu->l_begin_p[lu->pinv[i]];

    /* L row-wise, build_factors.rs:242-274 */
    memset(iwork1, 0, (size_t)m * sizeof(lint));
    for (lint get = 0; get < l_nz + m; get++)
        if (l_index[get] >= 0) iwork1[l_index[get]]++;
    put = l_nz + m;
    for (lint k = 0; k < m; k++) {
        lint i = pivotrow[k];
        lu->lt_begin_p[k] = put;
        lu->lt_begin[i] = put;
        put += iwork1[i];
        l_index[put++] = -1;
        iwork1[i] = lu->lt_begin_p[k];
    }
    assert(put == 2 * (l_nz + m));
    for (lint k = 0; k < m; k++) {
        lint ipivot = pivotrow[k];
        for (lint get = lu->l_begin_p[k]; l_index[get] >= 0; get++) {
            lint dst = iwork1[l_index[get]]++;
            l_index[dst] = ipivot;
            l_value[dst] = l_value[get];
        }
    }
    lu->r_begin[0] = 2 * (l_nz + m);

    /* U row-wise into the W file, build_factors.rs:286-351 */
    blo_file_empty(m, lu->w_begin, lu->w_end, lu->w_flink, lu->w_blink, lu->w_mem);
    memset(iwork1, 0, (size_t)m * sizeof(lint));
    put = 0;
    if (rank == m) {
        for (lint k = 0; k < m; k++) {
            lint jpivot = pivotcol[k];
            lu->w_begin[jpivot] = put;
            lint nz = 0;
            for (lint pos = lu->u_begin[k]; pos < lu->u_begin[k + 1]; pos++) {
                lint j = u_index[pos];
                w_index[put] = j;
                w_value[put++] = u_value[pos];
                iwork1[j]++;
                nz++;
            }
            lu->w_end[jpivot] = put;
            put += (lint)(stretch * (double)nz) + pad;
            blo_list_move(jpivot, 0, lu->w_flink, lu->w_blink, m, NULL);
        }
    } else {
        u_nz = 0;
        for (lint k = 0; k < rank; k++) {
            lint jpivot = pivotcol[k];
            lu->w_begin[jpivot] = put;
            lint nz = 0;
            for (lint pos = lu->u_begin[k]; pos < lu->u_begin[k + 1]; pos++) {
                lint j = u_index[pos];
                if (lu->qinv[j] < rank) {
                    w_index[put] = j;
                    w_value[put++] = u_value[pos];
                    iwork1[j]++;
                    nz++;
                }
            }
            lu->w_end[jpivot] = put;
            put += (lint)(stretch * (double)nz) + pad;
            blo_list_move(jpivot, 0, lu->w_flink, lu->w_blink, m, NULL);
            u_nz += nz;
        }
        for (lint k = rank; k < m; k++) {
            lint jpivot = pivotcol[k];
            lu->w_begin[jpivot] = lu->w_end[jpivot] = put;
            put += pad;
            blo_list_move(jpivot, 0, lu->w_flink, lu->w_blink, m, NULL);
        }
    }
    assert(put <= lu->w_end[m]);
    lu->w_begin[m] = put;

    /* U column-wise, build_factors.rs:354-384 */
    u_index[0] = -1;
    put = 1;
    for (lint k = 0; k < m; k++) {
        lint j = pivotcol[k], i = pivotrow[k];
        lint nz = iwork1[j];
        if (nz == 0) {
            lu->u_begin[i] = 0;
        } else {
            lu->u_begin[i] = put;
            put += nz;
            u_index[put++] = -1;
        }
        iwork1[j] = lu->u_begin[i];
    }
    lu->u_begin[m] = put;
    for (lint k = 0; k < m; k++) {
        lint jpivot = pivotcol[k], i = pivotrow[k];
        for (lint pos = lu->w_begin[jpivot]; pos < lu->w_end[jpivot]; pos++) {
            lint j = w_index[pos];
            lint dst = iwork1[j]++;
            assert(dst >= 1);
            u_index[dst] = i;
            u_value[dst] = w_value[pos];
        }
    }

    /* pmap/qmap overwrite pinv/qinv, build_factors.rs:395-400 */
    for (lint k = 0; k < m; k++) {
        lint i = pivotrow[k], j = pivotcol[k];
        lu->pinv[j] = i; /* pmap */
        lu->qinv[i] = j; /* qmap */
    }

    /* row_pivot, min/max pivot, build_factors.rs:403-410 */
    double max_pivot = 0.0, min_pivot = INFINITY;
    for (lint i = 0; i < m; i++) {
        lu->row_pivot[i] = lu->col_pivot[lu->qinv[i]];
        double pv = fabs(lu->row_pivot[i]);
        max_pivot = fmax(pv, max_pivot);
        min_pivot = fmin(pv, min_pivot);
    }
    memcpy(lu->p, pivotrow, (size_t)m * sizeof(lint));

    lu->min_pivot = min_pivot;
    lu->max_pivot = max_pivot;
    lu->pivotlen = m;
    lu->l_nz = l_nz;
    lu->u_nz = u_nz;
    lu->r_nz = 0;
    return BLO_OK;
}

/* ------------------------------------------------------------------ */
/* condest.rs:15-157                                                   */
/* ------------------------------------------------------------------ */
static double normest(lint m, const lint *u_begin, const lint *u_i, const double *u_x,
                      const double *pivot, const lint *perm, int upper, double *work) {
    double x1norm = 0.0, xinfnorm = 0.0, y1norm = 0.0;
    lint kbeg, kend, kinc;
    if (upper) { kbeg = 0; kend = m; kinc = 1; } else { kbeg = m - 1; kend = -1; kinc = -1; }
    for (lint k = kbeg; k != kend; k += kinc) {
        lint j = perm ? perm[k] : k;
        double temp = 0.0;
        for (lint p = u_begin[j]; u_i[p] >= 0; p++) temp -= work[u_i[p]] * u_x[p];
        temp += temp >= 0.0 ? 1.0 : -1.0;
        if (pivot) temp /= pivot[j];
        work[j] = temp;
        x1norm += fabs(temp);
        xinfnorm = fmax(xinfnorm, fabs(temp));
    }
    if (upper) { kbeg = m - 1; kend = -1; kinc = -1; } else { kbeg = 0; kend = m; kinc = 1; }
    for (lint k = kbeg; k != kend; k += kinc) {
        lint j = perm ? perm[k] : k;
        if (pivot) work[j] /= pivot[j];
        double temp = work[j];
        for (lint p = u_begin[j]; u_i[p] >= 0; p++) work[u_i[p]] -= temp * u_x[p];
        y1norm += fabs(temp);
    }
    return fmax(y1norm / x1norm, xinfnorm);
}

double blo_condest(lint m, const lint *u_begin, const lint *u_i, const double *u_x,
                   const double *pivot, const lint *perm, int upper, double *work,
                   double *norm, double *norminv) {
    double u_norm = 0.0;
    for (lint j = 0; j < m; j++) {
        double colsum = pivot ? fabs(pivot[j]) : 1.0;
        for (lint p = u_begin[j]; u_i[p] >= 0; p++) colsum += fabs(u_x[p]);
        u_norm = fmax(u_norm, colsum);
    }
    double u_invnorm = normest(m, u_begin, u_i, u_x, pivot, perm, upper, work);
    if (norm) *norm = u_norm;
    if (norminv) *norminv = u_invnorm;
    return u_norm * u_invnorm;
}

/* matrix_norm.rs:8-48 */
void blo_matrix_norm(blo_lu *lu, const lint *b_begin, const lint *b_end, const lint *b_i, const double *b_x) {
    const lint m = lu->m, rank = lu->rank;
    double *rowsum = lu->work1;
    assert(lu->nupdate == 0);
    for (lint i = 0; i < m; i++) rowsum[i] = 0.0;
    double onenorm = 0.0, infnorm = 0.0;
    for (lint k = 0; k < rank; k++) {
        lint jpivot = lu->pivotcol[k];
        double colsum = 0.0;
        for (lint pos = b_begin[jpivot]; pos < b_end[jpivot]; pos++) {
            colsum += fabs(b_x[pos]);
            rowsum[b_i[pos]] += fabs(b_x[pos]);
        }
        onenorm = fmax(onenorm, colsum);
    }
    for (lint k = rank; k < m; k++) {
        rowsum[lu->pivotrow[k]] += 1.0;
        onenorm = fmax(onenorm, 1.0);
    }
    for (lint i = 0; i < m; i++) infnorm = fmax(infnorm, rowsum[i]);
    lu->onenorm = onenorm;
    lu->infnorm = infnorm;
}

static double vec_onenorm(lint m, const double *x) {
    double d = 0.0;
    for (lint i = 0; i < m; i++) d += fabs(x[i]);
    return d;
}

/* residual_test.rs:16-152 */
void blo_residual_test(blo_lu *lu, const lint *b_begin, const lint *b_end, const lint *b_i, const double *b_x) {
    const lint m = lu->m, rank = lu->rank;
    const lint *p = lu->p, *pivotcol = lu->pivotcol, *pivotrow = lu->pivotrow;
    const lint *l_index = lu->l_index, *u_index = lu->u_index;
    const double *l_value = lu->l_value, *u_value = lu->u_value, *row_pivot = lu->row_pivot;
    double *rhs = lu->work0, *lhs = lu->work1;
    assert(lu->nupdate == 0);

    /* forward system */
    for (lint k = 0; k < m; k++) {
        double d = 0.0;
        for (lint pos = lu->lt_begin_p[k]; l_index[pos] >= 0; pos++) d += lhs[l_index[pos]] * l_value[pos];
        lint ipivot = p[k];
        rhs[ipivot] = d <= 0.0 ? 1.0 : -1.0;
        lhs[ipivot] = rhs[ipivot] - d;
    }
    for (lint k = m - 1; k >= 0; k--) {
        lint ipivot = pivotrow[k];
        lhs[ipivot] /= row_pivot[ipivot];
        double d = lhs[ipivot];
        for (lint pos = lu->u_begin[ipivot]; u_index[pos] >= 0; pos++) lhs[u_index[pos]] -= d * u_value[pos];
    }
    for (lint k = 0; k < rank; k++) {
        lint ipivot = pivotrow[k], jpivot = pivotcol[k];
        double d = lhs[ipivot];
        for (lint pos = b_begin[jpivot]; pos < b_end[jpivot]; pos++) rhs[b_i[pos]] -= d * b_x[pos];
    }
    for (lint k = rank; k < m; k++) {
        lint ipivot = pivotrow[k];
        rhs[ipivot] -= lhs[ipivot];
    }
    double norm_ftran = vec_onenorm(m, lhs);
    double norm_ftran_res = vec_onenorm(m, rhs);

    /* transposed system */
    for (lint k = 0; k < m; k++) {
        lint ipivot = pivotrow[k];
        double d = 0.0;
        for (lint pos = lu->u_begin[ipivot]; u_index[pos] >= 0; pos++) d += lhs[u_index[pos]] * u_value[pos];
        rhs[ipivot] = d <= 0.0 ? 1.0 : -1.0;
        lhs[ipivot] = (rhs[ipivot] - d) / row_pivot[ipivot];
    }
    for (lint k = m - 1; k >= 0; k--) {
        double d = 0.0;
        for (lint pos = lu->l_begin_p[k]; l_index[pos] >= 0; pos++) d += lhs[l_index[pos]] * l_value[pos];
        lhs[p[k]] -= d;
    }
    for (lint k = 0; k < rank; k++) {
        lint ipivot = pivotrow[k], jpivot = pivotcol[k];
        double d = 0.0;
        for (lint pos = b_begin[jpivot]; pos < b_end[jpivot]; pos++) d += lhs[b_i[pos]] * b_x[pos];
        rhs[ipivot] -= d;
    }
    for (lint k = rank; k < m; k++) {
        lint ipivot = pivotrow[k];
        rhs[ipivot] -= lhs[ipivot];
    }
    double norm_btran = vec_onenorm(m, lhs);
    double norm_btran_res = vec_onenorm(m, rhs);

    blo_matrix_norm(lu, b_begin, b_end, b_i, b_x);
    assert(lu->onenorm > 0.0);
    assert(lu->infnorm > 0.0);
    lu->residual_test = fmax(norm_ftran_res / ((double)m + lu->onenorm * norm_ftran),
                             norm_btran_res / ((double)m + lu->infnorm * norm_btran));
    for (lint i = 0; i < m; i++) lu->work0[i] = 0.0;
}

/* ------------------------------------------------------------------ */
/* factorize.rs:34-182                                                 */
/* ------------------------------------------------------------------ */
int blo_lu_factorize(blo_lu *lu, const lint *b_begin, const lint *b_end,
                     const lint *b_i, const double *b_x, int c0ntinue) {
    double tic = blo_now();
    int st = BLO_OK;

    if (!c0ntinue) {
        blo_lu_reset(lu);
        lu->task = BLO_TASK_SINGLETONS;
    }
#if BLO_REPAIR_D7
    /* D7 repair: what BASICLU's lu_load does on every entry */
    lu->addmem_l = lu->addmem_u = lu->addmem_w = 0;
    if (lu->task != BLO_TASK_NONE) lu->w_end[2 * lu->m] = lu->w_mem;
#endif

    switch (lu->task) {
    case BLO_TASK_SINGLETONS:
        st = blo_singletons(lu, b_begin, b_end, b_i, b_x);
        if (st != BLO_OK) goto out;
        lu->task = BLO_TASK_SETUP_BUMP;
        /* fallthrough */
    case BLO_TASK_SETUP_BUMP:
        st = blo_setup_bump(lu, b_begin, b_end, b_i, b_x);
        if (st != BLO_OK) goto out;
        lu->task = BLO_TASK_FACTORIZE_BUMP;
        /* fallthrough */
    case BLO_TASK_FACTORIZE_BUMP:
        st = blo_factorize_bump(lu);
        if (st != BLO_OK) goto out;
        /* fallthrough */
    case BLO_TASK_BUILD_FACTORS:
        break;
    default:
        return BLO_ERROR_INVALID_CALL; /* factorize.rs:102-105 */
    }

    lu->task = BLO_TASK_BUILD_FACTORS;
    st = blo_build_factors(lu);
    if (st != BLO_OK) goto out;

    lu->task = BLO_TASK_NONE;
    lu->nupdate = 0;
    lu->ftran_for_update = lu->btran_for_update = -1;
    lu->nfactorize++;

    /* factorize.rs:121-147 */
    lu->condest_l = blo_condest(lu->m, lu->l_begin, lu->l_index, lu->l_value, NULL, lu->p, 0,
                                lu->work1, &lu->norm_l, &lu->normest_l_inv);
    lu->condest_u = blo_condest(lu->m, lu->u_begin, lu->u_index, lu->u_value, lu->row_pivot, lu->p, 1,
                                lu->work1, &lu->norm_u, &lu->normest_u_inv);
    blo_residual_test(lu, b_begin, b_end, b_i, b_x);

    /* factorize.rs:160-166 */
    {
        double factor_cost = 0.04 * (double)lu->m + 0.07 * (double)lu->matrix_nz +
                             0.20 * (double)lu->bump_nz + 0.20 * (double)lu->nsearch_pivot +
                             0.008 * (double)lu->factor_flops;
        lu->update_cost_denom = factor_cost * 250.0;
    }
    st = lu->rank < lu->m ? BLO_WARNING_SINGULAR_MATRIX : BLO_OK;

out: {
        double el = blo_now() - tic;
        lu->time_factorize += el;
        lu->time_factorize_total += el;
    }
    return st;
}
