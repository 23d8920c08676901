/* blo_pivot.c -- CPU oracle (test infrastructure): one elimination step.
 * Follows /root/reference/src/lu/pivot.rs.  Floating point: a*b and the
 * subtraction are separate roundings (compile with -ffp-contract=off), as in
 * the Rust reference. */
#include "blo_int.h"

#define MAXROW_SMALL 64 /* pivot.rs:22 */

/* algorithmic-bytes accounting, SURVEY.md section 8(d): 12 B per stored
 * nonzero, 4 B per pattern-only index */
#define ACC_COL(lu, oldnz, newnz) ((lu)->elim_bytes += 12.0 * (double)((oldnz) + (newnz)))
#define ACC_ROW(lu, oldnz, newnz) ((lu)->elim_bytes += 4.0 * (double)((oldnz) + (newnz)))

static void remove_col(blo_lu *lu, lint j);

/* shared prologue of pivot_any / pivot_small: pivot.rs:142-208 and 490-558.
 * Moves the pivot to the front of its column and row, bounds the file growth,
 * garbage-collects if needed.  Returns BLO_REALLOCATE if W is too small. */
static int prologue(blo_lu *lu, lint *pcbeg, lint *pcend, lint *prbeg, lint *prend, double *ppivot) {
    const lint m = lu->m, pad = lu->pad;
    const double stretch = lu->stretch;
    const lint pivot_col = lu->pivot_col, pivot_row = lu->pivot_row;
    lint *w_begin = lu->w_begin, *w_end = lu->w_end, *w_index = lu->w_index;
    double *w_value = lu->w_value;
    lint cbeg = w_begin[pivot_col], cend = w_end[pivot_col];
    lint rbeg = w_begin[m + pivot_row], rend = w_end[m + pivot_row];
    const lint cnz1 = cend - cbeg - 1, rnz1 = rend - rbeg - 1;

    lint grow = 0, where = -1;
    for (lint pos = cbeg; pos < cend; pos++) {
        lint i = w_index[pos];
        if (i == pivot_row) where = pos;
        else {
            lint nz = w_end[m + i] - w_begin[m + i];
            grow += nz + rnz1 + (lint)(stretch * (double)(nz + rnz1)) + pad;
        }
    }
    assert(where >= 0);
    blo_iswap(w_index, cbeg, where);
    blo_fswap(w_value, cbeg, where);
    double pivot = w_value[cbeg];
    assert(pivot != 0.0);
    where = -1;
    for (lint rpos = rbeg; rpos < rend; rpos++) {
        lint j = w_index[rpos];
        if (j == pivot_col) where = rpos;
        else {
            lint nz = w_end[j] - w_begin[j];
            grow += nz + cnz1 + (lint)(stretch * (double)(nz + cnz1)) + pad;
        }
    }
    assert(where >= 0);
    blo_iswap(w_index, rbeg, where);
    lint room = w_end[2 * m] - w_begin[2 * m];
    if (grow > room) {
        blo_file_compress(2 * m, w_begin, w_end, lu->w_flink, w_index, w_value, stretch, pad);
        cbeg = w_begin[pivot_col]; cend = w_end[pivot_col];
        rbeg = w_begin[m + pivot_row]; rend = w_end[m + pivot_row];
        room = w_end[2 * m] - w_begin[2 * m];
        lu->ngarbage++;
    }
    if (grow > room) { lu->addmem_w = grow - room; return BLO_REALLOCATE; }
    *pcbeg = cbeg; *pcend = cend; *prbeg = rbeg; *prend = rend; *ppivot = pivot;
    return BLO_OK;
}

/* shared: compact column j around the marked rows, swap the pivot-row entry to
 * the front, make room for cnz1 appended entries.  pivot.rs:231-284 / 584-637.
 * Returns xrj; *pput is where appended entries go; *pcmx the max of the kept part. */
static double col_prepare(blo_lu *lu, lint j, lint cnz1, lint *pput, double *pcmx) {
    const lint m = lu->m, pad = lu->pad;
    const double stretch = lu->stretch;
    const lint pivot_row = lu->pivot_row;
    lint *w_begin = lu->w_begin, *w_end = lu->w_end, *w_index = lu->w_index;
    double *w_value = lu->w_value, *work = lu->work0;
    const lint *marked = lu->iwork0;
    double cmx = 0.0;
    lint where = -1;
    lint put = w_begin[j];
    const lint pos1 = w_begin[j];
    for (lint pos = pos1; pos < w_end[j]; pos++) {
        lint i = w_index[pos];
        lint position = marked[i];
        if (position > 0) {
            assert(i != pivot_row);
            work[position] = w_value[pos];
        } else {
            assert(position == 0);
            double x = fabs(w_value[pos]);
            if (i == pivot_row) where = put;
            else if (x > cmx) cmx = x;
            w_index[put] = w_index[pos];
            w_value[put] = w_value[pos];
            put++;
        }
    }
    assert(where >= 0);
    w_end[j] = put;
    blo_iswap(w_index, pos1, where);
    blo_fswap(w_value, pos1, where);
    double xrj = w_value[pos1];

    lint room = w_begin[lu->w_flink[j]] - put;
    if (room < cnz1) {
        lint nz = w_end[j] - w_begin[j];
        room = cnz1 + (lint)(stretch * (double)(nz + cnz1)) + pad;
        blo_file_reappend(j, 2 * m, w_begin, w_end, lu->w_flink, lu->w_blink, w_index, w_value, room);
        put = w_end[j];
        assert(w_begin[lu->w_flink[j]] - put == room);
        lu->nexpand++;
    }
    *pput = put;
    *pcmx = cmx;
    return xrj;
}

/* shared: compact row i (drop every column that is in the pivot row), make room
 * for rnz1 appended indices.  pivot.rs:346-380 / 711-746.  Returns put. */
static lint row_prepare(blo_lu *lu, lint i, lint rnz1) {
    const lint m = lu->m, pad = lu->pad;
    const double stretch = lu->stretch;
    const lint pivot_col = lu->pivot_col;
    lint *w_begin = lu->w_begin, *w_end = lu->w_end, *w_index = lu->w_index;
    const lint *marked = lu->iwork0;
    int found = 0;
    lint put = w_begin[m + i];
    for (lint rpos = w_begin[m + i]; rpos < w_end[m + i]; rpos++) {
        lint j = w_index[rpos];
        if (j == pivot_col) found = 1;
        if (marked[j] == 0) w_index[put++] = j;
    }
    assert(found);
    w_end[m + i] = put;
    lint room = w_begin[lu->w_flink[m + i]] - put;
    if (room < rnz1) {
        lint nz = w_end[m + i] - w_begin[m + i];
        room = rnz1 + (lint)(stretch * (double)(nz + rnz1)) + pad;
        blo_file_reappend(m + i, 2 * m, w_begin, w_end, lu->w_flink, lu->w_blink,
                          w_index, lu->w_value, room);
        put = w_end[m + i];
        assert(w_begin[lu->w_flink[m + i]] - put == room);
        lu->nexpand++;
    }
    return put;
}

/* shared epilogue: L column, pointers, unlink pivot row/col.  pivot.rs:403-426 / 776-799 */
static void epilogue(blo_lu *lu, lint cbeg, lint cend, lint rbeg, double pivot, lint u_put) {
    const lint m = lu->m, rank = lu->rank;
    const double droptol = lu->droptol;
    lint put = lu->l_begin_p[rank];
    for (lint pos = cbeg + 1; pos < cend; pos++) {
        double x = lu->w_value[pos] / pivot;
        if (fabs(x) > droptol) {
            lu->l_index[put] = lu->w_index[pos];
            lu->l_value[put] = x;
            put++;
        }
    }
    lu->l_index[put++] = -1;
    lu->l_begin_p[rank + 1] = put;
    lu->u_begin[rank + 1] = u_put;
    lu->col_pivot[lu->pivot_col] = pivot;
    lu->w_end[lu->pivot_col] = cbeg;
    lu->w_end[m + lu->pivot_row] = rbeg;
    blo_list_remove(lu->colcount_flink, lu->colcount_blink, lu->pivot_col);
    blo_list_remove(lu->rowcount_flink, lu->rowcount_blink, lu->pivot_row);
}

/* pivot.rs:114-458 */
static int pivot_any(blo_lu *lu) {
    const lint m = lu->m, rank = lu->rank;
    const double droptol = lu->droptol;
    lint *w_begin = lu->w_begin, *w_end = lu->w_end, *w_index = lu->w_index;
    double *w_value = lu->w_value, *work = lu->work0, *colmax = lu->col_pivot;
    lint *marked = lu->iwork0;
    lint cbeg, cend, rbeg, rend;
    double pivot;

    int st = prologue(lu, &cbeg, &cend, &rbeg, &rend, &pivot);
    if (st != BLO_OK) return st;
    const lint cnz1 = cend - cbeg - 1, rnz1 = rend - rbeg - 1;

    lint u_put = lu->u_begin[rank];
    assert(u_put < lu->u_mem);

    /* column file update, pivot.rs:219-331 */
    lint position = 1;
    for (lint pos = cbeg + 1; pos < cend; pos++) marked[w_index[pos]] = position++;
    for (lint rpos = rbeg + 1; rpos < rend; rpos++) {
        lint j = w_index[rpos];
        assert(j != lu->pivot_col);
        lint oldnz = w_end[j] - w_begin[j];
        lint put;
        double cmx;
        double xrj = col_prepare(lu, j, cnz1, &put, &cmx);
        double a = xrj / pivot;
        for (lint pos = 1; pos <= cnz1; pos++) work[pos] -= a * w_value[cbeg + pos];
        for (lint pos = 1; pos <= cnz1; pos++) {
            w_index[put] = w_index[cbeg + pos];
            w_value[put] = work[pos];
            put++;
            double x = fabs(work[pos]);
            if (x > cmx) cmx = x;
            work[pos] = 0.0;
        }
        w_end[j] = put;
        if (fabs(xrj) > droptol) {
            assert(u_put < lu->u_mem);
            lu->u_index[u_put] = j;
            lu->u_value[u_put] = xrj;
            u_put++;
        }
        assert(w_index[w_begin[j]] == lu->pivot_row);
        w_begin[j]++;
        lint nz = w_end[j] - w_begin[j];
        blo_list_move(j, nz, lu->colcount_flink, lu->colcount_blink, m, &lu->min_colnz);
        colmax[j] = cmx;
        ACC_COL(lu, oldnz, nz);
    }
    for (lint pos = cbeg + 1; pos < cend; pos++) marked[w_index[pos]] = 0;

    /* row file update, pivot.rs:335-401 */
    for (lint rpos = rbeg; rpos < rend; rpos++) marked[w_index[rpos]] = 1;
    assert(marked[lu->pivot_col] == 1);
    for (lint pos = cbeg + 1; pos < cend; pos++) {
        lint i = w_index[pos];
        assert(i != lu->pivot_row);
        lint oldnz = w_end[m + i] - w_begin[m + i];
        lint put = row_prepare(lu, i, rnz1);
        for (lint rpos = rbeg + 1; rpos < rend; rpos++) w_index[put++] = w_index[rpos];
        w_end[m + i] = put;
        lint nz = w_end[m + i] - w_begin[m + i];
        blo_list_move(i, nz, lu->rowcount_flink, lu->rowcount_blink, m, &lu->min_rownz);
        ACC_ROW(lu, oldnz, nz);
    }
    for (lint rpos = rbeg; rpos < rend; rpos++) marked[w_index[rpos]] = 0;

    epilogue(lu, cbeg, cend, rbeg, pivot, u_put);
    return BLO_OK;
}

/* pivot.rs:460-833.  Like pivot_any but drops updated entries <= droptol and
 * records them per column in a 64-bit mask (D5 repaired: true 64-bit). */
static int pivot_small(blo_lu *lu) {
    const lint m = lu->m, rank = lu->rank;
    const double droptol = lu->droptol;
    lint *w_begin = lu->w_begin, *w_end = lu->w_end, *w_index = lu->w_index;
    double *w_value = lu->w_value, *work = lu->work0, *colmax = lu->col_pivot;
    lint *marked = lu->iwork0;
    uint64_t *cancelled = lu->cancelled;
    lint cbeg, cend, rbeg, rend;
    double pivot;

    {
        lint c = w_end[lu->pivot_col] - w_begin[lu->pivot_col] - 1;
        assert(c <= MAXROW_SMALL);
    }
    int st = prologue(lu, &cbeg, &cend, &rbeg, &rend, &pivot);
    if (st != BLO_OK) return st;
    const lint cnz1 = cend - cbeg - 1, rnz1 = rend - rbeg - 1;

    lint u_put = lu->u_begin[rank];
    assert(u_put < lu->u_mem);

    /* column file update, pivot.rs:569-693 */
    lint position = 1;
    for (lint pos = cbeg + 1; pos < cend; pos++) marked[w_index[pos]] = position++;
    lint col_number = 0;
    for (lint rpos = rbeg + 1; rpos < rend; rpos++, col_number++) {
        lint j = w_index[rpos];
        assert(j != lu->pivot_col);
        lint oldnz = w_end[j] - w_begin[j];
        lint put;
        double cmx;
        double xrj = col_prepare(lu, j, cnz1, &put, &cmx);
        double a = xrj / pivot;
        for (lint pos = 1; pos <= cnz1; pos++) work[pos] -= a * w_value[cbeg + pos];
        uint64_t mask = 0;
        for (lint pos = 1; pos <= cnz1; pos++) {
            double x = fabs(work[pos]);
            if (x > droptol) {
                w_index[put] = w_index[cbeg + pos];
                w_value[put] = work[pos];
                put++;
                if (x > cmx) cmx = x;
            } else {
#if BLO_REPAIR_D5
                mask |= (uint64_t)1 << (pos - 1); /* cancellation in row w_index[cbeg+pos] */
#else
                if (pos - 1 >= 31) BLO_DEFECT_TRAP("D5", "pivot.rs:659 shifts an i32 mask by >= 31 (debug: panic; release: wrong row pattern)");
                mask |= (uint64_t)1 << (pos - 1);
#endif
            }
            work[pos] = 0.0;
        }
        w_end[j] = put;
        cancelled[col_number] = mask;
        if (fabs(xrj) > droptol) {
            assert(u_put < lu->u_mem);
            lu->u_index[u_put] = j;
            lu->u_value[u_put] = xrj;
            u_put++;
        }
        assert(w_index[w_begin[j]] == lu->pivot_row);
        w_begin[j]++;
        lint nz = w_end[j] - w_begin[j];
        blo_list_move(j, nz, lu->colcount_flink, lu->colcount_blink, m, &lu->min_colnz);
        colmax[j] = cmx;
        ACC_COL(lu, oldnz, nz);
    }
    for (lint pos = cbeg + 1; pos < cend; pos++) marked[w_index[pos]] = 0;

    /* row file update, pivot.rs:697-774 */
    for (lint rpos = rbeg; rpos < rend; rpos++) marked[w_index[rpos]] = 1;
    assert(marked[lu->pivot_col] == 1);
    uint64_t mask = 1;
    for (lint pos = cbeg + 1; pos < cend; pos++, mask <<= 1) {
        lint i = w_index[pos];
        assert(i != lu->pivot_row);
        lint oldnz = w_end[m + i] - w_begin[m + i];
        lint put = row_prepare(lu, i, rnz1);
        col_number = 0;
        for (lint rpos = rbeg + 1; rpos < rend; rpos++, col_number++)
            if ((cancelled[col_number] & mask) == 0) w_index[put++] = w_index[rpos];
        w_end[m + i] = put;
        lint nz = w_end[m + i] - w_begin[m + i];
        blo_list_move(i, nz, lu->rowcount_flink, lu->rowcount_blink, m, &lu->min_rownz);
        ACC_ROW(lu, oldnz, nz);
    }
    for (lint rpos = rbeg; rpos < rend; rpos++) marked[w_index[rpos]] = 0;

    epilogue(lu, cbeg, cend, rbeg, pivot, u_put);
    return BLO_OK;
}

/* pivot.rs:835-926: pivot row has a single entry => only L and the row file change */
static int pivot_singleton_row(blo_lu *lu) {
    const lint m = lu->m, rank = lu->rank;
    const double droptol = lu->droptol;
    const lint pivot_col = lu->pivot_col, pivot_row = lu->pivot_row;
    lint *w_begin = lu->w_begin, *w_end = lu->w_end, *w_index = lu->w_index;
    double *w_value = lu->w_value;
    const lint cbeg = w_begin[pivot_col], cend = w_end[pivot_col];
    const lint rbeg = w_begin[m + pivot_row], rend = w_end[m + pivot_row];
    assert(rend - rbeg - 1 == 0);

    lint where = cbeg;
    while (w_index[where] != pivot_row) { assert(where < cend - 1); where++; }
    double pivot = w_value[where];
    assert(pivot != 0.0);

    lint put = lu->l_begin_p[rank];
    for (lint pos = cbeg; pos < cend; pos++) {
        double x = w_value[pos] / pivot;
        if (pos != where && fabs(x) > droptol) {
            lu->l_index[put] = w_index[pos];
            lu->l_value[put] = x;
            put++;
        }
    }
    lu->l_index[put++] = -1;
    lu->l_begin_p[rank + 1] = put;
    lu->u_begin[rank + 1] = lu->u_begin[rank];

    for (lint pos = cbeg; pos < cend; pos++) {
        lint i = w_index[pos];
        if (i == pivot_row) continue;
        lint oldnz = w_end[m + i] - w_begin[m + i];
        lint wh = w_begin[m + i];
        while (w_index[wh] != pivot_col) { assert(wh < w_end[m + i] - 1); wh++; }
        w_index[wh] = w_index[--w_end[m + i]];
        lint nz = w_end[m + i] - w_begin[m + i];
        blo_list_move(i, nz, lu->rowcount_flink, lu->rowcount_blink, m, &lu->min_rownz);
        ACC_ROW(lu, oldnz, nz);
    }

    lu->col_pivot[pivot_col] = pivot;
    w_end[pivot_col] = cbeg;
    w_end[m + pivot_row] = rbeg;
    blo_list_remove(lu->colcount_flink, lu->colcount_blink, pivot_col);
    blo_list_remove(lu->rowcount_flink, lu->rowcount_blink, pivot_row);
    return BLO_OK;
}

/* pivot.rs:928-1025: pivot column has a single entry => only U and the column file change */
static int pivot_singleton_col(blo_lu *lu) {
    const lint m = lu->m, rank = lu->rank;
    const double droptol = lu->droptol;
    const lint pivot_col = lu->pivot_col, pivot_row = lu->pivot_row;
    lint *w_begin = lu->w_begin, *w_end = lu->w_end, *w_index = lu->w_index;
    double *w_value = lu->w_value;
    const lint cbeg = w_begin[pivot_col], cend = w_end[pivot_col];
    const lint rbeg = w_begin[m + pivot_row], rend = w_end[m + pivot_row];
    assert(cend - cbeg - 1 == 0);

    lint put = lu->u_begin[rank];
    double pivot = w_value[cbeg];
    assert(pivot != 0.0);
    int found = 0;
    double xrj = 0.0;
    for (lint rpos = rbeg; rpos < rend; rpos++) {
        lint j = w_index[rpos];
        if (j == pivot_col) { found = 1; continue; }
        lint where = -1;
        double cmx = 0.0;
        for (lint pos = w_begin[j]; pos < w_end[j]; pos++) {
            double x = fabs(w_value[pos]);
            if (w_index[pos] == pivot_row) { where = pos; xrj = w_value[pos]; }
            else if (x > cmx) cmx = x;
        }
        assert(where >= 0);
        if (fabs(xrj) > droptol) {
            lu->u_index[put] = j;
            lu->u_value[put] = xrj;
            put++;
        }
        lint oldnz = w_end[j] - w_begin[j];
        w_end[j]--;
        w_index[where] = w_index[w_end[j]];
        w_value[where] = w_value[w_end[j]];
        lint nz = w_end[j] - w_begin[j];
        blo_list_move(j, nz, lu->colcount_flink, lu->colcount_blink, m, &lu->min_colnz);
        lu->col_pivot[j] = cmx;
        ACC_COL(lu, oldnz, nz);
    }
    assert(found);
    lu->u_begin[rank + 1] = put;

    put = lu->l_begin_p[rank];
    lu->l_index[put++] = -1;
    lu->l_begin_p[rank + 1] = put;

    lu->col_pivot[pivot_col] = pivot;
    w_end[pivot_col] = cbeg;
    w_end[m + pivot_row] = rbeg;
    blo_list_remove(lu->colcount_flink, lu->colcount_blink, pivot_col);
    blo_list_remove(lu->rowcount_flink, lu->rowcount_blink, pivot_row);
    return BLO_OK;
}

/* pivot.rs:1027-1331: pivot column has exactly one off-diagonal entry */
static int pivot_doubleton_col(blo_lu *lu) {
    const lint m = lu->m, rank = lu->rank, pad = lu->pad;
    const double droptol = lu->droptol, stretch = lu->stretch;
    const lint pivot_col = lu->pivot_col, pivot_row = lu->pivot_row;
    lint *w_begin = lu->w_begin, *w_end = lu->w_end, *w_index = lu->w_index;
    double *w_value = lu->w_value, *colmax = lu->col_pivot;
    lint *marked = lu->iwork0;
    lint cbeg = w_begin[pivot_col];
    const lint cend = w_end[pivot_col];
    lint rbeg = w_begin[m + pivot_row], rend = w_end[m + pivot_row];
    const lint cnz1 = cend - cbeg - 1, rnz1 = rend - rbeg - 1;
    assert(cnz1 == 1);

    /* pivot to the front of column and row, pivot.rs:1068-1082 */
    if (w_index[cbeg] != pivot_row) {
        blo_iswap(w_index, cbeg, cbeg + 1);
        blo_fswap(w_value, cbeg, cbeg + 1);
    }
    assert(w_index[cbeg] == pivot_row);
    const double pivot = w_value[cbeg];
    assert(pivot != 0.0);
    const lint other_row = w_index[cbeg + 1];
    const double other_value = w_value[cbeg + 1];
    lint where = rbeg;
    while (w_index[where] != pivot_col) { assert(where < rend - 1); where++; }
    blo_iswap(w_index, rbeg, where);

    /* room for the other row, pivot.rs:1087-1111 */
    lint nz = w_end[m + other_row] - w_begin[m + other_row];
    const lint other_oldnz = nz;
    lint grow = nz + rnz1 + (lint)(stretch * (double)(nz + rnz1)) + pad;
    lint room = w_end[2 * m] - w_begin[2 * m];
    if (grow > room) {
        blo_file_compress(2 * m, w_begin, w_end, lu->w_flink, w_index, w_value, stretch, pad);
        cbeg = w_begin[pivot_col];
        rbeg = w_begin[m + pivot_row];
        rend = w_end[m + pivot_row];
        room = w_end[2 * m] - w_begin[2 * m];
        lu->ngarbage++;
    }
    if (grow > room) { lu->addmem_w = grow - room; return BLO_REALLOCATE; }

    /* column file update, pivot.rs:1115-1222 */
    lint u_put = lu->u_begin[rank];
    lint put = rbeg + 1;
    lint ncancelled = 0;
    for (lint rpos = rbeg + 1; rpos < rend; rpos++) {
        lint j = w_index[rpos];
        assert(j != pivot_col);
        double cmx = 0.0;
        lint where_pivot = -1, where_other = -1;
        lint end = w_end[j];
        const lint oldnz = end - w_begin[j];
        for (lint pos = w_begin[j]; pos < end; pos++) {
            double x = fabs(w_value[pos]);
            if (w_index[pos] == pivot_row) where_pivot = pos;
            else if (w_index[pos] == other_row) where_other = pos;
            else if (x > cmx) cmx = x;
        }
        assert(where_pivot >= 0);
        const double xrj = w_value[where_pivot];
        if (fabs(xrj) > droptol) {
            lu->u_index[u_put] = j;
            lu->u_value[u_put] = xrj;
            u_put++;
        }
        if (where_other < 0) {
            /* fill-in goes into the slot of the pivot-row entry (no re-bucketing) */
            double x = -xrj * (other_value / pivot);
            double xabs = fabs(x);
            if (xabs > droptol) {
                w_index[where_pivot] = other_row;
                w_value[where_pivot] = x;
                w_index[put++] = j;
                if (xabs > cmx) cmx = xabs;
            } else {
                end = --w_end[j];
                w_index[where_pivot] = w_index[end];
                w_value[where_pivot] = w_value[end];
                nz = end - w_begin[j];
                blo_list_move(j, nz, lu->colcount_flink, lu->colcount_blink, m, &lu->min_colnz);
            }
        } else {
            end = --w_end[j];
            w_index[where_pivot] = w_index[end];
            w_value[where_pivot] = w_value[end];
            if (where_other == end) where_other = where_pivot;
            w_value[where_other] -= xrj * (other_value / pivot);
            double x = fabs(w_value[where_other]);
            if (x <= droptol) {
                end = --w_end[j];
                w_index[where_other] = w_index[end];
                w_value[where_other] = w_value[end];
                marked[j] = 1;
                ncancelled++;
            } else if (x > cmx) {
                cmx = x;
            }
            nz = w_end[j] - w_begin[j];
            blo_list_move(j, nz, lu->colcount_flink, lu->colcount_blink, m, &lu->min_colnz);
        }
        colmax[j] = cmx;
        ACC_COL(lu, oldnz, w_end[j] - w_begin[j]);
    }
    rend = put;
    lu->u_begin[rank + 1] = u_put;

    /* row file update, pivot.rs:1228-1293 */
    if (ncancelled) {
        assert(marked[pivot_col] == 0);
        marked[pivot_col] = 1;
        lint rput = w_begin[m + other_row];
        lint end = w_end[m + other_row];
        for (lint pos = rput; pos < end; pos++) {
            lint j = w_index[pos];
            if (marked[j]) marked[j] = 0;
            else w_index[rput++] = j;
        }
        assert(end - rput == ncancelled + 1);
        w_end[m + other_row] = rput;
    } else {
        where = w_begin[m + other_row];
        while (w_index[where] != pivot_col) { assert(where < w_end[m + other_row] - 1); where++; }
        lint end = --w_end[m + other_row];
        w_index[where] = w_index[end];
    }
    const lint nfill = rend - (rbeg + 1);
    room = w_begin[lu->w_flink[m + other_row]] - w_end[m + other_row];
    if (nfill > room) {
        nz = w_end[m + other_row] - w_begin[m + other_row];
        lint space = nfill + (lint)(stretch * (double)(nz + nfill)) + pad;
        blo_file_reappend(m + other_row, 2 * m, w_begin, w_end, lu->w_flink, lu->w_blink,
                          w_index, w_value, space);
        lu->nexpand++;
    }
    put = w_end[m + other_row];
    for (lint pos = rbeg + 1; pos < rend; pos++) w_index[put++] = w_index[pos];
    w_end[m + other_row] = put;
    nz = w_end[m + other_row] - w_begin[m + other_row];
    blo_list_move(other_row, nz, lu->rowcount_flink, lu->rowcount_blink, m, &lu->min_rownz);
    ACC_ROW(lu, other_oldnz, nz);

    /* L column, pivot.rs:1296-1305 */
    put = lu->l_begin_p[rank];
    {
        double x = other_value / pivot;
        if (fabs(x) > droptol) {
            lu->l_index[put] = other_row;
            lu->l_value[put] = x;
            put++;
        }
    }
    lu->l_index[put++] = -1;
    lu->l_begin_p[rank + 1] = put;

    colmax[pivot_col] = pivot;
    w_end[pivot_col] = cbeg;
    w_end[m + pivot_row] = rbeg;
    blo_list_remove(lu->colcount_flink, lu->colcount_blink, pivot_col);
    blo_list_remove(lu->rowcount_flink, lu->rowcount_blink, pivot_row);
    return BLO_OK;
}

/* pivot.rs:1333-1381: empty a column whose max dropped below abstol */
static void remove_col(blo_lu *lu, lint j) {
    const lint m = lu->m;
    lint *w_begin = lu->w_begin, *w_end = lu->w_end, *w_index = lu->w_index;
    const lint cbeg = w_begin[j], cend = w_end[j];
    for (lint pos = cbeg; pos < cend; pos++) {
        lint i = w_index[pos];
        lint where = w_begin[m + i];
        while (w_index[where] != j) { assert(where < w_end[m + i] - 1); where++; }
        w_index[where] = w_index[--w_end[m + i]];
        lint nz = w_end[m + i] - w_begin[m + i];
        blo_list_move(i, nz, lu->rowcount_flink, lu->rowcount_blink, m, &lu->min_rownz);
    }
    lu->col_pivot[j] = 0.0;
    w_end[j] = cbeg;
    blo_list_move(j, 0, lu->colcount_flink, lu->colcount_blink, m, &lu->min_colnz);
}

/* pivot.rs:48-112 */
int blo_pivot(blo_lu *lu) {
    const lint m = lu->m, rank = lu->rank;
    const lint pivot_col = lu->pivot_col, pivot_row = lu->pivot_row;
    const lint nz_col = lu->w_end[pivot_col] - lu->w_begin[pivot_col];
    const lint nz_row = lu->w_end[m + pivot_row] - lu->w_begin[m + pivot_row];
    double tic = blo_now();
    assert(nz_row >= 1);
    assert(nz_col >= 1);

    lint room = lu->l_mem - lu->l_begin_p[rank];
    lint need = nz_col;
    if (room < need) { lu->addmem_l = need - room; return BLO_REALLOCATE; }
    room = lu->u_mem - lu->u_begin[rank];
    need = nz_row - 1;
    if (room < need) { lu->addmem_u = need - room; return BLO_REALLOCATE; }

    int st, kind;
    if (nz_row == 1) { st = pivot_singleton_row(lu); kind = 2; }
    else if (nz_col == 1) { st = pivot_singleton_col(lu); kind = 3; }
    else if (nz_col == 2) { st = pivot_doubleton_col(lu); kind = 4; }
    else if (nz_col - 1 <= MAXROW_SMALL) { st = pivot_small(lu); kind = 5; }
    else { st = pivot_any(lu); kind = 6; }

    if (st == BLO_OK) {
        /* pivot.rs:96-106 */
        for (lint pos = lu->u_begin[rank]; pos < lu->u_begin[rank + 1]; pos++) {
            lint j = lu->u_index[pos];
            assert(j != pivot_col);
            if (lu->col_pivot[j] == 0.0 || lu->col_pivot[j] < lu->abstol) remove_col(lu, j);
        }
        blo_trace_push(lu, pivot_row, pivot_col, lu->col_pivot[pivot_col], kind, nz_row, nz_col);
        lu->elim_bytes += 12.0 * (double)nz_col + 4.0 * (double)nz_row +
                          12.0 * (double)(nz_col - 1) + 12.0 * (double)(nz_row - 1);
        lu->nelim_div += nz_col - 1;
    }
    /* the reference adds the flops even when the variant asked for reallocation
     * of W (pivot.rs:108), so a retried step counts twice; reproduced. */
    lu->factor_flops += (nz_col - 1) * (nz_row - 1);
    lu->time_elim_pivot += blo_now() - tic;
    return st;
}
