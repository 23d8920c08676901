/* blo_core.c -- CPU oracle (test infrastructure): state object, bucket lists, line files.
 * Follows /root/reference/src/lu/lu.rs, list.rs, file.rs. */
#include "blo_int.h"

/* ------------------------------------------------------------------ */
/* bucket lists: list.rs:36-137                                        */
/* flink/blink have nelem link slots followed by nlist head slots.     */
/* ------------------------------------------------------------------ */

/* list.rs:36 */
void blo_list_init(lint *flink, lint *blink, lint nelem, lint nlist, lint *min_list) {
    for (lint i = 0; i < nelem + nlist; i++) flink[i] = blink[i] = i;
    if (min_list) *min_list = nlist > 1 ? nlist : 1;
}

/* list.rs:54 -- append at the TAIL (this is what makes buckets FIFO) */
void blo_list_add(lint elem, lint list, lint *flink, lint *blink, lint nelem, lint *min_list) {
    assert(flink[elem] == elem);
    assert(blink[elem] == elem);
    lint head = nelem + list;
    lint last = blink[head];
    blink[head] = elem;
    blink[elem] = last;
    flink[last] = elem;
    flink[elem] = head;
    if (min_list && list > 0 && list < *min_list) *min_list = list;
}

/* list.rs:81 */
void blo_list_remove(lint *flink, lint *blink, lint elem) {
    flink[blink[elem]] = flink[elem];
    blink[flink[elem]] = blink[elem];
    flink[elem] = elem;
    blink[elem] = elem;
}

/* list.rs:89 */
void blo_list_move(lint elem, lint list, lint *flink, lint *blink, lint nelem, lint *min_list) {
    blo_list_remove(flink, blink, elem);
    blo_list_add(elem, list, flink, blink, nelem, min_list);
}

/* list.rs:104 */
void blo_list_swap(lint *flink, lint *blink, lint e1, lint e2) {
    lint e1next = flink[e1], e2next = flink[e2];
    lint e1prev = blink[e1], e2prev = blink[e2];
    assert(e1next != e1);
    assert(e2next != e2);
    if (e1next == e2) {
        flink[e2] = e1; blink[e1] = e2;
        flink[e1prev] = e2; blink[e2] = e1prev;
        flink[e1] = e2next; blink[e2next] = e1;
    } else if (e2next == e1) {
        flink[e1] = e2; blink[e2] = e1;
        flink[e2] = e1next; blink[e1next] = e2;
        flink[e2prev] = e1; blink[e1] = e2prev;
    } else {
        flink[e2] = e1next; blink[e1next] = e2;
        flink[e2prev] = e1; blink[e1] = e2prev;
        flink[e1prev] = e2; blink[e2] = e1prev;
        flink[e1] = e2next; blink[e2next] = e1;
    }
}

/* ------------------------------------------------------------------ */
/* line files: file.rs:32-181                                          */
/* ------------------------------------------------------------------ */

/* file.rs:32 */
void blo_file_empty(lint nlines, lint *begin, lint *end, lint *next, lint *prev, lint fmem) {
    begin[nlines] = 0;
    end[nlines] = fmem;
    for (lint i = 0; i < nlines; i++) begin[i] = end[i] = 0;
    for (lint i = 0; i < nlines; i++) { next[i] = i + 1; prev[i + 1] = i; }
    next[nlines] = 0;
    prev[0] = nlines;
}

/* file.rs:56 -- move a line to the file end (keeps in-line order), relink memory order */
void blo_file_reappend(lint line, lint nlines, lint *begin, lint *end, lint *next, lint *prev,
                       lint *index, double *value, lint extra_space) {
    lint fmem = end[nlines];
    lint used = begin[nlines];
    lint room = fmem - used;
    lint ibeg = begin[line], iend = end[line];
    begin[line] = used;
    assert(iend - ibeg <= room);
    for (lint pos = ibeg; pos < iend; pos++) {
        index[used] = index[pos];
        value[used] = value[pos];
        used++;
    }
    end[line] = used;
    room = fmem - used;
    assert(room >= extra_space);
    used += extra_space;
    begin[nlines] = used;
    blo_list_move(line, 0, next, prev, nlines, NULL);
}

/* file.rs:92 -- garbage collection; memory order and in-line order preserved */
lint blo_file_compress(lint nlines, lint *begin, lint *end, const lint *next,
                       lint *index, double *value, double stretch, lint pad) {
    lint nz = 0, used = 0, extra_space = 0;
    for (lint i = next[nlines]; i < nlines; i = next[i]) {
        lint ibeg = begin[i], iend = end[i];
        assert(ibeg >= used);
        used += extra_space;
        if (used > ibeg) used = ibeg; /* chop extra space added before */
        begin[i] = used;
        for (lint pos = ibeg; pos < iend; pos++) {
            index[used] = index[pos];
            value[used] = value[pos];
            used++;
        }
        end[i] = used;
        extra_space = (lint)(stretch * (double)(iend - ibeg)) + pad;
        nz += iend - ibeg;
    }
    assert(used <= begin[nlines]);
    used += extra_space;
    if (used > begin[nlines]) used = begin[nlines]; /* never use more space than before */
    begin[nlines] = used;
    return nz;
}

/* file.rs:151 */
lint blo_file_diff(lint nrow, const lint *begin_row, const lint *end_row,
                   const lint *begin_col, const lint *end_col,
                   const lint *index, const double *value) {
    lint ndiff = 0;
    for (lint i = 0; i < nrow; i++) {
        for (lint pos = begin_row[i]; pos < end_row[i]; pos++) {
            lint j = index[pos];
            lint where = begin_col[j];
            while (where < end_col[j] && index[where] != i) where++;
            if (where == end_col[j]) ndiff++;
            else if (value && value[pos] != value[where]) ndiff++;
        }
    }
    return ndiff;
}

/* ------------------------------------------------------------------ */
/* state object: lu.rs:243-396                                         */
/* ------------------------------------------------------------------ */

static void *zalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (!p) abort();
    return p;
}

/* lu.rs:243 */
void blo_lu_init(blo_lu *lu, lint m, lint b_nz) {
    memset(lu, 0, sizeof *lu);
    lu->l_mem = lu->u_mem = lu->w_mem = b_nz;
    lu->droptol = 1e-20;
    lu->abstol = 1e-14;
    lu->reltol = 0.1;
    lu->nzbias = 1;
    lu->maxsearch = 3;
    lu->pad = 4;
    lu->stretch = 0.3;
    lu->compress_thres = 0.5;
    lu->sparse_thres = 0.05;
    lu->search_rows = 0; /* D8: the crate's default is 0 (doc says 1) */
    lu->check_file_diff = 1;
    lu->m = m;
    lu->nupdate = -1;
    lu->pivot_row = lu->pivot_col = -1;
    lu->ftran_for_update = lu->btran_for_update = -1;

    lu->l_index = zalloc(b_nz, sizeof(lint));
    lu->u_index = zalloc(b_nz, sizeof(lint));
    lu->w_index = zalloc(b_nz, sizeof(lint));
    lu->l_value = zalloc(b_nz, sizeof(double));
    lu->u_value = zalloc(b_nz, sizeof(double));
    lu->w_value = zalloc(b_nz, sizeof(double));
    size_t n2 = (size_t)(2 * m + 2), n1 = (size_t)(m + 1);
    lu->colcount_flink = zalloc(n2, sizeof(lint));
    lu->colcount_blink = zalloc(n2, sizeof(lint));
    lu->rowcount_flink = zalloc(n2, sizeof(lint));
    lu->rowcount_blink = zalloc(n2, sizeof(lint));
    lu->w_begin = zalloc(n2, sizeof(lint));
    lu->w_end = zalloc(n2, sizeof(lint));
    lu->w_flink = zalloc(n2, sizeof(lint));
    lu->w_blink = zalloc(n2, sizeof(lint));
    lu->pivotcol = zalloc(n2, sizeof(lint));
    lu->pivotrow = zalloc(n2, sizeof(lint));
    lu->iwork1 = zalloc(n2, sizeof(lint));
    lu->pinv = zalloc(n1, sizeof(lint));
    lu->qinv = zalloc(n1, sizeof(lint));
    lu->l_begin_p = zalloc(n1, sizeof(lint));
    lu->u_begin = zalloc(n1, sizeof(lint));
    lu->l_begin = zalloc(n1, sizeof(lint));
    lu->lt_begin = zalloc(n1, sizeof(lint));
    lu->lt_begin_p = zalloc(n1, sizeof(lint));
    lu->p = zalloc(n1, sizeof(lint));
    lu->r_begin = zalloc(n1, sizeof(lint));
    lu->eta_row = zalloc(n1, sizeof(lint));
    lu->iwork0 = zalloc(n1, sizeof(lint));
    lu->pstack = zalloc(n1, sizeof(lint));
    lu->work0 = zalloc(n1, sizeof(double));
    lu->work1 = zalloc(n1, sizeof(double));
    lu->col_pivot = zalloc(n1, sizeof(double));
    lu->row_pivot = zalloc(n1, sizeof(double));
    lu->cancelled = zalloc(n1, sizeof(uint64_t));

    lu->w_end[2 * m] = lu->w_mem; /* lu.rs:308-314 */
    blo_lu_reset(lu);
}

void blo_lu_release(blo_lu *lu) {
    free(lu->l_index); free(lu->u_index); free(lu->w_index);
    free(lu->l_value); free(lu->u_value); free(lu->w_value);
    free(lu->colcount_flink); free(lu->colcount_blink);
    free(lu->rowcount_flink); free(lu->rowcount_blink);
    free(lu->w_begin); free(lu->w_end); free(lu->w_flink); free(lu->w_blink);
    free(lu->pivotcol); free(lu->pivotrow); free(lu->iwork1);
    free(lu->pinv); free(lu->qinv); free(lu->l_begin_p); free(lu->u_begin);
    free(lu->l_begin); free(lu->lt_begin); free(lu->lt_begin_p); free(lu->p);
    free(lu->r_begin); free(lu->eta_row); free(lu->iwork0); free(lu->pstack);
    free(lu->work0); free(lu->work1); free(lu->col_pivot); free(lu->row_pivot);
    free(lu->cancelled); free(lu->trace);
    memset(lu, 0, sizeof *lu);
}

/* lu.rs:329 */
void blo_lu_reset(blo_lu *lu) {
    lu->nupdate = -1;
    lu->nforrest = 0;
    lu->l_nz = lu->u_nz = lu->r_nz = 0;
    lu->min_pivot = lu->max_pivot = lu->max_eta = 0.0;
    lu->update_cost_numer = 0.0;
    lu->update_cost_denom = 1.0;
    lu->time_factorize = lu->time_solve = lu->time_update = 0.0;
    lu->l_flops = lu->u_flops = lu->r_flops = 0;
    lu->condest_l = lu->condest_u = 0.0;
    lu->norm_l = lu->norm_u = 0.0;
    lu->normest_l_inv = lu->normest_u_inv = 0.0;
    lu->onenorm = lu->infnorm = lu->residual_test = 0.0;
    lu->matrix_nz = lu->rank = lu->bump_size = lu->bump_nz = 0;
    lu->nsearch_pivot = lu->nexpand = lu->ngarbage = lu->factor_flops = 0;
    lu->time_singletons = lu->time_search_pivot = lu->time_elim_pivot = 0.0;
    lu->pivot_error = 0.0;
    lu->task = BLO_TASK_NONE;
    lu->pivot_row = lu->pivot_col = -1;
    lu->ftran_for_update = lu->btran_for_update = -1;
    lu->marker = 0;
    lu->pivotlen = 0;
    lu->rankdef = 0;
    lu->min_colnz = lu->min_rownz = 1;
    lu->w_end[2 * lu->m] = lu->w_mem;
    memset(lu->iwork0, 0, (size_t)lu->m * sizeof(lint));
    memset(lu->work0, 0, (size_t)lu->m * sizeof(double));
    lu->elim_bytes = 0.0;
    lu->nelim_div = 0;
    lu->trace_len = 0;
}

/* lu.rs:324 */
double blo_lu_update_cost(const blo_lu *lu) { return lu->update_cost_numer / lu->update_cost_denom; }

void blo_trace_push(blo_lu *lu, lint row, lint col, double pivot, int kind, lint nz_row, lint nz_col) {
    if (!lu->trace_on) return;
    if (lu->trace_len == lu->trace_cap) {
        lu->trace_cap = lu->trace_cap ? 2 * lu->trace_cap : 1024;
        lu->trace = realloc(lu->trace, (size_t)lu->trace_cap * sizeof(blo_trace));
        if (!lu->trace) abort();
    }
    blo_trace *t = &lu->trace[lu->trace_len++];
    t->row = row; t->col = col; t->pivot = pivot; t->kind = kind;
    t->nz_row = nz_row; t->nz_col = nz_col;
}
