/* blo_update.c -- CPU oracle (test infrastructure): Forrest-Tomlin update with
 * permutation shortcuts.  Follows /root/reference/src/lu/update.rs with the
 * port defects D2, D3, D4, D12 repaired (BASICLU semantics). */
#include "blo_int.h"

#define GAP (-1)
#define FLIP(i) (-(i) - 1)

/* update.rs:26-42 */
static lint find(lint j, const lint *index, lint start, lint end) {
    if (end >= 0) {
        while (start < end && index[start] != j) start++;
        return start;
    }
    while (index[start] != j && index[start] >= 0) start++;
    return index[start] == j ? start : end;
}

/* update.rs:51-105: BFS for a cycle j0 -> ... -> j0 in the row file graph */
static lint bfs_path(lint m, lint j0, const lint *begin, const lint *end, const lint *index,
                     lint *jlist, lint *marked, lint *queue) {
    lint j = -1, tail = 1, top = m;
    int found = 0;
    queue[0] = j0;
    /* D14: update.rs:68 reads `for front in 0..tail` with `tail` grown inside the loop; a Rust range is
     * evaluated once, so the reference's search never expands beyond j0 and the assert at update.rs:687
     * fires whenever the path has more than one edge.  Repaired: BASICLU's dynamic bound. */
#if BLO_REPAIR_D14
    for (lint front = 0; front < tail && !found; front++) {
#else
    const lint tail_at_entry = tail;
    for (lint front = 0; front < tail_at_entry && !found; front++) {
#endif
        j = queue[front];
        for (lint pos = begin[j]; pos < end[j]; pos++) {
            lint k = index[pos];
            if (k == j0) { found = 1; break; }
            if (marked[k] >= 0) {
                marked[k] = FLIP(j); /* parent[k] = j */
                queue[tail++] = k;
            }
        }
    }
    if (found) {
        while (j != j0) {
            jlist[--top] = j;
            j = FLIP(marked[j]);
            assert(j >= 0);
        }
        jlist[--top] = j0;
    }
    for (lint pos = 0; pos < tail; pos++) marked[queue[pos]] = 0;
    return top;
}

/* update.rs:115-162 */
static lint compress_packed(lint m, lint *begin, lint *index, double *value) {
    lint nz = 0;
    const lint end = begin[m];
    for (lint i = 0; i < m; i++) {
        lint p = begin[i];
        if (index[p] == GAP) {
            begin[i] = 0;
        } else {
            assert(index[p] > GAP);
            begin[i] = index[p];
            index[p] = GAP - i - 1;
        }
    }
    assert(index[0] == GAP);
    lint i = -1, put = 1;
    for (lint get = 1; get < end; get++) {
        if (index[get] > GAP) {
            assert(i >= 0);
            index[put] = index[get];
            value[put++] = value[get];
            nz++;
        } else if (index[get] < GAP) {
            assert(i == -1);
            i = GAP - index[get] - 1;
            index[put] = begin[i];
            begin[i] = put;
            value[put++] = value[get];
            nz++;
        } else if (i >= 0) {
            i = -1;
            index[put++] = GAP;
        }
    }
    assert(i == -1);
    begin[m] = put;
    return nz;
}

/* update.rs:176-314.  jlist has nswap+1 entries (D4 repaired). */
static void permute(blo_lu *lu, const lint *jlist, lint nswap) {
    lint *pmap = lu->pinv, *qmap = lu->qinv;
    lint *u_begin = lu->u_begin, *w_begin = lu->w_begin, *w_end = lu->w_end;
    double *col_pivot = lu->col_pivot, *row_pivot = lu->row_pivot;
    lint *u_index = lu->u_index, *w_index = lu->w_index;
    double *u_value = lu->u_value, *w_value = lu->w_value;

    const lint j0 = jlist[0], jn = jlist[nswap];
    const lint i0 = pmap[j0], in_ = pmap[jn];
    assert(nswap >= 1);
    assert(qmap[i0] == j0);
    assert(qmap[in_] == jn);
    assert(row_pivot[i0] == 0.0);
    assert(col_pivot[j0] == 0.0);

    /* row file */
    lint begin = w_begin[jn], end = w_end[jn];
    double piv = col_pivot[jn];
    for (lint n = nswap; n > 0; n--) {
        lint j = jlist[n], jprev = jlist[n - 1];
        w_begin[j] = w_begin[jprev];
        w_end[j] = w_end[jprev];
        blo_list_swap(lu->w_flink, lu->w_blink, j, jprev);
        lint where = find(j, w_index, w_begin[j], w_end[j]);
        assert(where < w_end[j]);
        if (n > 1) {
            assert(jprev != j0);
            w_index[where] = jprev;
            col_pivot[j] = w_value[where];
            assert(col_pivot[j] != 0.0);
            w_value[where] = col_pivot[jprev];
        } else {
            assert(jprev == j0);
            col_pivot[j] = w_value[where];
            assert(col_pivot[j] != 0.0);
            w_end[j]--;
            w_index[where] = w_index[w_end[j]];
            w_value[where] = w_value[w_end[j]];
        }
        lu->min_pivot = fmin(lu->min_pivot, fabs(col_pivot[j]));
        lu->max_pivot = fmax(lu->max_pivot, fabs(col_pivot[j]));
    }
    w_begin[j0] = begin;
    w_end[j0] = end;
    lint where = find(j0, w_index, w_begin[j0], w_end[j0]);
    assert(where < w_end[j0]);
    w_index[where] = jn;
    col_pivot[j0] = w_value[where];
    assert(col_pivot[j0] != 0.0);
    w_value[where] = piv;
    lu->min_pivot = fmin(lu->min_pivot, fabs(col_pivot[j0]));
    lu->max_pivot = fmax(lu->max_pivot, fabs(col_pivot[j0]));

    /* column file */
    begin = u_begin[i0];
    for (lint n = 0; n < nswap; n++) {
        lint i = pmap[jlist[n]], inext = pmap[jlist[n + 1]];
        u_begin[i] = u_begin[inext];
        where = find(i, u_index, u_begin[i], -1);
        assert(where >= 0);
        u_index[where] = inext;
        row_pivot[i] = u_value[where];
        assert(row_pivot[i] != 0.0);
        u_value[where] = row_pivot[inext];
    }
    u_begin[in_] = begin;
    where = find(in_, u_index, u_begin[in_], -1);
    assert(where >= 0);
    row_pivot[in_] = u_value[where];
    assert(row_pivot[in_] != 0.0);
    for (end = where; u_index[end] >= 0; end++) ;
    u_index[where] = u_index[end - 1];
    u_value[where] = u_value[end - 1];
    u_index[end - 1] = -1;

    /* mappings */
    for (lint n = nswap; n > 0; n--) {
        lint j = jlist[n], i = pmap[jlist[n - 1]];
        pmap[j] = i;
        qmap[i] = j;
    }
    pmap[j0] = in_;
    qmap[in_] = j0;
}

/* update.rs:388-959 */
int blo_k_update(blo_lu *lu, double xtbl) {
    const lint m = lu->m, nforrest = lu->nforrest, pad = lu->pad;
    const double stretch = lu->stretch;
    lint u_nz = lu->u_nz;
    lint *pmap = lu->pinv, *qmap = lu->qinv;
    lint *u_begin = lu->u_begin, *r_begin = lu->r_begin;
    lint *w_begin = lu->w_begin, *w_end = lu->w_end, *w_flink = lu->w_flink, *w_blink = lu->w_blink;
    lint *l_index = lu->l_index, *u_index = lu->u_index, *w_index = lu->w_index;
    double *l_value = lu->l_value, *u_value = lu->u_value, *w_value = lu->w_value;
    lint *marked = lu->iwork0;
    lint *iwork1 = lu->iwork1, *iwork2 = lu->iwork1 + m;
    double *work1 = lu->work1;

    const lint jpivot = lu->btran_for_update;
    const lint ipivot = pmap[jpivot];
    const double oldpiv = lu->col_pivot[jpivot];
    lint ipivot_vec = ipivot, jpivot_vec = jpivot; /* D2 repaired: one-element reach */
#if !BLO_REPAIR_D2
    ipivot_vec = 0; jpivot_vec = 0;      /* update.rs:422-423: vec![0; ipivot] -- zeros (and a panic when the length is 0) */
    if (ipivot == 0 || jpivot == 0) BLO_DEFECT_TRAP("D2", "update.rs:877-878 indexes an empty vec![0; 0] (panic)");
#endif
    double tic = blo_now();
    assert(nforrest < m);

    /* move the diagonal element to the end of the spike, update.rs:442-463 */
    double spike_diag = 0.0;
    int have_diag = 0;
    lint put = u_begin[m];
    for (lint pos = put; u_index[pos] >= 0; pos++) {
        lint i = u_index[pos];
        if (i != ipivot) {
            u_index[put] = i;
            u_value[put++] = u_value[pos];
        } else {
            spike_diag = u_value[pos];
            have_diag = 1;
        }
    }
    if (have_diag) {
        u_index[put] = ipivot;
        u_value[put] = spike_diag;
    }
    const lint nz_spike = put - u_begin[m];
    const lint nz_roweta = r_begin[nforrest + 1] - r_begin[nforrest];

    /* new pivot, update.rs:485-513 */
    lint marker = ++lu->marker;
    for (lint pos = r_begin[nforrest]; pos < r_begin[nforrest + 1]; pos++) {
        lint i = l_index[pos];
        marked[i] = marker;
        work1[i] = l_value[pos];
    }
    double newpiv = spike_diag;
    lint intersect = 0;
    for (lint pos = u_begin[m]; pos < u_begin[m] + nz_spike; pos++) {
        lint i = u_index[pos];
        assert(i != ipivot);
        if (marked[i] == marker) {
            newpiv -= u_value[pos] * work1[i];
            intersect++;
        }
    }
    if (newpiv == 0.0 || fabs(newpiv) < lu->abstol) return BLO_ERROR_SINGULAR_UPDATE;
    const double piverr = fabs(newpiv - xtbl * oldpiv);

    /* bound on file growth, update.rs:517-536 */
    lint grow = 0;
    for (lint pos = u_begin[m]; pos < u_begin[m] + nz_spike; pos++) {
        lint i = u_index[pos];
        lint j = qmap[i];
        lint jnext = w_flink[j];
        if (w_end[j] == w_begin[jnext]) {
            lint nz = w_end[j] - w_begin[j];
            grow += nz + 1;
            grow += (lint)(stretch * (double)(nz + 1)) + pad;
        }
    }
    lint room = w_end[m] - w_begin[m];
    if (grow > room) { lu->addmem_w = grow - room; return BLO_REALLOCATE; }

    /* remove column jpivot from the row file, update.rs:538-555 */
    lint nz = 0;
    for (lint pos = u_begin[ipivot]; u_index[pos] >= 0; pos++) {
        lint j = qmap[u_index[pos]];
        lint end = w_end[j]--;
        lint where = find(jpivot, w_index, w_begin[j], end);
        assert(where < end);
        w_index[where] = w_index[end - 1];
        w_value[where] = w_value[end - 1];
        nz++;
    }
    u_nz -= nz;
    /* erase column jpivot in the column file, update.rs:557-563 */
    for (lint pos = u_begin[ipivot]; u_index[pos] >= 0; pos++) u_index[pos] = GAP;
    /* column pointer to the spike, chop the diagonal, update.rs:565-570 */
    u_begin[ipivot] = u_begin[m];
    u_begin[m] += nz_spike;
    u_index[u_begin[m]++] = GAP;
    /* insert the spike into the row file, update.rs:572-601 */
    for (lint pos = u_begin[ipivot]; u_index[pos] >= 0; pos++) {
        lint j = qmap[u_index[pos]];
        lint jnext = w_flink[j];
        if (w_end[j] == w_begin[jnext]) {
            nz = w_end[j] - w_begin[j];
            lint space = 1 + (lint)(stretch * (double)(nz + 1)) + pad;
            blo_file_reappend(j, m, w_begin, w_end, w_flink, w_blink, w_index, w_value, space);
        }
        lint end = w_end[j]++;
        w_index[end] = jpivot;
        w_value[end] = u_value[pos];
    }
    u_nz += nz_spike;
    lu->col_pivot[jpivot] = spike_diag;
    lu->row_pivot[ipivot] = spike_diag;

    /* triangularity test, update.rs:609-818 */
    int istriangular;
    lint nreach = 0;
    lint *row_reach = NULL, *col_reach = NULL;
    if (have_diag) {
        istriangular = intersect == 0;
        if (istriangular) {
            lu->min_pivot = fmin(lu->min_pivot, fabs(newpiv));
            lu->max_pivot = fmax(lu->max_pivot, fabs(newpiv));
#if !BLO_REPAIR_D3
            BLO_DEFECT_TRAP("D3", "update.rs:634-643 indexes reach vectors of nreach-1 elements at nreach-1 (panic)");
#endif
            nreach = nz_roweta + 1; /* D3 repaired: nreach elements */
            row_reach = iwork1;
            col_reach = iwork2;
            row_reach[0] = ipivot;
            col_reach[0] = jpivot;
            lint pos = r_begin[nforrest];
            for (lint n = 1; n < nreach; n++) {
                lint i = l_index[pos++];
                row_reach[n] = i;
                col_reach[n] = qmap[i];
            }
            lu->nsymperm_total++;
        }
    } else {
        lint *path = iwork1, *reach = iwork2;
        lint *pstack = lu->pstack;
        lint top = bfs_path(m, jpivot, w_begin, w_end, w_index, path, marked, iwork2);
        assert(top < m - 1);
        assert(path[top] == jpivot);

        istriangular = 1;
        lint rtop = m;
        marker = ++lu->marker;
        for (lint t = top; t < m - 1 && istriangular; t++) {
            lint j = path[t], jnext = path[t + 1];
            lint where = find(jnext, w_index, w_begin[j], w_end[j]);
            assert(where < w_end[j]);
            w_index[where] = j; /* take the path edge out for a moment */
            rtop = blo_dfs(j, w_begin, w_end, w_index, rtop, reach, pstack, marked, marker);
            assert(reach[rtop] == j);
            reach[rtop] = jnext;
            w_index[where] = jnext;
            istriangular = marked[jnext] != marker;
        }
        if (istriangular) {
            lint j = path[m - 1];
            rtop = blo_dfs(j, w_begin, w_end, w_index, rtop, reach, pstack, marked, marker);
            assert(reach[rtop] == j);
            reach[rtop] = jpivot;
            marked[j]--; /* unmark for a moment */
            for (lint pos = u_begin[ipivot]; u_index[pos] >= 0; pos++)
                if (marked[qmap[u_index[pos]]] == marker) istriangular = 0;
            marked[j]++;
        }
        if (istriangular) {
            lint nswap = m - top - 1;
#if !BLO_REPAIR_D4
            BLO_DEFECT_TRAP("D4", "update.rs:797 hands permute() a slice of nswap entries, it reads jlist[nswap] (panic)");
#endif
            permute(lu, path + top, nswap); /* D4 repaired: nswap+1 entries visible */
            u_nz--;
            assert(reach[rtop] == jpivot);
            col_reach = reach + rtop;
            row_reach = iwork1 + rtop;
            nreach = m - rtop;
            for (lint n = 0; n < nreach; n++) row_reach[n] = pmap[col_reach[n]];
        }
    }

    /* Forrest-Tomlin update, update.rs:822-883 */
    if (!istriangular) {
        for (lint pos = w_begin[jpivot]; pos < w_end[jpivot]; pos++) {
            lint j = w_index[pos];
            assert(j != jpivot);
            lint where = -1, end;
            for (end = u_begin[pmap[j]]; u_index[end] >= 0; end++)
                if (u_index[end] == ipivot) where = end;
            assert(where >= 0);
            u_index[where] = u_index[end - 1];
            u_value[where] = u_value[end - 1];
            u_index[end - 1] = -1;
            u_nz--;
        }
        w_end[jpivot] = w_begin[jpivot];
        lu->col_pivot[jpivot] = newpiv;
        lu->row_pivot[ipivot] = newpiv;
        lu->min_pivot = fmin(lu->min_pivot, fabs(newpiv));
        lu->max_pivot = fmax(lu->max_pivot, fabs(newpiv));

        nz = 0;
        put = r_begin[nforrest];
        double max_eta = 0.0;
        for (lint pos = put; pos < r_begin[nforrest + 1]; pos++) {
            if (l_value[pos] != 0.0) {
                max_eta = fmax(max_eta, fabs(l_value[pos]));
                l_index[put] = l_index[pos];
                l_value[put++] = l_value[pos];
                nz++;
            }
        }
        r_begin[nforrest + 1] = put;
        lu->r_nz += nz;
        lu->max_eta = fmax(lu->max_eta, max_eta);

        nreach = 1;
        row_reach = &ipivot_vec;
        col_reach = &jpivot_vec;
        lu->nforrest++;
        lu->nforrest_total++;
    }

    /* append the reach to the pivot sequence, update.rs:891-911 */
    if (lu->pivotlen + nreach > 2 * m) blo_garbage_perm(lu);
    put = lu->pivotlen;
    for (lint n = 0; n < nreach; n++) lu->pivotrow[put++] = row_reach[n];
    put = lu->pivotlen;
    for (lint n = 0; n < nreach; n++) lu->pivotcol[put++] = col_reach[n];
    lu->pivotlen += nreach;

    /* compress U and W when enough was wasted, update.rs:915-937 (D12: signed) */
    lint used = u_begin[m];
#if BLO_REPAIR_D12
    if (used - u_nz - m > (lint)(lu->compress_thres * (double)used)) {
#else
    if ((uint64_t)(used - u_nz - m) > (uint64_t)(lu->compress_thres * (double)used)) {      /* usize arithmetic wraps (release build) */
#endif
        nz = compress_packed(m, u_begin, u_index, u_value);
        assert(nz == u_nz);
    }
    used = w_begin[m];
    lint need = u_nz + (lint)(stretch * (double)u_nz) + m * pad;
#if BLO_REPAIR_D12
    if (used - need > (lint)(lu->compress_thres * (double)used)) {
#else
    if ((uint64_t)(used - need) > (uint64_t)(lu->compress_thres * (double)used)) {
#endif
        nz = blo_file_compress(m, w_begin, w_end, w_flink, w_index, w_value, stretch, pad);
        assert(nz == u_nz);
    }

    double el = blo_now() - tic;
    lu->time_update += el;
    lu->time_update_total += el;
    lu->pivot_error = piverr / (1.0 + fabs(newpiv));
    lu->u_nz = u_nz;
    lu->btran_for_update = -1;
    lu->ftran_for_update = -1;
    lu->update_cost_numer += (double)nz_roweta;
    lu->nupdate++;
    lu->nupdate_total++;
    return BLO_OK;
}

/* update.rs:49-55 */
int blo_lu_update(blo_lu *lu, double xtbl) {
    if (lu->nupdate < 0 || lu->ftran_for_update < 0 || lu->btran_for_update < 0)
        return BLO_ERROR_INVALID_CALL;
#if BLO_REPAIR_D7
    /* D7 repair (lu_load semantics): refresh the file-size sentinel of the m-line file */
    lu->addmem_l = lu->addmem_u = lu->addmem_w = 0;
    lu->w_end[lu->m] = lu->w_mem;
#endif
    return blo_k_update(lu, xtbl);
}
