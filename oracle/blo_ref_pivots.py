"""Second, independent restatement of the pivot-choosing half of rwl/blu, written directly from the Rust
sources (NOT from oracle/*.c): singleton peel (src/lu/singletons.rs:287-503), bump set-up
(src/lu/setup_bump.rs:100-264), Markowitz search incl. the row search (src/lu/markowitz.rs:34-219) and the
five elimination variants + remove_col (src/lu/pivot.rs:48-1381), driven as factorize_bump.rs:12-49 does.

TEST INFRASTRUCTURE ONLY (see oracle/blo.h).  Its purpose: the C oracle and the CUDA path were written by one
reader of the Rust; this file is a second reading with different data structures (Python lists per line,
ordered dicts per count bucket instead of the flink/blink arrays and the line file), so that a shared
misreading of tie-breaking, storage order or bucket order shows up as a different pivot sequence
(tests/test_oracle.py cross-checks >= 1000 random and structured cases).

Defects of the Rust port are resolved as in SURVEY.md section 0: D5 (64-bit cancellation mask) and D6 (a column
whose maximum is below abstol is passed over) follow BASICLU; everything else is as written.
"""
from collections import OrderedDict

MAXROW_SMALL = 64      # pivot.rs:22


class Buckets:
    """list.rs:36-137: elements in at most one of the lists 0..nlist-1, FIFO by insertion (list_add appends)."""

    def __init__(self):
        self.lists = {}
        self.where = {}

    def add(self, e, k):
        assert e not in self.where
        self.lists.setdefault(k, OrderedDict())[e] = None
        self.where[e] = k

    def remove(self, e):
        k = self.where.pop(e, None)
        if k is not None:
            del self.lists[k][e]

    def move(self, e, k):          # list_move: remove, then append at the tail
        self.remove(e)
        self.add(e, k)

    def items(self, k):
        return list(self.lists.get(k, ()))

    def first(self, k):
        for e in self.lists.get(k, ()):
            return e
        return None


class RefLU:
    def __init__(self, m, colptr, rowidx, values, abstol=1e-14, reltol=0.1, droptol=1e-20, maxsearch=3,
                 search_rows=0, nzbias=1):
        self.m = m
        self.abstol, self.reltol, self.droptol = abstol, reltol, droptol
        self.maxsearch, self.search_rows, self.nzbias = maxsearch, search_rows, nzbias
        self.bcols = [[(int(rowidx[p]), float(values[p])) for p in range(colptr[j], colptr[j + 1])] for j in range(m)]
        self.pinv = [-1] * m
        self.qinv = [-1] * m
        self.pivots = []          # (row, col, pivot value) in rank order
        self.dropped = []         # columns removed as rank deficiency, in order
        self.colpiv = [0.0] * m
        self.nsearch = 0

    # ------------------------------------------------------------ singletons.rs
    def singletons(self):
        m = self.m
        # row-wise copy filled by scanning j = 0..m (singletons.rs:186-198): ascending column inside a row
        self.brows = [[] for _ in range(m)]
        for j in range(m):
            for i, x in self.bcols[j]:
                self.brows[i].append((j, x))
        if self.nzbias >= 0:
            self.singleton_cols(); self.singleton_rows()
        else:
            self.singleton_rows(); self.singleton_cols()

    def singleton_cols(self):          # singletons.rs:287-393
        m = self.m
        cnt = {j: len(self.bcols[j]) for j in range(m) if self.qinv[j] < 0}
        rowset = {}
        for j in cnt:
            x = 0
            for i, _ in self.bcols[j]:
                x ^= i
            rowset[j] = x
        queue = [j for j in range(m) if j in cnt and cnt[j] == 1]
        front = 0
        while front < len(queue):
            j = queue[front]; front += 1
            if cnt[j] == 0:
                continue
            i = rowset[j]
            piv = next(x for (jj, x) in self.brows[i] if jj == j)
            if piv == 0.0 or abs(piv) < self.abstol:
                continue
            rank = len(self.pivots)
            self.qinv[j] = rank; self.pinv[i] = rank
            del cnt[j]
            for j2, _ in self.brows[i]:
                if self.qinv[j2] < 0:
                    rowset[j2] ^= i
                    cnt[j2] -= 1
                    if cnt[j2] == 1:
                        queue.append(j2)
            self.pivots.append((i, j, piv))
            self.colpiv[j] = piv

    def singleton_rows(self):          # singletons.rs:398-503
        m = self.m
        cnt = {i: len(self.brows[i]) for i in range(m) if self.pinv[i] < 0}
        colset = {}
        for i in cnt:
            x = 0
            for j, _ in self.brows[i]:
                x ^= j
            colset[i] = x
        queue = [i for i in range(m) if i in cnt and cnt[i] == 1]
        front = 0
        while front < len(queue):
            i = queue[front]; front += 1
            if cnt[i] == 0:
                continue
            j = colset[i]
            piv = next(x for (ii, x) in self.bcols[j] if ii == i)
            if piv == 0.0 or abs(piv) < self.abstol:
                continue
            rank = len(self.pivots)
            self.qinv[j] = rank; self.pinv[i] = rank
            del cnt[i]
            for i2, _ in self.bcols[j]:
                if self.pinv[i2] < 0:
                    colset[i2] ^= j
                    cnt[i2] -= 1
                    if cnt[i2] == 1:
                        queue.append(i2)
            self.pivots.append((i, j, piv))
            self.colpiv[j] = piv

    # ------------------------------------------------------------ setup_bump.rs
    def setup_bump(self):
        m = self.m
        self.col = {}                  # j -> [[i, x], ...] in storage order
        self.row = {}                  # i -> [j, ...]
        self.colmax = {}
        self.cb, self.rb = Buckets(), Buckets()
        for j in range(m):
            if self.qinv[j] >= 0:
                continue
            ent = [[i, x] for (i, x) in self.bcols[j] if self.pinv[i] < 0]
            cmx = max([abs(x) for _, x in ent], default=0.0)
            if cmx == 0.0 or cmx < self.abstol:
                self.colmax[j] = 0.0
                self.col[j] = []
                self.cb.add(j, 0)
            else:
                self.colmax[j] = cmx
                self.col[j] = ent
                self.cb.add(j, len(ent))
        for i in range(m):
            if self.pinv[i] >= 0:
                continue
            self.row[i] = []
        for j in range(m):                      # rows filled by scanning the column file in index order
            for i, _ in self.col.get(j, ()):
                self.row[i].append(j)
        for i in range(m):
            if i in self.row:
                self.rb.add(i, len(self.row[i]))

    # ------------------------------------------------------------ markowitz.rs
    def markowitz(self):
        m = self.m
        e = self.cb.first(0)
        if e is not None:                        # markowitz.rs:73-78
            return None, e
        best = None                              # (mc, i, j)
        mc64 = m * m
        nsearch = 0
        maxcount = max([k for k, v in self.cb.lists.items() if v] + [k for k, v in self.rb.lists.items() if v and k <= m] + [1])
        for nz in range(1, maxcount + 1):
            for j in self.cb.items(nz):
                cmx = self.colmax[j]
                if cmx == 0.0 or cmx < self.abstol:
                    continue                     # D6: passed over, not counted
                tol = max(self.abstol, self.reltol * cmx)
                for i, x in self.col[j]:
                    ax = abs(x)
                    if ax == 0.0 or ax < tol:
                        continue
                    mc = (nz - 1) * (len(self.row[i]) - 1)
                    if mc < mc64:
                        mc64 = mc; best = (i, j)
                        if self.search_rows and mc64 <= (nz - 1) * (nz - 1):
                            self.nsearch += nsearch
                            return best
                nsearch += 1
                if nsearch >= self.maxsearch:
                    self.nsearch += nsearch
                    return best
            if not self.search_rows:
                continue
            for i in self.rb.items(nz):          # (a copy: parking below changes the list, markowitz.rs:137)
                cheap = found = False
                for j in self.row[i]:
                    mc = (nz - 1) * (len(self.col[j]) - 1)
                    if mc >= mc64:
                        continue
                    cheap = True
                    cmx = self.colmax[j]
                    if cmx == 0.0 or cmx < self.abstol:
                        continue
                    ax = abs(next(x for (ii, x) in self.col[j] if ii == i))
                    if ax >= self.abstol and ax >= self.reltol * cmx:
                        found = True
                        mc64 = mc; best = (i, j)
                        if mc64 <= nz * (nz - 1):
                            self.nsearch += nsearch
                            return best
                if cheap and not found:
                    self.rb.move(i, m + 1)       # parked until the row is updated (markowitz.rs:178-179)
                else:
                    nsearch += 1
                    if nsearch >= self.maxsearch:
                        self.nsearch += nsearch
                        return best
        self.nsearch += nsearch
        return best

    # ------------------------------------------------------------ pivot.rs
    def pivot(self, pr, pc):
        nz_col, nz_row = len(self.col[pc]), len(self.row[pr])
        if nz_row == 1:
            urow = self.pivot_singleton_row(pr, pc)
        elif nz_col == 1:
            urow = self.pivot_singleton_col(pr, pc)
        elif nz_col == 2:
            urow = self.pivot_doubleton_col(pr, pc)
        else:
            urow = self.pivot_general(pr, pc, small=(nz_col - 1 <= MAXROW_SMALL))
        for j in urow:                           # pivot.rs:96-106
            if self.colmax[j] == 0.0 or self.colmax[j] < self.abstol:
                self.remove_col(j)

    def finish(self, pr, pc, pivot):
        self.colpiv[pc] = pivot
        self.col[pc] = []; self.row[pr] = []
        self.cb.remove(pc); self.rb.remove(pr)
        rank = len(self.pivots)
        self.pinv[pr] = rank; self.qinv[pc] = rank
        self.pivots.append((pr, pc, pivot))

    def pivot_general(self, pr, pc, small):      # pivot_any pivot.rs:114-458, pivot_small :460-833
        C = self.col[pc]
        w = next(k for k, (i, _) in enumerate(C) if i == pr)
        C[0], C[w] = C[w], C[0]                  # pivot to the front of its column
        pivot = C[0][1]
        R = self.row[pr]
        w = R.index(pc)
        R[0], R[w] = R[w], R[0]
        crow = [i for i, _ in C[1:]]
        cval = [x for _, x in C[1:]]
        position = {i: p for p, i in enumerate(crow)}
        cancelled = {}                           # column -> set of positions whose update cancelled (small)
        urow = []
        for j in R[1:]:
            work = [0.0] * len(crow)
            keep = []
            where = None
            cmx = 0.0
            for i, x in self.col[j]:
                p = position.get(i)
                if p is not None:
                    work[p] = x
                else:
                    if i == pr:
                        where = len(keep)
                    elif abs(x) > cmx:
                        cmx = abs(x)
                    keep.append([i, x])
            keep[0], keep[where] = keep[where], keep[0]
            xrj = keep[0][1]
            a = xrj / pivot
            drop = set()
            for p in range(len(crow)):
                v = work[p] - a * cval[p]
                if small and not abs(v) > self.droptol:
                    drop.add(p)
                    continue
                keep.append([crow[p], v])
                if abs(v) > cmx:
                    cmx = abs(v)
            cancelled[j] = drop
            if abs(xrj) > self.droptol:
                urow.append(j)
            self.col[j] = keep[1:]               # the pivot-row entry leaves the column
            self.cb.move(j, len(self.col[j]))
            self.colmax[j] = cmx
        inrow = set(R)
        for p, i in enumerate(crow):
            new = [j for j in self.row[i] if j not in inrow]
            new += [j for j in R[1:] if p not in cancelled[j]]
            self.row[i] = new
            self.rb.move(i, len(new))
        self.finish(pr, pc, pivot)
        return urow

    def pivot_singleton_row(self, pr, pc):       # pivot.rs:835-926
        C = self.col[pc]
        pivot = next(x for i, x in C if i == pr)
        for i, _ in C:
            if i == pr:
                continue
            r = self.row[i]
            w = r.index(pc)
            r[w] = r[-1]; r.pop()                # move-last-into-hole
            self.rb.move(i, len(r))
        self.finish(pr, pc, pivot)
        return []

    def pivot_singleton_col(self, pr, pc):       # pivot.rs:928-1025
        pivot = self.col[pc][0][1]
        urow = []
        for j in list(self.row[pr]):
            if j == pc:
                continue
            c = self.col[j]
            w = next(k for k, (i, _) in enumerate(c) if i == pr)
            xrj = c[w][1]
            cmx = max([abs(x) for k, (i, x) in enumerate(c) if k != w], default=0.0)
            if abs(xrj) > self.droptol:
                urow.append(j)
            c[w] = c[-1]; c.pop()
            self.cb.move(j, len(c))
            self.colmax[j] = cmx
        self.finish(pr, pc, pivot)
        return urow

    def pivot_doubleton_col(self, pr, pc):       # pivot.rs:1027-1331
        C = self.col[pc]
        if C[0][0] != pr:
            C[0], C[1] = C[1], C[0]
        pivot = C[0][1]
        other_row, other_value = C[1]
        R = self.row[pr]
        w = R.index(pc)
        R[0], R[w] = R[w], R[0]
        urow, fill, cancelled = [], [], set()
        for j in R[1:]:
            c = self.col[j]
            wp = wo = None
            cmx = 0.0
            for k, (i, x) in enumerate(c):
                if i == pr:
                    wp = k
                elif i == other_row:
                    wo = k
                elif abs(x) > cmx:
                    cmx = abs(x)
            xrj = c[wp][1]
            if abs(xrj) > self.droptol:
                urow.append(j)
            if wo is None:
                x = -xrj * (other_value / pivot)
                if abs(x) > self.droptol:
                    c[wp] = [other_row, x]       # fill-in takes the slot of the pivot-row entry; no re-bucketing
                    fill.append(j)
                    if abs(x) > cmx:
                        cmx = abs(x)
                else:
                    c[wp] = c[-1]; c.pop()
                    self.cb.move(j, len(c))
            else:
                last = len(c) - 1
                c[wp] = c[last]; c.pop()
                if wo == last:
                    wo = wp
                c[wo][1] -= xrj * (other_value / pivot)
                ax = abs(c[wo][1])
                if ax <= self.droptol:
                    c[wo] = c[-1]; c.pop()
                    cancelled.add(j)
                elif ax > cmx:
                    cmx = ax
                self.cb.move(j, len(c))
            self.colmax[j] = cmx
        r = self.row[other_row]
        if cancelled:
            gone = cancelled | {pc}
            r[:] = [j for j in r if j not in gone]
        else:
            w = r.index(pc)
            r[w] = r[-1]; r.pop()
        r.extend(fill)
        self.rb.move(other_row, len(r))
        self.finish(pr, pc, pivot)
        return urow

    def remove_col(self, j):                     # pivot.rs:1333-1381
        for i, _ in self.col[j]:
            r = self.row[i]
            w = r.index(j)
            r[w] = r[-1]; r.pop()
            self.rb.move(i, len(r))
        self.colmax[j] = 0.0
        self.col[j] = []
        self.cb.move(j, 0)

    # ------------------------------------------------------------ factorize_bump.rs
    def factorize(self):
        self.singletons()
        self.setup_bump()
        m = self.m
        while len(self.pivots) + len(self.dropped) < m:
            got = self.markowitz()
            if got is None:
                raise AssertionError("no eligible pivot (factorize_bump.rs:22 asserts)")
            pr, pc = got
            if pr is None:                       # empty column: rank deficiency (factorize_bump.rs:23-31)
                self.cb.remove(pc)
                self.dropped.append(pc)
                continue
            self.pivot(pr, pc)
        return self

    def permutations(self):
        """rowperm / colperm as get_factors.rs:17-20 defines them (build_factors.rs:192-209: non-pivotal rows and
        columns follow in index order)."""
        rows = [p[0] for p in self.pivots] + [i for i in range(self.m) if self.pinv[i] < 0]
        cols = [p[1] for p in self.pivots] + [j for j in range(self.m) if self.qinv[j] < 0]
        return rows, cols
