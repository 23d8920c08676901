"""Single-basis configurations of BASELINE.json (configs[0], [2], [3], [4]) on one B200, each next to
the CPU oracle (C restatement of rwl/blu, one thread) timed in the same run on the box's host.
Not the headline bench (that is bench.py on configs[1]); these are SURVEY.md 8(d)'s other rows.
Every GPU result is also checked against the oracle (bit-exact factors / solutions).

usage: python scripts/configs_bench.py [--quick] [--out gpurun_out/configs.jsonl] [c1 c3 c4 c5 ...]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from blu_b200 import BLU, gen  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

HBM_PEAK = 6542.1
try:
    HBM_PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def wall(f):
    t = time.perf_counter()
    r = f()
    return r, time.perf_counter() - t


def factor_bytes(g, m):
    """algorithmic bytes of one factorization, SURVEY.md 8(d) (same formula as bench.py)"""
    nnz, lnz, unz = g.info("matrix_nz"), g.info("l_nz"), g.info("u_nz")
    bump_nz, bump = g.info("bump_nz"), g.info("bump_size")
    return 36.0 * nnz + 12.0 * (lnz + unz) + 16.0 * m + 28.0 * bump_nz + 16.0 * bump + g.info("elim_bytes") + 48.0 * (lnz + unz) + 40.0 * m


def same_factors(g, o):
    _, fo = o.get_factors()
    _, fg = g.get_factors()
    return all(np.array_equal(fo[k], fg[k]) for k in fo)


def make_pair(m, nnz, threads, mem=None, ofactor=None):
    g = BLU(m, nnz)
    g.threads_per_basis = threads
    if mem:
        g.l_mem, g.u_mem, g.w_mem = mem
    o = Oracle(m, ofactor or (mem[0] if mem else 60 * nnz))
    return g, o


def run_factorize(name, cp, ri, v, m, threads, mem, rhs_seed, check_factors=True, file_diff=True, note=""):
    g, o = make_pair(m, len(v), threads, mem, ofactor=max(mem) if mem else None)
    o.set_param("check_file_diff", 1 if file_diff else 0)
    g.factorize(cp[:-1], cp[1:], ri, v)                       # warm-up (module load, Reallocate growth)
    sg, tg = wall(lambda: g.factorize(cp[:-1], cp[1:], ri, v))
    kern_ms = 1e3 * g.info("time_factorize") / 2 if g.info("nrealloc") == 0 else None
    so, to = wall(lambda: o.factorize(cp[:-1], cp[1:], ri, v))
    b = gen.rhs(rhs_seed, m)
    (ss, xg), tsg = wall(lambda: g.solve_dense(b, "N"))
    (_, xo), tso = wall(lambda: o.solve_dense(b, "N"))
    byt = factor_bytes(g, m)
    rec = {"config": name, "m": m, "nnz": int(len(v)), "threads_per_basis": threads, "note": note,
           "gpu": {"factorize_ms_wall": 1e3 * tg, "factorize_ms_device_avg": kern_ms, "solve_dense_ms_wall": 1e3 * tsg,
                   "status": sg, "nrealloc": g.info("nrealloc")},
           "cpu_oracle_1thread": {"factorize_ms": 1e3 * to, "solve_dense_ms": 1e3 * tso, "status": so,
                                  "file_diff_asserts": bool(file_diff)},
           "stats": {k: g.info(k) for k in ("rank", "bump_size", "bump_nz", "l_nz", "u_nz", "factor_flops", "nsearch_pivot", "ngarbage", "nexpand")},
           "roofline": {"bound": "hbm (single basis: latency-bound, one CTA = 1/148 of the GPU)", "algorithmic_bytes": byt,
                        "achieved_gbs": byt / tg / 1e9, "peak_gbs": HBM_PEAK, "frac": byt / tg / 1e9 / HBM_PEAK},
           "parity": {"status_equal": sg == so, "factors_bit_identical": same_factors(g, o) if check_factors else None,
                      "solve_dense_bit_identical": bool(np.array_equal(xg, xo)),
                      "stats_equal": all(g.info(k) == o.info(k) for k in ("rank", "l_nz", "u_nz", "factor_flops", "nsearch_pivot"))},
           "speedup_gpu_over_cpu_factorize": to / tg}
    return rec, g, o


def c1(args):
    (cp, ri, v), rhs = gen.config1()
    rec, g, o = run_factorize("configs[0]: 1000x1000, ~5 nnz/col, 30% slack", cp, ri, v, 1000, 256, (60000, 60000, 200000), 1002)
    return [rec]


def c3(args):
    m, bump = (100000, 4000) if not args.quick else (20000, 1000)
    cp, ri, v = gen.config3(m, bump)
    dense = bump * bump
    mem = (int(2.2 * (dense // 2 + 10 * m)), int(1.2 * (dense // 2 + 10 * m)), int(3.0 * dense + 40 * m))
    rec, g, o = run_factorize(f"configs[2]: {m}x{m}, structured bump {bump}x{bump}, ~8 nnz/col", cp, ri, v, m, 1024, mem, 4002,
                              file_diff=False, note="oracle without the O(sum rownz*colnz) file_diff asserts (D11), favourable to the CPU")
    return [rec]


def c4(args):
    n = 300 if not args.quick else 100
    m = n * n
    cp, ri, v = gen.config4(n, rail_deg=max(50, m // 200))
    mem = (60 * len(v), 60 * len(v), 120 * len(v))
    rec, g, o = run_factorize(f"configs[3] scaled: circuit-like {n}x{n} grid (m={m}); full size is 1000x1000", cp, ri, v, m, 1024, mem, 5002,
                              file_diff=False, note="scaled down: the Markowitz search of the device path scans the active columns (O(bump) per pivot), which does not scale to a 10^6 bump yet")
    # Gilbert-Peierls solves: 0.1 % dense right-hand sides, both systems
    nrhs = args.nrhs or (1000 if not args.quick else 100)
    nz = max(1, m // 1000)
    tg = to = 0.0
    same = True
    nzl = 0
    for r in range(nrhs):
        idx, val = gen.sparse_rhs_np(6000 + r, m, nz)
        tr = "N" if r % 2 == 0 else "T"
        _, dt = wall(lambda: g.solve_sparse(nz, idx, val, tr)); tg += dt
        _, dt = wall(lambda: o.solve_sparse(nz, idx, val, tr)); to += dt
        n_ = o.nzlhs
        nzl += n_
        same = same and g.nzlhs == n_ and np.array_equal(g.ilhs[:n_], o.ilhs[:n_]) and np.array_equal(g.lhs, o.lhs)
    # the same right-hand sides, all at once, through the multi-RHS dense solve
    R = np.zeros((nrhs, m))
    for r in range(nrhs):
        idx, val = gen.sparse_rhs_np(6000 + r, m, nz)
        R[r, idx] = val
    g.solve_dense_multi(R[:8], "N")                              # warm-up / staging allocation
    (sm, X), tm = wall(lambda: g.solve_dense_multi(R, "N"))
    ns = min(nrhs, 50)
    tcs = 0.0
    same_multi = sm == 0
    for r in range(ns):
        (_, xo), dt = wall(lambda: o.solve_dense(R[r], "N")); tcs += dt
        same_multi = same_multi and np.array_equal(X[r], xo)
    rec["solve_dense_multi"] = {"nrhs": nrhs, "gpu_ms_total": 1e3 * tm, "gpu_ms_per_rhs": 1e3 * tm / nrhs,
                                "cpu_oracle_ms_per_rhs": 1e3 * tcs / ns, "bit_identical_to_oracle_solve_dense": bool(same_multi),
                                "note": "one warp per right-hand side, all in flight; includes H2D of rhs and D2H of the solutions"}
    # and through the multi-RHS SPARSE solve: the Gilbert-Peierls solve of every right-hand side, one warp each
    rl = [gen.sparse_rhs_np(6000 + r, m, nz) for r in range(nrhs)]
    g.solve_sparse_multi(rl[:8], "N")
    (ssm, outm, statm), tsm = wall(lambda: g.solve_sparse_multi(rl, "N"))
    same_sm = ssm == 0
    tco = 0.0
    for r in range(ns):
        _, dt = wall(lambda: o.solve_sparse(nz, rl[r][0], rl[r][1], "N")); tco += dt
        n_ = o.nzlhs
        same_sm = same_sm and len(outm[r][0]) == n_ and np.array_equal(outm[r][0], o.ilhs[:n_]) and np.array_equal(outm[r][1], o.lhs[o.ilhs[:n_]])
    rec["solve_sparse_multi"] = {"nrhs": nrhs, "gpu_ms_total": 1e3 * tsm, "gpu_ms_per_rhs": 1e3 * tsm / nrhs,
                                 "cpu_oracle_ms_per_rhs": 1e3 * tco / ns, "bit_identical_pattern_order_and_values": bool(same_sm),
                                 "note": "blu_solve_sparse_multi: one warp per right-hand side, all in flight; H2D/D2H included"}
    rec["solve_sparse"] = {"calls": nrhs, "nzrhs": nz, "avg_nzlhs": nzl / nrhs, "gpu_ms_per_call": 1e3 * tg / nrhs,
                           "cpu_oracle_ms_per_call": 1e3 * to / nrhs, "bit_identical_pattern_order_and_values": bool(same),
                           "bound": "latency (one warp; DFS on one lane)"}
    return [rec]


def c4full(args):
    """configs[3] at full size: 10^6 rows (ladder network, gen.config4_ladder), factorize + 1,000 Gilbert-Peierls
    solves with 0.1 %-dense right-hand sides."""
    cp, ri, v = gen.config4_ladder()
    m = len(cp) - 1
    g = BLU(m, len(v)); g.threads_per_basis = 1024
    o = Oracle(m, 40 * len(v)); o.set_param("check_file_diff", 0)
    sg, tg = wall(lambda: g.factorize(cp[:-1], cp[1:], ri, v))
    so, to = wall(lambda: o.factorize(cp[:-1], cp[1:], ri, v))
    byt = factor_bytes(g, m)
    rec = {"config": f"configs[3] FULL SIZE: circuit-like ladder network, m={m}, nnz={len(v)}", "m": m, "nnz": int(len(v)),
           "gpu": {"factorize_ms_wall": 1e3 * tg, "status": sg, "nrealloc": g.info("nrealloc"), "threads_per_basis": 1024,
                   "search_kcycles": g.info("t_phase3") / 1e3, "total_kcycles": g.info("t_phase11") / 1e3},
           "cpu_oracle_1thread": {"factorize_ms": 1e3 * to, "status": so},
           "stats": {k: g.info(k) for k in ("rank", "bump_size", "bump_nz", "l_nz", "u_nz", "factor_flops", "nsearch_pivot", "ngarbage", "nexpand")},
           "roofline": {"bound": "latency (one basis = one CTA)", "algorithmic_bytes": byt, "achieved_gbs": byt / tg / 1e9, "peak_gbs": HBM_PEAK, "frac": byt / tg / 1e9 / HBM_PEAK},
           "parity": {"status_equal": sg == so, "factors_bit_identical": same_factors(g, o),
                      "stats_equal": all(g.info(k) == o.info(k) for k in ("rank", "l_nz", "u_nz", "factor_flops", "nsearch_pivot"))}}
    nrhs, nz, chunk = args.nrhs or 1000, m // 1000, 50
    rl = [gen.sparse_rhs_np(6000 + r, m, nz) for r in range(nrhs)]
    tsm, same, checked, nzl = 0.0, True, 0, 0
    tco = 0.0
    for c0 in range(0, nrhs, chunk):
        (ssm, outm, statm), dt = wall(lambda: g.solve_sparse_multi(rl[c0:c0 + chunk], "N"))
        tsm += dt
        same = same and ssm == 0
        nzl += sum(len(x[0]) for x in outm)
        for r in (c0, c0 + chunk - 1)[:1 if c0 else 2]:      # bit-exact comparison of a sample against the oracle
            _, dt = wall(lambda: o.solve_sparse(nz, rl[r][0], rl[r][1], "N")); tco += dt
            n_ = o.nzlhs
            same = same and len(outm[r - c0][0]) == n_ and np.array_equal(outm[r - c0][0], o.ilhs[:n_]) and np.array_equal(outm[r - c0][1], o.lhs[o.ilhs[:n_]])
            checked += 1
    rec["solve_sparse_multi"] = {"nrhs": nrhs, "nzrhs": nz, "avg_nzlhs": nzl / nrhs, "gpu_ms_total": 1e3 * tsm, "gpu_ms_per_rhs": 1e3 * tsm / nrhs,
                                 "cpu_oracle_ms_per_rhs": 1e3 * tco / max(checked, 1), "compared_with_oracle": checked,
                                 "bit_identical_pattern_order_and_values": bool(same)}
    _, dt = wall(lambda: g.solve_sparse(nz, rl[0][0], rl[0][1], "N"))
    rec["solve_sparse_single_call_ms"] = 1e3 * dt
    return [rec]


def c5(args):
    m, bump, nupd = (100000, 2000, 500) if not args.quick else (20000, 1000, 60)
    cp, ri, v = gen.config3(m, bump, seed=7001)
    pool = gen.column_pool(7002, m, nupd)
    dense = bump * bump
    mem = (int(2.2 * (dense // 2 + 10 * m)) + 40 * m, int(1.2 * (dense // 2 + 10 * m)) + 40 * m, int(3.0 * dense + 40 * m))
    g, o = make_pair(m, len(v), 1024, mem, ofactor=max(mem))
    o.set_param("check_file_diff", 0)
    assert g.factorize(cp[:-1], cp[1:], ri, v) == o.factorize(cp[:-1], cp[1:], ri, v) == 0
    pcp, pri, pv = pool
    colptr, rowidx, vals = cp.copy(), [ri[cp[j]:cp[j + 1]] for j in range(m)], [v[cp[j]:cp[j + 1]] for j in range(m)]
    tg = to = 0.0
    tparts = [0.0, 0.0, 0.0]; oparts = [0.0, 0.0, 0.0]; nzl = 0
    same = True
    nft = 0
    done = 0
    for it in range(nupd):
        idx, val = pri[pcp[it]:pcp[it + 1]], pv[pcp[it]:pcp[it + 1]]
        s1, d1 = wall(lambda: g.solve_for_update(len(idx), idx, val, "N", 1)); tg += d1; tparts[0] += d1
        s2, d2 = wall(lambda: o.solve_for_update(len(idx), idx, val, "N", 1)); to += d2; oparts[0] += d2
        n_ = o.nzlhs; nzl += n_
        same = same and s1 == s2 == 0 and g.nzlhs == n_ and np.array_equal(g.ilhs[:n_], o.ilhs[:n_]) and np.array_equal(g.lhs, o.lhs)
        lhs = o.lhs
        j = int(np.argmax(np.abs(lhs)))                         # maxvolume.rs:120-131
        xtbl = lhs[j]
        jj = np.array([j])
        s1, d1 = wall(lambda: g.solve_for_update(1, jj, None, "T", 0)); tg += d1; tparts[1] += d1
        s2, d2 = wall(lambda: o.solve_for_update(1, jj, None, "T", 0)); to += d2; oparts[1] += d2
        nf0 = o.info("nforrest")
        s1, d1 = wall(lambda: g.update(xtbl)); tg += d1; tparts[2] += d1
        s2, d2 = wall(lambda: o.update(xtbl)); to += d2; oparts[2] += d2
        same = same and s1 == s2
        if s2 != 0:
            break
        nft += o.info("nforrest") > nf0
        rowidx[j], vals[j] = idx, val
        done += 1
    same = same and all(g.info(k) == o.info(k) for k in ("nforrest", "nupdate", "u_nz", "r_nz", "pivot_error", "max_eta"))
    b = gen.rhs(7003, m)
    _, xg = g.solve_dense(b, "N")
    _, xo = o.solve_dense(b, "N")
    same_dense = bool(np.array_equal(xg, xo))
    rec = {"config": f"configs[4]: replay on a {m}-row basis, {done} column replacements (solve_for_update N+T, update)",
           "m": m, "forrest_tomlin_updates": int(nft), "permutation_updates": int(done - nft),
           "gpu": {"us_per_replacement": 1e6 * tg / max(done, 1), "us_ftran_btran_update": [1e6 * t / max(done, 1) for t in tparts]},
           "cpu_oracle_1thread": {"us_per_replacement": 1e6 * to / max(done, 1), "us_ftran_btran_update": [1e6 * t / max(done, 1) for t in oparts]},
           "avg_nzlhs_ftran": nzl / max(done, 1),
           "parity": {"every_solution_and_counter_bit_identical": bool(same), "solve_dense_after_replay_bit_identical": same_dense},
           "bound": "latency (one warp per call + one H2D/D2H round trip per call)"}
    yield rec
    # refactorize the final basis
    lens = np.array([len(r) for r in rowidx])
    cp2 = np.concatenate([[0], np.cumsum(lens)])
    ri2, v2 = np.concatenate(rowidx), np.concatenate(vals)
    sg, trg = wall(lambda: g.factorize(cp2[:-1], cp2[1:], ri2, v2))
    so, tro = wall(lambda: o.factorize(cp2[:-1], cp2[1:], ri2, v2))
    yield {"config": f"configs[4]: refactorization of the {m}-row basis after {done} replacements", "m": m,
           "gpu": {"refactorize_ms": 1e3 * trg, "status": sg, "threads_per_basis": 1024},
           "cpu_oracle_1thread": {"refactorize_ms": 1e3 * tro, "status": so},
           "stats": {k: g.info(k) for k in ("rank", "bump_size", "bump_nz", "l_nz", "u_nz", "factor_flops")},
           "parity": {"refactorization_bit_identical": same_factors(g, o) if sg == so == 0 else None}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--nrhs", type=int, default=0, help="right-hand sides of the configs[3] solve series")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.jsonl"))
    ap.add_argument("which", nargs="*", default=["c1", "c3", "c4", "c5"])
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    fns = dict(c1=c1, c3=c3, c4=c4, c4full=c4full, c5=c5)
    with open(args.out, "a") as f:
        for w in args.which:
            t = time.time()
            for rec in fns[w](args):
                rec["wall_s_total"] = time.time() - t
                line = json.dumps(rec)
                print(line, flush=True)
                f.write(line + "\n")
                f.flush()


if __name__ == "__main__":
    main()
