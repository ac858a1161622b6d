"""Expected result rows of bench.py's full-size configurations, from tests/golden/bench_golden.json.

The golden file holds exact integers computed by an independent numpy evaluation of the SQL (tools/make_bench_golden.py:
no executor, no hashing, no joins - the synthetic tables are counter based, so a lineitem row's order, customer, supplier
and nation follow from its row index).  This module turns them into the text the executor's rows print as - `numeric`
display scales and `avg` rounding restated here from utils/adt/numeric.c (numeric_sum :6091, numeric_avg :6056,
select_div_scale :9194-9254, div_var round-half-away) - so that bench.py can compare every run's rows with them
(`result_check`) at every N.  Nothing here touches oracle/ or the native libraries.
"""
import json
import os

from . import ssb, tpch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "tests", "golden", "bench_golden.json")
SOURCE = "tests/golden/bench_golden.json (independent numpy evaluation of the SQL, tools/make_bench_golden.py)"
_cache = {}


def load(path=None):
    path = path or os.environ.get("CBGPU_BENCH_GOLDEN") or PATH
    if path not in _cache:
        _cache[path] = json.load(open(path)) if os.path.exists(path) else None
    return _cache[path]


def scaled_text(v, scale):
    """a scaled integer as numeric_out prints it at display scale `scale`"""
    v = int(v)
    sign = "-" if v < 0 else ""
    v = abs(v)
    if scale == 0:
        return sign + str(v)
    s = str(v).rjust(scale + 1, "0")
    return sign + s[:-scale] + "." + s[-scale:]


def _weight_first(v, scale):
    """(weight, first base-10000 digit) of |v| / 10^scale as a NumericVar holds it (numeric.c: digits base NBASE = 10000,
    weight = position of the first digit)"""
    v = abs(int(v))
    if v == 0:
        return 0, 0
    ip = v // 10 ** scale
    if ip > 0:
        s = str(ip)
        return (len(s) - 1) // 4, int(s[:(len(s) - 1) % 4 + 1])
    frac = str(v % 10 ** scale).rjust(scale, "0")
    frac += "0" * (-len(frac) % 4)
    for g in range(0, len(frac), 4):
        if int(frac[g:g + 4]) != 0:
            return -1 - g // 4, int(frac[g:g + 4])
    return 0, 0


def avg_text(total, scale, n):
    """numeric_avg: numeric_div(sum, N) with select_div_scale's result scale, rounded half away from zero"""
    w1, f1 = _weight_first(total, scale)
    w2, f2 = _weight_first(n, 0)
    qweight = w1 - w2 - (1 if f1 <= f2 else 0)
    rscale = min(max(16 - 4 * qweight, scale, 0), 1000)
    num = abs(int(total)) * 10 ** rscale
    den = int(n) * 10 ** scale
    q, r = divmod(num, den)
    if 2 * r >= den:
        q += 1
    return scaled_text(-q if int(total) < 0 else q, rscale)


def q1_rows(gold_q1, nranks):
    """format_q1-style rows of Q1 over the union of shards 0 .. nranks-1 (bench.py's weak-scaling distribution: rank r
    holds generator rows [r * n, (r + 1) * n))"""
    acc = {}
    for sh in gold_q1["shards"][:nranks]:
        for g, st in sh.items():
            a = acc.setdefault(g, [0] * 6)
            for i in range(6):
                a[i] += int(st[i])
    rows = []
    for g in sorted(acc):
        n, qty, ext, dp, ch, disc = acc[g]
        rows.append([g[0], g[1], scaled_text(qty, 2), scaled_text(ext, 2), scaled_text(dp, 4), scaled_text(ch, 6),
                     avg_text(qty, 2, n), avg_text(ext, 2, n), avg_text(disc, 2, n), str(n)])
    return rows


def q3_rows(gold_q3):
    return [[r[0], scaled_text(r[1], 4), tpch.days_to_text(r[2]), str(r[3])] for r in gold_q3["rows"]]


def q5_rows(gold_q5):
    return [[r[0], scaled_text(r[1], 4)] for r in gold_q5["rows"]]


def ssb_rows(gold_ssb, q):
    return [[int(x) for x in r[:-1]] + [str(r[-1])] for r in gold_ssb[q]]


def check(name, got, want):
    """'ok', or a short description of the first difference"""
    if got == want:
        return "ok"
    if len(got) != len(want):
        return "MISMATCH %s: %d rows, expected %d" % (name, len(got), len(want))
    for i, (a, b) in enumerate(zip(got, want)):
        if a != b:
            return "MISMATCH %s row %d: got %r, expected %r" % (name, i, a, b)
    return "MISMATCH %s" % name


def key(kind, sf):
    sf = int(sf) if float(sf).is_integer() else sf
    return "%s_sf%s" % (kind, sf)


__all__ = ["load", "q1_rows", "q3_rows", "q5_rows", "ssb_rows", "check", "key", "SOURCE", "ssb"]
