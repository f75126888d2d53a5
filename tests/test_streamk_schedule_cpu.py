"""The stream-K schedule of the FP64 GEMM (csrc/gemm_f64.cu: sk_schedule on the host, SkCursor / sk_owner on the device).
The host half is called through the C ABI (adb_gemm_sk_plan, no GPU needed); the device half is restated here in Python,
integer for integer, and the two are checked against the invariants the kernel relies on:
  * the shares of the K-sharing CTAs tile the units of the shared region exactly once, in order;
  * a CTA has at most one segment that starts inside a tile and at most one that starts a tile but ends inside it
    (the two workspace slots 2c and 2c + 1);
  * for every split tile, the CTAs sk_owner() enumerates (first .. last) are exactly the CTAs that hold a part of it, in k order,
    and the arrival count the last arriver waits for equals the number of parts."""
import ctypes as C
import random

import pytest

from autodiff_b200 import _lib


def plan(M, N, K, batch, bm, bn, slots, min_units):
    out = (C.c_longlong * 6)()
    _lib.check(_lib.lib().adb_gemm_sk_plan(M, N, K, batch, bm, bn, slots, min_units, out))
    return dict(total_kt=out[0], tiles=out[1], dp_tiles=out[2], sk_units=out[3], sk_ctas=out[4], grid=out[5])


def unit_begin(p, c):                      # sk_unit_begin
    return c * p["sk_units"] // p["sk_ctas"]


def owner(p, u):                           # sk_owner
    c = min(u * p["sk_ctas"] // p["sk_units"], p["sk_ctas"] - 1)
    while unit_begin(p, c) > u:
        c -= 1
    while unit_begin(p, c + 1) <= u:
        c += 1
    return c


def segments(p, c):                        # SkCursor / sk_advance
    u, ue, kt, out = unit_begin(p, c), unit_begin(p, c + 1), p["total_kt"], []
    while u < ue:
        t = u // kt
        lo = u - t * kt
        hi = min(kt, lo + (ue - u))
        out.append((t, lo, hi))
        u += hi - lo
    return out


CASES = [(8192, 8192, 8192, 1, 128, 128, 148, 8), (1024, 4096, 4097, 1, 128, 128, 148, 8), (512, 1024, 1024, 1, 128, 128, 148, 8),
         (512, 1024, 1024, 1, 64, 64, 296, 8), (1024, 1024, 1024, 64, 128, 128, 148, 8), (130, 130, 1, 1, 128, 128, 148, 8),
         (257, 129, 50, 3, 64, 64, 296, 8), (4097, 4096, 8192, 1, 128, 128, 148, 8), (300, 200, 8000, 1, 128, 128, 148, 8),
         (128 * 148, 128, 16, 1, 128, 128, 148, 8), (128 * 149, 128, 16 * 7, 1, 128, 128, 148, 8), (640, 640, 4097, 1, 128, 128, 148, 1)]


@pytest.mark.parametrize("case", CASES + list(range(40)))
def test_stream_k_schedule_invariants(case):
    if isinstance(case, int):                                                        # a seeded random problem
        rnd = random.Random(1000 + case)
        bm = bn = rnd.choice([64, 128])
        case = (rnd.randint(1, 6000), rnd.randint(1, 6000), rnd.randint(1, 9000), rnd.choice([1, 1, 1, 2, 5]), bm, bn,
                148 * (2 if bm == 64 else 1), rnd.choice([1, 8, 8, 16, 48]))
    M, N, K, batch, bm, bn, slots, mu = case
    p = plan(*case)
    kt = p["total_kt"]
    assert kt == (K + 15) // 16 and p["tiles"] == ((M + bm - 1) // bm) * ((N + bn - 1) // bn) * batch
    assert p["dp_tiles"] % slots == 0 and 0 <= p["tiles"] - p["dp_tiles"] < slots
    rem = p["tiles"] - p["dp_tiles"]
    assert p["sk_units"] == rem * kt and p["grid"] == p["dp_tiles"] + p["sk_ctas"]
    if rem == 0:
        assert p["sk_ctas"] == 0
        return
    assert rem <= p["sk_ctas"] <= slots
    if p["sk_ctas"] > rem:
        assert p["sk_units"] // p["sk_ctas"] >= mu                                   # no share shorter than min_units
    covered, parts = 0, {}
    for c in range(p["sk_ctas"]):
        segs = segments(p, c)
        assert segs, "an empty share: the reducer would miscount the parts of a tile"
        assert unit_begin(p, c) == covered                                           # shares are contiguous and in order
        mid_start = [s for s in segs if s[1] > 0]
        head_only = [s for s in segs if s[1] == 0 and s[2] < kt]
        assert len(mid_start) <= 1 and len(head_only) <= 1                           # the two workspace slots of a CTA
        for t, lo, hi in segs:
            assert 0 <= lo < hi <= kt and 0 <= t < rem
            covered += hi - lo
            if lo > 0 or hi < kt:
                parts.setdefault(t, []).append((lo, hi, c))
    assert covered == p["sk_units"]
    for t, ps in parts.items():
        first, last = owner(p, t * kt), owner(p, t * kt + kt - 1)
        assert [c for _, _, c in ps] == list(range(first, last + 1))                 # the reducer's enumeration, in k order
        assert ps[0][0] == 0 and ps[-1][1] == kt and all(a[1] == b[0] for a, b in zip(ps, ps[1:]))
        assert len(ps) == last - first + 1 >= 2                                      # arrival count the last arriver sees


def test_stream_k_plan_rejects_oversized_and_bad_arguments():
    with pytest.raises(_lib.BackendError):
        plan(0, 10, 10, 1, 128, 128, 148, 8)
    with pytest.raises(_lib.BackendError):
        plan(2 ** 31 - 1, 2 ** 31 - 1, 64, 1, 128, 128, 148, 8)                      # more than 2^31 tiles
