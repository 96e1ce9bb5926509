"""CPU tests: the oracle restatement (oracle/rt_oracle.cpp) against the golden vectors produced by
the UNMODIFIED reference (tests/golden/make_golden.py), and against the survey's published hashes."""
import gzip
import json
import os

import numpy as np
import pytest

import raytrace_b200 as R
from parity_util import oracle_render

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))["cases"]


def case_id(c):
    return f"{c['scene']}-{c['w']}x{c['h']}-l{c['level']}-n{c['n']}-t{c['type']}"


def test_survey_hashes_are_in_the_golden_set():
    # hashes of the reference output recorded by the survey
    want = {("c1", 1088, 576, 1): "e674753ec4f3b630", ("c1", 1088, 576, 5): "a02c8e981712e3d3", ("c1", 1920, 1080, 5): "2bda0bcb0aaf620e"}
    got = {(c["scene"], c["w"], c["h"], c["level"]): c["hash"] for c in GOLDEN if c["type"] == 0x80}
    for k, v in want.items():
        assert got[k] == v


@pytest.mark.parametrize("case", GOLDEN, ids=case_id)
def test_oracle_matches_reference_golden(case):
    sc = R.Scene(case["scene"], case["w"], case["h"], case["n"], case["parts"])
    img, ids, cnt = oracle_render(sc, case["level"], case["type"], want_ids=case["type"] == 0x80)
    assert R.fnv1a64(img) == case["hash"]
    rays = case["rays"]
    if case["type"] == 7:   # RTflec shoots untyped rays: the proxy files them all under "shadow"
        assert cnt.primary == rays["primary"] and cnt.shadow + cnt.reflect == rays["shadow"] + rays["reflect"]
    else:
        assert (cnt.primary, cnt.shadow, cnt.reflect, cnt.refract) == (rays["primary"], rays["shadow"], rays["reflect"], rays["refract"])
    if "ids_hash" in case:
        assert R.fnv1a64(ids) == case["ids_hash"]
        assert int((ids["object"] >= 0).sum()) == case["hit_pixels"]
    if "frame" in case:
        raw = gzip.open(os.path.join(HERE, "golden", case["frame"])).read()
        ref = np.frombuffer(raw, np.uint8).reshape(case["h"], case["w"], 3)
        assert np.array_equal(ref, img)


def test_unrendered_margin_stays_127():
    # RayTracer.cpp:13,620: only floor(W/64)*64 x floor(H/64)*64 pixels are written
    sc = R.Scene("c1", 320, 200)
    img, _, _ = oracle_render(sc, 1, want_ids=False)
    assert (img[192:] == 127).all() and not (img[:192] == 127).all()
    sc = R.Scene("c1", 100, 50)
    img, _, cnt = oracle_render(sc, 1, want_ids=False)
    assert (img == 127).all() and cnt.primary == 0


def test_oracle_row_shards_tile_the_frame():
    sc = R.Scene("t_mixed", 320, 256)
    full, _, cfull = oracle_render(sc, 3, want_ids=False)
    acc = np.full_like(full, 127)
    total = 0
    for r in range(3):
        part, _, c = oracle_render(sc, 3, want_ids=False, rank=r, world=3)
        rows = [y for y in range(256) if (y // 64) % 3 == r]
        acc[rows] = part[rows]
        other = [y for y in range(256) if (y // 64) % 3 != r]
        assert (part[other] == 127).all()
        total += c.primary + c.shadow + c.reflect + c.refract
    assert np.array_equal(acc, full)
    assert total == cfull.primary + cfull.shadow + cfull.reflect + cfull.refract


def test_oracle_serpentine_shards_tile_the_frame():
    # RT_FLAG_SERPENTINE: odd groups of `world` tiles go to the ranks in reverse order
    from raytrace_b200.distributed import bands_of
    sc = R.Scene("t_mixed", 320, 256)
    full, _, cfull = oracle_render(sc, 2, want_ids=False)
    acc = np.full_like(full, 127)
    seen = []
    for r in range(3):
        part, _, c = oracle_render(sc, 2, want_ids=False, rank=r, world=3, tile_rows=16, flags=R.RT_FLAG_SERPENTINE)
        tiles = bands_of(r, 3, 256, 16, serpentine=True)
        seen += tiles
        rows = [y for y in range(256) if y // 16 in tiles]
        other = [y for y in range(256) if y // 16 not in tiles]
        assert (part[other] == 127).all()
        acc[rows] = part[rows]
    assert sorted(seen) == list(range(16))
    assert bands_of(0, 3, 256, 16, serpentine=True) == [0, 5, 6, 11, 12] and bands_of(2, 3, 256, 16, serpentine=True) == [2, 3, 8, 9, 14, 15]
    assert np.array_equal(acc, full)


def test_tile_window_of_the_oracle_partitions_a_shard():
    # rt_render_params::tile_first / tile_count (the band loop of supersampled frames): the windows of a shard tile it
    import ctypes as C

    import numpy as np

    import raytrace_b200 as R
    from parity_util import oracle_lib
    w, h, level = 192, 192, 2
    sc = R.Scene("t_mixed", w, h)
    lib = oracle_lib()

    def render(rank, world, flags, tr, first, count):
        out = np.empty((h, w, 3), np.uint8)
        cnt = R.Counters()
        p = R.RenderParams(R.MY_MODEL_RAYTRACE, level, rank, world, flags, tr, first, count)
        assert lib.rto_render(sc.flatten(), C.byref(p), out.ctypes.data, None, C.byref(cnt), 4) == 0
        return out, cnt.primary + cnt.shadow + cnt.reflect + cnt.refract
    for rank, world, flags, tr in ((0, 1, 0, 64), (1, 2, R.RT_FLAG_SERPENTINE, 8), (2, 3, 0, 16)):
        whole, rays = render(rank, world, flags, tr, 0, 0)
        acc, total, first = np.full_like(whole, 127), 0, 0
        for count in (1, 2, 0):
            part, r = render(rank, world, flags, tr, first, count)
            sel = (part != 127).any(axis=(1, 2))
            assert (acc[sel] == 127).all()          # windows do not overlap
            acc[sel] = part[sel]
            total += r
            first += count
        assert np.array_equal(acc, whole) and total == rays
