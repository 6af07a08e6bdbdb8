"""Host logic: layout + planner through the C ABI, against the oracle restatement, the
reference scheduler compiled in place, and the schedule table of SURVEY.md section 8."""
import random

import pytest

# (w, h) -> [(pipeline, inputLevel, levelCount, workgroups)]  -- SURVEY.md section 8 table
SURVEY_SCHEDULES = {
    (4096, 4096): [(1, 0, 6, 4096), (1, 6, 6, 1)],
    (16384, 16384): [(1, 0, 6, 65536), (1, 6, 6, 16), (1, 12, 2, 1)],
    (4095, 4095): [(0, 0, 2, 16384), (0, 2, 2, 1024), (0, 4, 2, 64), (0, 6, 2, 4), (0, 8, 2, 1), (0, 10, 1, 1)],
    (2047, 2047): [(0, 0, 2, 4096), (0, 2, 2, 256), (0, 4, 2, 16), (0, 6, 2, 1), (0, 8, 2, 1)],
    (1920, 1080): [(1, 0, 3, 2025), (0, 3, 2, 40), (0, 5, 2, 2), (0, 7, 2, 1), (0, 9, 1, 1)],
    (2560, 1440): [(1, 0, 5, 3600), (0, 5, 2, 6), (0, 7, 2, 1), (0, 9, 2, 1)],
    (1080, 4096): [(1, 0, 3, 4320), (0, 3, 2, 80), (0, 5, 2, 4), (1, 7, 3, 1), (0, 10, 2, 1)],
    (2052, 2052): [(1, 0, 2, 4113), (0, 2, 2, 256), (1, 4, 6, 4), (0, 10, 1, 1)],
    (4094, 4094): [(0, 0, 2, 16384), (0, 2, 2, 1024), (0, 4, 2, 64), (0, 6, 2, 4), (0, 8, 2, 1), (0, 10, 1, 1)],
}


def abi_plan(nv, w, h, levels=0, **kw):
    return [(s["pipeline"], s["inputLevel"], s["levelCount"], s["workgroups"], s["pushConstant"], s["bindPipeline"],
             s["barrierAfter"]) for s in nv.get_plan(w, h, levels, **kw)]


@pytest.mark.parametrize("size", sorted(SURVEY_SCHEDULES))
def test_plan_matches_survey_table(nv, size):
    got = [s[:4] for s in abi_plan(nv, *size)]
    assert got == SURVEY_SCHEDULES[size]


def test_plan_matches_oracle_and_reference(nv, oracle, ref):
    rnd = random.Random(7)
    sizes = list(SURVEY_SCHEDULES) + [(1, 7), (7, 1), (2, 2), (3, 3), (5, 64), (64, 5), (1, 2), (2, 1), (65535, 3)]
    sizes += [(rnd.randint(1, 6000), rnd.randint(1, 6000)) for _ in range(400)]
    sizes += [(4 * rnd.randint(1, 1500), 4 * rnd.randint(1, 1500)) for _ in range(200)]
    for w, h in sizes:
        for have_fast in (1, 0):
            a = abi_plan(nv, w, h, flags=0 if have_fast else nv.FLAG_FORCE_GENERAL)
            o = [s.key() for s in oracle.plan(w, h, 0, have_fast)]
            r = [s.key() for s in ref.plan(w, h, 0, have_fast)]
            assert a == o == r, (w, h, have_fast)


def test_plan_partial_level_counts(nv, oracle, ref):
    for w, h in [(4096, 4096), (1920, 1080), (333, 77)]:
        for levels in range(2, nv.level_count(w, h) + 1):
            a = abi_plan(nv, w, h, levels)
            assert a == [s.key() for s in oracle.plan(w, h, levels)] == [s.key() for s in ref.plan(w, h, levels)]
            assert sum(s[2] for s in a) == levels - 1


@pytest.mark.parametrize("div,max_levels", [(2, 6), (2, 5), (2, 3), (8, 3)])
def test_plan_template_variants(nv, oracle, ref, div, max_levels):
    """<DivisibilityRequirement, MaxLevels> variants (levels_1_6, levels_1_5, levels_1_3, levels_3_3)."""
    rnd = random.Random(div * 10 + max_levels)
    sizes = [(4096, 4096), (2048, 2048), (1920, 1080), (4094, 4094), (2052, 2052)]
    sizes += [(2 * rnd.randint(1, 3000), 2 * rnd.randint(1, 3000)) for _ in range(100)]
    for w, h in sizes:
        a = abi_plan(nv, w, h, fast_divisibility=div, fast_max_levels=max_levels)
        o = [s.key() for s in oracle.plan(w, h, 0, 1, div, max_levels)]
        r = [s.key() for s in ref.plan_variant(w, h, 0, div, max_levels)]
        assert a == o == r, (w, h)


def test_plan_invariants(nv):
    rnd = random.Random(3)
    for _ in range(300):
        w, h = rnd.randint(1, 70000), rnd.randint(1, 70000)
        n = nv.level_count(w, h)
        if n == 1:
            continue
        plan = nv.get_plan(w, h)
        lvl = 0
        for i, s in enumerate(plan):
            assert s["inputLevel"] == lvl and s["levelCount"] >= 1
            assert s["pushConstant"] == (s["inputLevel"] << 5 | s["levelCount"])
            assert s["srcWidth"] == max(1, w >> lvl) and s["srcHeight"] == max(1, h >> lvl)
            if s["pipeline"] == 1:
                assert s["levelCount"] <= 6
                assert s["srcWidth"] % 4 == 0 and s["srcHeight"] % 4 == 0
                assert s["srcWidth"] % (1 << s["levelCount"]) == 0 and s["srcHeight"] % (1 << s["levelCount"]) == 0
            else:
                assert s["levelCount"] <= 2
            assert s["barrierAfter"] == (i != len(plan) - 1)
            lvl += s["levelCount"]
        assert lvl == n - 1


def test_layout(nv, oracle, ref):
    rnd = random.Random(11)
    for w, h in [(1, 1), (16384, 16384), (4095, 4095), (1, 9), (9, 1)] + [(rnd.randint(1, 9000), rnd.randint(1, 9000))
                                                                          for _ in range(100)]:
        n = nv.level_count(w, h)
        lay = ref.layout(w, h)
        assert n == oracle.level_count(w, h) == len(lay)
        for i in range(n):
            assert (nv.level_offset_texels(w, h, i), *nv.level_extent(w, h, i)) == lay[i]
        assert nv.chain_texels(w, h) == oracle.chain_texels(w, h) == lay[-1][0] + 1
        assert nv.chain_bytes(w, h, 0, nv.FORMAT_RGBA32F) == 16 * nv.chain_texels(w, h)


def test_headline_sizes(nv):
    assert nv.level_count(16384, 16384) == 15
    assert nv.chain_bytes(16384, 16384) == 1431655764
    assert nv.chain_bytes(4096, 4096) == 89478484
    assert nv.chain_bytes(4095, 4095) == 89413008
    assert nv.chain_bytes(2047, 2047) == 22336908
