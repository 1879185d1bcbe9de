"""The rule behind proj_resolve_kernel (orb_slam3_b200/csrc/match.cu): the reference's greedy, order-dependent
loop (ORBmatcher.cc:75-137 / :1740-1806) equals rounds in which
  * an unresolved point claims only candidates it could TAKE (distance <= TH_HIGH, not taken by a lower index),
  * a point finalises iff none of its free candidates is claimed by a lower unresolved point,
  * a finalising point ignores takes made by HIGHER-index points (they come later in the reference's loop),
  * takes are applied after the round.
Pure host logic on random, heavily overlapping candidate lists, the points of a round visited in random order (the
kernel visits them concurrently).  The kernel itself is checked against the oracle in tests/test_match_gpu.py."""
import random

TH_HIGH = 100


def _scan(cands, free, kind, ratio, lvl):
    bd, bl, bd2, bl2, bi = 256, -1, 256, -1, -1
    for c, d in cands:
        if not free(c):
            continue
        if d < bd:
            bd2, bd, bl2, bl, bi = bd, d, bl, lvl[c], c
        elif kind == 0 and d < bd2:
            bl2, bd2 = lvl[c], d
    acc = bd <= TH_HIGH
    if acc and kind == 0 and bl == bl2 and bd > ratio * bd2:
        acc = False
    return bi if acc else -1


def sequential(nq, nk, cands, has_obs, taken0, kind, ratio, lvl):
    taken, assign, nm = list(taken0), [-1] * nk, 0
    for j in range(nq):
        bi = _scan(cands[j], lambda c: not taken[c], kind, ratio, lvl)
        if bi >= 0:
            assign[bi] = j
            nm += 1
            if has_obs[j]:
                taken[bi] = True
    return assign, nm


def in_rounds(nq, nk, cands, has_obs, taken0, kind, ratio, lvl, rng):
    INF = 1 << 30
    takenby = [-1 if t else INF for t in taken0]
    assign, nm = [-1] * nk, 0
    state = [1 if cands[j] else 0 for j in range(nq)]
    acc_kp = [-1] * nq
    rounds = 0
    while True:
        rounds += 1
        for j in range(nq):
            if state[j] == 2:
                takenby[acc_kp[j]] = j
                state[j] = 0
        minidx = [INF] * nk
        for j in range(nq):
            if state[j] == 1:
                for c, d in cands[j]:
                    if d <= TH_HIGH and takenby[c] > j:
                        minidx[c] = min(minidx[c], j)
        order = [j for j in range(nq) if state[j] == 1]
        rng.shuffle(order)
        unresolved, new_state = 0, {}
        for j in order:
            if any(takenby[c] > j and minidx[c] < j for c, _ in cands[j]):
                unresolved += 1
                continue
            bi = _scan(cands[j], lambda c: takenby[c] >= j, kind, ratio, lvl)
            st = 0
            if bi >= 0:
                assign[bi] = j
                acc_kp[j] = bi
                nm += 1
                if has_obs[j]:
                    st = 2
            new_state[j] = st
        for j, st in new_state.items():
            state[j] = st
        if unresolved == 0:
            return assign, nm, rounds


def test_rounds_equal_the_sequential_loop():
    rng = random.Random(1)
    for trial in range(1500):
        nk, nq, kind = rng.randint(5, 60), rng.randint(1, 80), rng.randint(0, 1)
        lvl = [rng.randint(0, 3) for _ in range(nk)]
        low = rng.choice([0.05, 0.3, 0.8])
        cands = []
        for _ in range(nq):
            cs = rng.sample(range(nk), rng.randint(0, min(nk, 12)))
            cands.append([(c, rng.randint(20, 100) if rng.random() < low else rng.randint(101, 180)) for c in cs])
        has_obs = [rng.random() < 0.7 for _ in range(nq)]
        taken0 = [rng.random() < 0.15 for _ in range(nk)]
        a, n = sequential(nq, nk, cands, has_obs, taken0, kind, 0.8, lvl)
        b, m, _ = in_rounds(nq, nk, cands, has_obs, taken0, kind, 0.8, lvl, rng)
        assert a == b and n == m, trial


def test_far_candidates_do_not_serialise_the_points():
    """Points whose shared candidates are all beyond TH_HIGH finish in ONE round (the previous rule needed nq)."""
    nq, nk = 50, 60
    cands = [[(j, 40)] + [(c, 130) for c in range(50, 60)] for j in range(nq)]
    a, n, rounds = in_rounds(nq, nk, cands, [True] * nq, [False] * nk, 0, 0.8, [0] * nk, random.Random(0))
    assert rounds == 1 and n == nq and a[:50] == list(range(50))


def test_second_best_is_the_next_distance_position_key():
    """The kernel finds best / second best as the two smallest (distance, list position) keys with a warp
    reduction; the reference's scan (strict `<` updates, ORBmatcher.cc:99-116) ends in the same state, ties
    included."""
    rng = random.Random(5)
    for _ in range(20000):
        m = rng.randint(0, 12)
        items = [(rng.randint(0, 7), rng.choice([30, 30, 31, 40, 40, 55, 90, 120])) for _ in range(m)]  # (level, dist)
        bd, bl, bd2, bl2, bi = 256, -1, 256, -1, -1
        for pos, (lv, d) in enumerate(items):
            if d < bd:
                bd2, bd, bl2, bl, bi = bd, d, bl, lv, pos
            elif d < bd2:
                bl2, bd2 = lv, d
        keys = sorted((d << 16) | pos for pos, (_, d) in enumerate(items))
        if not keys:
            assert bi == -1
            continue
        assert (keys[0] >> 16, keys[0] & 0xffff) == (bd, bi)
        if len(keys) > 1:
            assert keys[1] >> 16 == bd2 and items[keys[1] & 0xffff][0] == bl2
        else:
            assert bd2 == 256 and bl2 == -1


def test_dropping_irrelevant_candidates_keeps_the_result():
    """proj_candidates_kernel drops candidates beyond TH_HIGH that cannot act as second best either
    (ratio * dist >= TH_HIGH); SearchByProjection(Cur, Last) keeps only those within TH_HIGH."""
    import numpy as np
    rng = random.Random(9)
    for trial in range(1500):
        nk, nq, kind = rng.randint(5, 40), rng.randint(1, 50), rng.randint(0, 1)
        ratio = rng.choice([0.6, 0.7, 0.8, 0.9])
        lvl = [rng.randint(0, 2) for _ in range(nk)]
        cands = []
        for _ in range(nq):
            cs = rng.sample(range(nk), rng.randint(0, min(nk, 10)))
            cands.append([(c, rng.randint(60, 180)) for c in cs])
        has_obs = [rng.random() < 0.7 for _ in range(nq)]
        taken0 = [rng.random() < 0.15 for _ in range(nk)]

        def relevant(d):
            return d <= TH_HIGH or (kind == 0 and float(np.float32(ratio) * np.float32(d)) < TH_HIGH)
        kept = [[(c, d) for c, d in lst if relevant(d)] for lst in cands]
        r32 = float(np.float32(ratio))
        assert sequential(nq, nk, cands, has_obs, taken0, kind, r32, lvl) == \
            sequential(nq, nk, kept, has_obs, taken0, kind, r32, lvl), trial
