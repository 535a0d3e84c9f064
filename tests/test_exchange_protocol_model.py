"""Exhaustive interleaving check of the buffer-reuse protocol of the sharded fused loop (CPU, no GPU).

This is a MODEL of the stream/event/flag dependencies in csrc/spx_fused.cu (`fused_run`,
`shard_price_kernel`, `launch_fused_update`), not a run of the kernels.  Per rank r and pass q:

    Pb(r,q)    pricing kernel starts on the side stream     needs Pe(r,q-1) [stream order] and, with
                                                            look-ahead, Ue(r,q-2) [ev_upd wait]
    Pw(r,q,t)  rank r stores its candidate-column planes    needs Pb(r,q)
               (slot q % PLANE_SLOTS) and keys (slot q % 2)
               into rank t's XBOX over NVLink
    Pe(r,q)    pricing kernel ends                          needs Pw(s,q,r) from every rank s [flag waits]
                                                            and its own Pw(r,q,*)
    Ub/Ue(r,q) update kernel of pass q on the main stream   needs Pe(r,q) [ev_priced] and Ue(r,q-1)

Without look-ahead everything is one stream: Pb(r,q) also needs Ue(r,q-1).  The PERSISTENT pricing engine (one pricing
kernel per call) has the same graph with device flags in place of the events plus the edge "the first update kernel waits
until the engine is resident" (last two tests; the second one reproduces the deadlock met on hardware when CUDA's lazy
kernel loading held the first update launch back until the engine had ended).  The model is COARSER than
the kernels (one exchange per pass instead of one per level), i.e. it admits more interleavings, so
"no violation in the model" carries over.  A read is a violation if, when the reading kernel ends, a
location it read holds data of another pass (stores into one location are issued by one rank in pass
order, so the location's content is the newest completed store).

The check explores EVERY reachable set of completed events.  It must pass with the 3 plane slots the
code uses and must find the overwrite with 2 (the race that motivated the third slot).
"""

import pytest


def explore(ranks, passes, plane_slots, key_slots, lookahead, extra_deps=None):
    """extra_deps(event) -> further events it waits for (used to model the persistent engine's extra edges)."""
    R, Q = ranks, passes
    events = []
    for r in range(R):
        for q in range(1, Q + 1):
            events.append(("Pb", r, q, -1))
            events.extend(("Pw", r, q, t) for t in range(R))
            events.append(("Pe", r, q, -1))
            events.append(("Ub", r, q, -1))
            events.append(("Ue", r, q, -1))
    index = {e: i for i, e in enumerate(events)}

    def deps(e):
        kind, r, q, t = e
        d = []
        if kind == "Pb":
            if q > 1:
                d.append(("Pe", r, q - 1, -1))
            if lookahead:
                if q > 2:
                    d.append(("Ue", r, q - 2, -1))
            elif q > 1:
                d.append(("Ue", r, q - 1, -1))
        elif kind == "Pw":
            d.append(("Pb", r, q, -1))
        elif kind == "Pe":
            d.extend(("Pw", s, q, r) for s in range(R))
            d.extend(("Pw", r, q, t2) for t2 in range(R))
        elif kind == "Ub":
            d.append(("Pe", r, q, -1))
            if q > 1:
                d.append(("Ue", r, q - 1, -1))
        elif kind == "Ue":
            d.append(("Ub", r, q, -1))
        if extra_deps is not None:
            d.extend(extra_deps(e))
        return [index[x] for x in d]

    dep_mask = [sum(1 << i for i in set(deps(e))) for e in events]

    def newest_store(done, src, dst, slot, slots):
        """pass number of the newest completed store of `src` into `dst`'s slot"""
        best = 0
        for q in range(1, Q + 1):
            if q % slots == slot and done >> index[("Pw", src, q, dst)] & 1:
                best = q
        return best

    def violation(done, e):
        kind, r, q, _ = e
        if kind == "Ue":                                   # the update read the planes of pass q
            for s in range(R):
                if newest_store(done, s, r, q % plane_slots, plane_slots) != q:
                    return f"update {q} on rank {r}: planes from rank {s} overwritten"
        if kind == "Pe":                                   # pricing read keys of pass q and, replaying, planes q-1
            for s in range(R):
                if newest_store(done, s, r, q % key_slots, key_slots) != q:
                    return f"pricing {q} on rank {r}: keys from rank {s} overwritten"
                if lookahead and q > 1 and newest_store(done, s, r, (q - 1) % plane_slots, plane_slots) != q - 1:
                    return f"pricing {q} on rank {r}: previous planes from rank {s} overwritten"
        return None

    seen = {0}
    frontier = [0]
    full = (1 << len(events)) - 1
    reached_end = False
    while frontier:
        nxt = []
        for done in frontier:
            if done == full:
                reached_end = True
            for i, e in enumerate(events):
                if done >> i & 1 or (dep_mask[i] & ~done):
                    continue
                after = done | 1 << i
                bad = violation(after, e)
                if bad:
                    return bad, len(seen)
                if after not in seen:
                    seen.add(after)
                    nxt.append(after)
        frontier = nxt
    assert reached_end, "the model deadlocked"
    return None, len(seen)


@pytest.mark.parametrize("ranks,passes", [(2, 5), (3, 4)])
def test_three_plane_slots_are_race_free_with_lookahead(ranks, passes):
    bad, states = explore(ranks, passes, plane_slots=3, key_slots=2, lookahead=True)
    assert bad is None, bad
    assert states > 1000                                    # the exploration really branched


def test_two_plane_slots_race_with_lookahead():
    bad, _ = explore(2, 4, plane_slots=2, key_slots=2, lookahead=True)
    assert bad is not None and "planes" in bad


@pytest.mark.parametrize("plane_slots", [2, 3])
def test_serial_passes_are_race_free(plane_slots):
    bad, _ = explore(2, 4, plane_slots=plane_slots, key_slots=2, lookahead=False)
    assert bad is None, bad


def test_one_key_slot_would_race():
    bad, _ = explore(2, 3, plane_slots=3, key_slots=1, lookahead=True)
    assert bad is not None and "keys" in bad


# ---- the persistent pricing engine (fused_run engine 2): ONE pricing kernel per call walks Pb/Pe of every pass; the same
# two dependencies travel through device flags (Ub(q) <- Pe(q): cs->plan_ready; Pb(q) <- Ue(q-2): cs->upd_done) instead of
# events, plus one edge: the first update kernel may not start before the engine is resident (wait_engine_kernel).
def _engine_edges(e):
    kind, r, q, _ = e
    return [("Pb", r, 1, -1)] if kind == "Ub" and q == 1 else []


@pytest.mark.parametrize("ranks,passes", [(2, 5), (3, 4)])
def test_persistent_engine_has_the_same_dependency_graph(ranks, passes):
    bad, states = explore(ranks, passes, plane_slots=3, key_slots=2, lookahead=True, extra_deps=_engine_edges)
    assert bad is None, bad
    assert states > 1000


def test_update_launches_held_back_until_the_engine_ends_deadlock():
    """What happened on 2 GPUs in fresh processes (profiles/r2/r2q_r2r_persistent_engine.md): the host blocked inside the
    first update launch (CUDA was loading that kernel lazily, which synchronises with the resident engine), i.e. Ub(1) could
    not happen before the engine's last pass had ended — while the engine's third pass waits for Ue(1)."""
    Q = 4

    def blocked(e):
        kind, r, q, _ = e
        return _engine_edges(e) + ([("Pe", r, Q, -1)] if kind == "Ub" and q == 1 else [])

    with pytest.raises(AssertionError, match="deadlocked"):
        explore(2, Q, plane_slots=3, key_slots=2, lookahead=True, extra_deps=blocked)
    # two passes per call could never show it: nothing in them waits for an update
    bad, _ = explore(2, 2, plane_slots=3, key_slots=2, lookahead=True,
                     extra_deps=lambda e: _engine_edges(e) + ([("Pe", e[1], 2, -1)] if e[0] == "Ub" and e[2] == 1 else []))
    assert bad is None
