"""The schedule of the pipelined device planner (gbp_pipeline.cuh), restated in Python on top of the oracle's PRIMITIVES and
compared with the oracle's sequential planner (orc_plan_ex, itself pinned to the unmodified reference's runRRTConnect loop by
tests/test_oracle_vs_ref.py).  CPU only.

What is being pinned is the argument the device planner's speculation rests on: an extend that comes back TRAPPED leaves both
trees unchanged (rrt.cpp:77-102), the STATE cell of half-iteration h is h and its ACTION cells are h * K + j whatever happened
before, so the next B valid halves of a query may all be evaluated against the SAME (stale) trees — nearest neighbour, the K
candidate pair checks — and taken in order: leading all-invalid segments are TRAPPED, the first segment with a valid candidate
is resolved (newConfig's acceptance, append, connect), everything behind it is dropped and evaluated again next round.  The
result must not depend on B."""
import numpy as np
import pytest

import pyoracle as po
from conftest import load_terrain


def queries(o, n, seed):
    q = o.sample_states(seed, 1, 0, 40 * n)
    q[:, 3:8] = 0; q[:, 3] = 0.5
    v, _ = o.valid_states(q, po.STANCE)
    q = q[v == 1]
    out = []
    for i in range(len(q)):
        d = np.hypot(q[:, 0] - q[i, 0], q[:, 1] - q[i, 1])
        j = np.nonzero((d > 1.0) & (d < 3.0))[0]
        if len(j):
            out.append((q[i], q[j[0]]))
        if len(out) == n:
            break
    return out


class Tree:
    def __init__(self, root):
        self.states, self.actions, self.parent = [np.array(root, dtype=np.float64)], [np.zeros(10)], [-1]

    def push(self, parent, s, a):
        self.states.append(np.array(s)); self.actions.append(np.array(a)); self.parent.append(int(parent))


def speculative_plan(o, start, goal, seed, query, K, max_iters, cap, B):
    """rounds of up to B speculated halves, as k_pipe_prep / k_walk_seg / k_pipe_triage / k_pipe_select / k_pipe_connect run them"""
    Ta, Tb = Tree(start), Tree(goal)
    cell, limit, nn_queries, rounds = 0, 2 * max_iters, 0, 0
    while True:
        it = cell >> 1
        if it >= max_iters:
            return Ta, Tb, False, max_iters, nn_queries, rounds          # budget used up
        if len(Ta.states) >= cap or len(Tb.states) >= cap:
            return Ta, Tb, False, it + 1, nn_queries, rounds             # a tree is full at the start of a half
        rounds += 1
        # prep: the query's next B VALID random states below the budget (invalid ones are skipped, rrt_connect.cpp:254) ...
        segs, c = [], cell
        while len(segs) < B and c < limit:
            s_rand = o.sample_states(seed, query, c, 1)[0]
            if o.valid_states(s_rand[None], po.STANCE)[0][0]:
                segs.append((c, s_rand))
            c += 1
        if not segs:
            cell = c
            continue
        # ... each with the nearest neighbour in the tree its half extends, all against the trees AS THEY ARE NOW; walk: K pair checks each
        evaluated = []
        for c, s_rand in segs:
            T = Ta if (c & 1) == 0 else Tb
            near = int(o.nearest(np.array(T.states), s_rand[None])[0][0])
            normal = o.surface_normal(s_rand[:1], s_rand[1:2])[0]
            a = o.sample_actions(seed, query, c * K, K, normal)
            v, _, sn, _, _ = o.validate_pairs(np.repeat(T.states[near][None], K, 0), a, po.FORWARD if (c & 1) == 0 else po.REVERSE)
            evaluated.append((near, a, v, sn))
        # triage / select / connect: in cell order
        cell = segs[-1][0] + 1
        for (c, s_rand), (near, a, v, sn) in zip(segs, evaluated):
            nn_queries += 1
            if not v.any():
                continue                                                 # TRAPPED at triage
            half = c & 1
            Tx, Ty = (Ta, Tb) if half == 0 else (Tb, Ta)
            j = int(np.argmax(v))                                        # the first valid action decides (rrt.cpp:44-47)
            cell = c + 1                                                 # the segments behind this one are dropped
            if o.distance(sn[j][None], s_rand[None], 1)[0] < o.distance(Tx.states[near][None], s_rand[None], 1)[0]:  # rrt.cpp:55-66
                Tx.push(near, sn[j], a[j])
                nn_queries += 1
                cn = int(o.nearest(np.array(Ty.states), sn[j][None])[0][0])
                st, s2, a2, _ = o.attempt_connect(Ty.states[cn][None], sn[j][None], po.REVERSE if half == 0 else po.FORWARD)
                if st[0] != po.TRAPPED:
                    Ty.push(cn, s2[0], a2[0])
                if st[0] == po.REACHED:
                    return Ta, Tb, True, (c >> 1) + 1, nn_queries, rounds
            break


@pytest.mark.parametrize("name,max_iters,cap", [("synth_mixed", 400, 12), ("slope", 400, 12), ("synth_mixed", 400, 3), ("synth_mixed", 7, 12)])
def test_speculated_halves_equal_the_sequential_planner(name, max_iters, cap):
    """budget stops, capacity stops (a tree full at the start of a half) and solved queries; B = 1 is the sequential schedule itself"""
    T = load_terrain(name)
    o = po.Oracle(T)
    K = 6
    solved_any, grew_any, fewer_rounds = False, False, False
    for qi, (start, goal) in enumerate(queries(o, 5, 5)):
        P = po.PlanParams(K, 0, max_iters, cap, 0, 0, 0)
        st, _, _, ta, tb = o.plan_ex(start, goal, 4, 50 + qi, P)
        rounds = {}
        for B in (1, 4, 8, 16):  # 8 and 16 are the shipped depths (from / below 32,768 queries)
            Ta, Tb, solved, iters, nn_queries, rounds[B] = speculative_plan(o, start, goal, 4, 50 + qi, K, max_iters, cap, B)
            assert (solved, iters, len(Ta.states), len(Tb.states), nn_queries) == (bool(st.solved), st.iters, st.nv_a, st.nv_b, st.nn_queries), (qi, B)
            for mine, ref in ((Ta, ta), (Tb, tb)):
                assert np.array_equal(np.array(mine.states), ref["states"]) and np.array_equal(np.array(mine.parent), ref["parent"]), (qi, B)
                assert np.array_equal(np.array(mine.actions)[1:], ref["actions"][1:]), (qi, B)
        solved_any |= bool(st.solved)
        grew_any |= st.nv_a + st.nv_b > 2
        fewer_rounds |= rounds[16] < rounds[4] < rounds[1]
    assert fewer_rounds and (grew_any or max_iters < 50)  # the comparison saw trees grow, and speculation did shorten the schedule
