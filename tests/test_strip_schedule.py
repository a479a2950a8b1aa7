"""CPU replay of the cluster strip kernel's schedule (lws_b200/csrc/kernels_batch.cu).

The kernel's correctness rests on a dependency argument: blocks of 8 (or 4) bins, frames 2 (or 3)
blocks apart, sweeps Q frames apart, strips NBr macro-steps apart, rows living in a ring of R slots
that TMA fills LEAD frames ahead and drains when a frame's last sweep is done, edge blocks
copied into the neighbour strip's halo.  This test replays exactly that control flow in numpy
-- per-strip rings with halos, slot reuse, lock-stepped strips, housekeeping at the same
macro-steps -- with *concurrent semantics*: inside a macro-step every thread reads what the
rings held when the step began (plus its own writes), and no thread may read a cell that a
different thread writes in the same step.  The result must equal the sequential oracle.
The plan comes from the product's own planner (lwsb_debug_plan_strips).
"""
import numpy as np
import pytest

from conftest import SMALL_CASES, golden, relF

import lws_b200
from lws_b200 import _native, dsp

SL, SLEAD = 5, 2


def _tables(W, fold, Q):
    return [_native.debug_terms(W, fold, Q, 1, p) for p in range(Q)]


class Strip(object):
    def __init__(self, c, plan, Nreal):
        self.c = c
        SBK = plan["block_bins"]
        self.NBr, self.R = plan["blocks_per_strip"], plan["ring_rows"]
        self.b0 = c * self.NBr * SBK
        self.nb_my = min(max((Nreal - self.b0 + SBK - 1) // SBK, 0), self.NBr)
        self.width = SBK * self.NBr + 2 * SL
        self.ring = np.full((self.R, self.width), np.nan + 0j, dtype=np.complex128)
        self.tag = [-1] * self.R      # which extended row a slot holds


def replay(E, A, terms, thr_list, mean, Q, T, Nreal, plan):
    """One call of `len(thr_list)` sweeps on the extended spectrogram E (modified in place)."""
    C, NBr, NBV, NS, G, R = (plan[k] for k in ("cluster", "blocks_per_strip", "virtual_blocks", "frame_slots",
                                                "sweeps_per_pass", "ring_rows"))
    QS = plan["sweep_lag"]
    GX, LEAD = plan["sweep_extra_from"], plan["load_lead"]   # sweep slots >= GX run one more frame behind; frames of load look-ahead
    off = lambda g: QS * g + (1 if g >= GX else 0)
    SBK = plan["block_bins"]
    LAGB = (SBK + SL + SBK - 1) // SBK   # blocks between consecutive frames: LAGB * SBK >= SBK + L
    assert QS >= Q and NBV % LAGB == 0 and NS * LAGB == NBV
    Tp = T + 2 * (Q - 1)
    amax = A[Q - 1:Q - 1 + T, SL:SL + Nreal].max()
    act = [i for i, th in enumerate(thr_list) if th * mean < amax]
    npass = (len(act) + G - 1) // G
    assert R >= off(G - 1) + 2 * Q + LEAD + NS and LEAD >= 1 and (GX == G or LEAD == SLEAD - 1)
    for ps in range(npass):
        Gp = min(G, len(act) - ps * G)
        nsteps = LAGB * (T - 1 + off(Gp - 1)) + NBV
        strips = [Strip(c, plan, Nreal) for c in range(C)]

        def load(st, e):
            g0 = st.b0  # ring column 0 <-> extended column b0 (= bin b0 - SL)
            st.ring[e % R, :] = 0
            hi = min(g0 + st.width, E.shape[1])
            st.ring[e % R, :hi - g0] = E[e, g0:hi]
            st.tag[e % R] = e

        for st in strips:
            for e in range(min(Tp, 2 * (Q - 1) + LEAD + 1)):
                load(st, e)
        for wall in range(nsteps + (C - 1) * NBr):
            snaps = [st.ring.copy() for st in strips]
            writes = {}   # (strip, slot, col) -> writer id
            reads = []    # (strip, slot, col, reader id)
            for st in strips:
                t = wall - st.c * NBr   # lock step: strip c runs NBr macro-steps behind strip c-1
                if not (0 <= t < nsteps):
                    continue
                for g in range(Gp):
                    for j in range(NS):
                        d = t - LAGB * j
                        if d < 0:
                            continue
                        xb, m = d % NBV, j + NS * (d // NBV) - off(g)
                        if not (xb < st.nb_my and 0 <= m < T):
                            continue
                        me = (st.c, g, j)
                        e = m + Q - 1
                        for dr in range(-(Q - 1), Q):
                            assert st.tag[(e + dr) % R] == e + dr, "row %d not resident (strip %d, t %d)" % (e + dr, st.c, t)
                        own = {}  # this thread's writes of this step: (slot, col) -> value
                        for i in range(SBK):
                            n = st.b0 + SBK * xb + i
                            if n >= Nreal:
                                continue
                            a = A[e, n + SL]
                            if not (a > thr_list[act[ps * G + g]] * mean):
                                continue
                            col = SL + SBK * xb + i
                            dr_, dk_, co = terms[n % Q]
                            vals = np.empty(len(co), dtype=np.complex128)
                            for q in range(len(co)):
                                key = ((e + dr_[q]) % R, col + dk_[q])
                                if key in own:
                                    vals[q] = own[key]
                                else:
                                    vals[q] = snaps[st.c][key]
                                    reads.append((st.c,) + key + (me,))
                            tsum = np.sum(co * vals)
                            if not (abs(tsum) > 0):
                                continue
                            val = tsum * a / abs(tsum)
                            targets = [(st.c, e % R, col, val)]
                            if 1 <= n <= SL and st.c == 0:
                                targets.append((st.c, e % R, SL - n, np.conj(val)))
                            elif Nreal - 1 - SL <= n <= Nreal - 2:
                                targets.append((st.c, e % R, SL + 2 * (Nreal - 1) - n - st.b0, np.conj(val)))
                            q = SBK * xb + i   # bin inside the strip: the first / last L bins also live in a neighbour's halo
                            if q < SL and st.c > 0:
                                targets.append((st.c - 1, e % R, SL + SBK * NBr + q, val))
                            if q >= SBK * NBr - SL and st.c < C - 1:
                                targets.append((st.c + 1, e % R, q - (SBK * NBr - SL), val))
                            for (sc, slot, cc, vv) in targets:
                                assert strips[sc].tag[slot] == e, "halo write into a slot holding another row"
                                strips[sc].ring[slot, cc] = vv
                                writes[(sc, slot, cc)] = me
                                if sc == st.c:
                                    own[(slot, cc)] = vv
            for (sc, slot, cc, reader) in reads:
                w = writes.get((sc, slot, cc))
                assert w is None or w == reader, "cell read and written by different threads in one macro-step"
            # housekeeping after the step (control warp)
            for st in strips:
                t = wall - st.c * NBr
                if not (0 <= t < nsteps):
                    continue
                tf = t - (st.nb_my - 1)
                if st.nb_my > 0 and tf >= 0 and tf % LAGB == 0:
                    m = tf // LAGB - off(Gp - 1)
                    if 0 <= m < T:
                        e = m + Q - 1
                        assert st.tag[e % R] == e
                        lo = 0 if st.c == 0 else SL
                        hi = SL + (Nreal - st.b0) + SL if st.c == C - 1 else SL + SBK * NBr
                        E[e, st.b0 + lo:st.b0 + hi] = st.ring[e % R, lo:hi]
                if (t + 1) % LAGB == 0:
                    e = (t + 1) // LAGB + 2 * (Q - 1) + LEAD
                    if e < Tp:
                        load(st, e)
    return E


CASES = [  # (golden case, frames, thresholds, smem budget, cluster, sweeps per pass[, bins per block])
    ("q4", 20, [0.5, 0.2, 0.3, 0.1, 0.05], 232448, 0, 0),          # one strip, all sweeps in one pass
    ("q4", 9, [0.5, 0.2, 0.3, 0.1, 0.05], 232448, 0, 2),           # several passes, a partial last pass
    ("q4b", 24, [0.9, 0.4, 0.2, 0.3], 232448, 2, 0),               # two strips (33 bins: 3 + 2 blocks)
    ("q4b", 3, [0.9, 0.4, 0.2], 232448, 2, 2),                     # fewer frames than Q, two strips, two passes
    ("q8", 12, [0.6, 0.3, 0.2], 232448, 2, 0),                     # Q = 8, two strips
    ("q2", 16, [0.5, 0.25, 0.2, 0.1], 232448, 0, 3),               # Q = 2
    ("cfg1_short", 10, [0.8, 0.3, 0.2, 0.25, 0.1, 0.3], 232448, 4, 2),  # 257 bins on 4 strips (9 blocks each, 6 in the last)
    ("rand513", 5, [0.8, 0.3], 232448, 8, 0),                      # 513 bins on 8 strips of 9 blocks (virtual 10)
    ("cfg1_short", 6, [0.8, 0.3, 0.5], 60000, 2, 0),               # a tight ring: small G forced by the budget
    ("rand513", 6, [0.8, 0.3, 0.2, 0.5, 0.1, 0.3, 0.2, 0.4, 0.1], 232448, 2, 7),  # BASELINE configs[1]'s plan: 17 frame slots, 7 sweeps per pass,
                                                                                     # one extra frame of lag from sweep slot 4 on, two passes
    # 4-bin blocks, frames 3 blocks apart; the L = 5 halo bins span two blocks
    ("q4", 20, [0.5, 0.2, 0.3, 0.1, 0.05], 232448, 0, 0, 4),
    ("q4", 9, [0.5, 0.2, 0.3, 0.1, 0.05], 232448, 0, 2, 4),
    ("q4b", 24, [0.9, 0.4, 0.2, 0.3], 232448, 2, 0, 4),
    ("q4b", 3, [0.9, 0.4, 0.2], 232448, 2, 2, 4),
    ("q2", 16, [0.5, 0.25, 0.2, 0.1], 232448, 0, 3, 4),
    ("cfg1_short", 10, [0.8, 0.3, 0.2, 0.25, 0.1, 0.3], 232448, 4, 2, 4),
    ("rand513", 5, [0.8, 0.3], 232448, 8, 0, 4),
    ("rand513", 4, [0.8, 0.3, 0.1], 232448, 4, 2, 4),
    ("cfg1_short", 6, [0.8, 0.3, 0.5], 60000, 2, 0, 4),
]
CASES = [c if len(c) == 7 else c + (8,) for c in CASES]


@pytest.mark.parametrize("name,T,thr,smem,cluster,sweeps,block", CASES,
                         ids=["%s-T%d-C%d-G%d-B%d" % (c[0], c[1], c[4], c[5], c[6]) for c in CASES])
def test_strip_schedule_equals_sequential(oracle, name, T, thr, smem, cluster, sweeps, block):
    if name == "cfg1_short":
        args, kw = (512, 128), {}
    elif name == "rand513":
        args, kw = (1024, 256), {}
    else:
        case = [c for c in SMALL_CASES if c["name"] == name][0]
        args, kw = case["args"], case["kwargs"]
    p = lws_b200.lws(*args, **kw)
    po = oracle.lws(*args, **kw)
    Q, L = p.W.shape[1], p.W.shape[2] - 1
    assert L == SL
    if name == "rand513":
        A0 = np.abs(np.random.default_rng(7).standard_normal((T, 513)))
    else:
        A0 = np.abs(golden(name)["X"])[:T]
    T, Nreal = A0.shape
    thr = np.asarray(thr, dtype=np.float64)
    plan = _native.debug_plan_strips(Nreal, Q, L, len(thr), T, 1, smem_limit=smem, cluster=cluster, sweeps=sweeps, block=block)
    assert plan is not None and plan["block_bins"] == block
    SBK = block
    if cluster:
        assert plan["cluster"] == cluster
    fold = {2: 2, 4: 4}.get(Q, 0)
    terms = _tables(p.W, fold, Q)
    # extended spectrogram wide enough for whole blocks + halo of the last strip
    E = dsp.extspec(A0.astype(np.complex128), L, Q)
    width = plan["cluster"] * plan["blocks_per_strip"] * SBK + 2 * SL
    Ew = np.zeros((E.shape[0], max(width, E.shape[1])), dtype=np.complex128)
    Ew[:, :E.shape[1]] = E
    Aw = np.abs(Ew)
    mean = np.mean(A0)
    replay(Ew, Aw, terms, thr, mean, Q, T, Nreal, plan)
    got = Ew[Q - 1:Q - 1 + T, L:L + Nreal]
    want = po.batch_lws(A0, thresholds=thr)
    assert relF(got, want) < 1e-11, plan


def test_planner_properties():
    """Every plan the planner can emit satisfies the constraints the kernel relies on."""
    for Nreal in (17, 33, 65, 129, 257, 513, 1025, 2049, 4097, 25, 41):
        for Q in (2, 4, 8):
            for iters in (1, 7, 100, 200):
                for smem in (232448, 100000, 30000):
                    for cluster, block in ((0, 0), (1, 0), (2, 0), (4, 0), (8, 0), (0, 4), (2, 4), (4, 4), (8, 4), (0, 8), (4, 8)):
                        pl = _native.debug_plan_strips(Nreal, Q, 5, iters, 600, 64, smem_limit=smem, cluster=cluster, block=block)
                        if pl is None:
                            continue
                        SBK = pl["block_bins"]
                        LAGB = (SBK + SL + SBK - 1) // SBK
                        assert SBK in (4, 8) and (block == 0 or SBK == block) and SBK % Q == 0 or Q % SBK == 0 and SBK == 8
                        C, NBr, NBV, NS, G, R = (pl[k] for k in ("cluster", "blocks_per_strip", "virtual_blocks",
                                                                 "frame_slots", "sweeps_per_pass", "ring_rows"))
                        assert NBV % LAGB == 0 and NBV >= NBr and NS * LAGB == NBV and NBr >= 2 and SBK * NBr >= 2 * SL
                        assert C * NBr * SBK >= Nreal                          # the strips cover every bin
                        assert (C - 1) * NBr * SBK <= Nreal - 1 - SL           # mirror zone inside the last strip
                        extra = 1 if pl["sweep_extra_from"] < G else 0
                        assert R >= pl["sweep_lag"] * (G - 1) + extra + 2 * Q + pl["load_lead"] + NS and 1 <= G <= iters and pl["sweep_lag"] >= Q
                        assert pl["load_lead"] == SLEAD - extra
                        assert pl["ring_pitch"] % 2 == 1 and pl["ring_pitch"] >= SBK * NBr + 2 * SL
                        assert pl["smem_bytes"] <= smem and pl["threads"] <= 256 and pl["threads"] >= NS * G + 32 and (not pl["tensor_memory"] or NS * G <= 128)
                        assert R * pl["ring_pitch"] * 16 + R * 8 + 32 + 4 * iters <= pl["smem_bytes"]
    assert _native.debug_plan_strips(513, 3, 5, 10, 100, 1) is None     # Q must divide the block size
    assert _native.debug_plan_strips(513, 4, 7, 10, 100, 1) is None     # L is fixed at 5


def test_work_items_are_pass_major_and_producers_come_first():
    """The strip kernel's work list: every pass of every utterance exactly once, and the previous pass of the same
    utterance always earlier in the list (clusters take items in increasing order, so a waiting item's producer is
    finished or running: no dead-lock)."""
    rng = np.random.default_rng(5)
    for B, G in ((1, 7), (5, 3), (64, 7), (9, 1), (3, 100)):
        act = [int(x) for x in rng.integers(0, 101, B)]
        act[0] = 100
        items = _native.debug_work_items(act, G)
        want = {(b, ps) for b in range(B) for ps in range((act[b] + G - 1) // G)}
        assert len(items) == len(want) and set(items) == want
        pos = {it: i for i, it in enumerate(items)}
        assert all(pos[(b, ps - 1)] < pos[(b, ps)] for (b, ps) in items if ps > 0)
        assert [ps for _, ps in items] == sorted(ps for _, ps in items)   # pass-major


def test_planner_picks_the_measured_best_plans_for_the_baseline_shapes():
    """The cost model was fitted on B200: for BASELINE configs[1] (67 of 100 sweeps can move a bin with
    the default thresholds) and all-active it must land on the plan measured fastest -- cluster 2, 7 sweeps per pass,
    8-bin blocks -- and for configs[4] (Q = 8, 4 utterances per GPU) on cluster 8; 4-bin blocks only on request."""
    for active in (60, 67, 71, 75, 80, 90, 100):
        pl = _native.debug_plan_strips(513, 4, 5, active, 628, 64)
        assert (pl["cluster"], pl["sweeps_per_pass"], pl["block_bins"], pl["sweep_lag"]) == (2, 7, 8, 4), pl
        assert pl["frame_slots"] == 17 and pl["sweep_fastest"] == 2, pl  # 16 + 1 frame slots: the rotating lane order
    pl = _native.debug_plan_strips(1025, 8, 5, 150, 5632, 4)
    assert pl["cluster"] == 8 and pl["block_bins"] == 8 and pl["sweep_lag"] >= 8, pl
    assert _native.debug_plan_strips(513, 4, 5, 67, 628, 64, block=4)["block_bins"] == 4
    # a single utterance: several passes in flight on different clusters rather than all sweeps in one pass
    pl = _native.debug_plan_strips(513, 4, 5, 67, 628, 1)
    assert (67 + pl["sweeps_per_pass"] - 1) // pl["sweeps_per_pass"] >= 4, pl


def test_rotating_lane_order_covers_every_task_and_is_conflict_free():
    """The rotating lane order of the strip kernel (strip_body.inc, GFAST == 2; kernels_batch.cu, planner order 2) for strips of
    16 + 1 frame slots: at every macro-step phase `ph` (slots 0 .. ph have wrapped to their next frame) half-warp h holds sweep
    slot h with the 16 frame slots other than `ph`, the left-over tasks share the lanes from 16 G on.  The map must hit every
    (frame slot, sweep slot) exactly once, and the frame residues mod 8 -- the shared-memory bank group of a lane, ring_off --
    must occur exactly twice in every full half-warp: no bank conflict."""
    NS, QS = 17, 4
    for G in range(1, 8):
        for ph in range(NS):
            seen = {}
            for tix in range(17 * G):
                spill = tix >= 16 * G
                g = tix - 16 * G if spill else tix >> 4
                i = tix & 15
                j = ph if spill else (i if i < ph else i + 1)
                assert 0 <= j < NS and 0 <= g < G
                assert (j, g) not in seen
                seen[(j, g)] = tix
            assert len(seen) == NS * G
            for h in range(G):
                res = [(j + (NS if j <= ph else 0) - QS * h) % 8 for j in range(NS) if j != ph]
                assert sorted(res) == sorted(list(range(8)) * 2), (G, ph, h, res)

