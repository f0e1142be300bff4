"""The mbarrier protocol of the tcgen05 prefill attention (csrc/attention_prefill_tc.cu), restated as three cooperating
roles over a model of mbarrier phase / parity semantics and run under randomised interleavings: it must never
deadlock, never overwrite a buffer another role still needs, and consume every block exactly once.  The kernel has not
run on a device yet; a protocol error there would be a hang, so the protocol is checked here.

Model: an mbarrier completes a phase when `count` arrivals have been made; `wait(parity)` passes when the phase of that
parity has completed, i.e. when (completed phases & 1) != parity -- a fresh barrier passes wait(1), blocks wait(0).
tcgen05.commit = one arrival that happens when all MMAs issued before it have completed; MMAs complete in issue order
after an arbitrary delay (the scheduler decides), TMA loads likewise.
"""
import random

import pytest

Q_FULL, KV_FULL, KV_EMPTY, S_FULL, S_EMPTY, P_FULL, P_EMPTY, O_FULL, O_EMPTY = 0, 1, 3, 5, 7, 9, 10, 11, 13
COUNTS = {0: 1, 1: 1, 2: 1, 3: 1, 4: 1, 5: 1, 6: 1, 7: 128, 8: 128, 9: 128, 10: 1, 11: 1, 12: 1, 13: 128, 14: 128}


class Sim:
    def __init__(self, nblk, rng):
        self.nblk, self.rng = nblk, rng
        self.done_phases = {b: 0 for b in COUNTS}
        self.pending = {b: 0 for b in COUNTS}
        self.async_q = []       # in-order queue of async operations (MMA groups, commits, TMA loads) -> callables
        # buffer state for hazard checks
        self.kv = [None, None]  # block held by K/V stage s
        self.s_buf = [None, None]   # ("written", j) / ("read", j)
        self.p_buf = None
        self.o_buf = [None, None]
        self.accumulated = []

    def arrive(self, b, n=1):
        self.pending[b] += n
        assert self.pending[b] <= COUNTS[b], f"barrier {b}: more arrivals than its count"
        if self.pending[b] == COUNTS[b]:
            self.pending[b] = 0
            self.done_phases[b] += 1

    def passes(self, b, parity):
        return (self.done_phases[b] & 1) != parity

    # ---- roles as generators: yield ("wait", bar, parity) to block; everything else runs atomically
    def producer(self):
        self.async_q.append(lambda: self.arrive(Q_FULL))
        for j in range(self.nblk):
            s, ph = j & 1, (j >> 1) & 1
            yield ("wait", KV_EMPTY + s, ph ^ 1)

            def land(j=j, s=s):
                assert self.kv[s] is None or self.kv[s][0] == "free", f"K/V stage {s} overwritten while block {self.kv[s]} in use"
                self.kv[s] = ("full", j)
                self.arrive(KV_FULL + s)
            self.async_q.append(land)

    def mma(self):
        def issue_s(j):
            s, ph = j & 1, (j >> 1) & 1
            yield ("wait", KV_FULL + s, ph)
            yield ("wait", S_EMPTY + s, ph ^ 1)

            def run(j=j, s=s):
                assert self.kv[s] == ("full", j), f"S_{j} read stage {s} holding {self.kv[s]}"
                assert self.s_buf[s] is None or self.s_buf[s][0] == "read", f"S buffer {s} overwritten before it was read"
                self.s_buf[s] = ("written", j)
            self.async_q.append(run)
            self.async_q.append(lambda s=s: self.arrive(S_FULL + s))

        def issue_pv(j):
            s, ph = j & 1, (j >> 1) & 1
            yield ("wait", P_FULL, j & 1)
            yield ("wait", O_EMPTY + s, ph ^ 1)

            def run(j=j, s=s):
                assert self.p_buf == ("written", j), f"PV_{j} read P holding {self.p_buf}"
                assert self.kv[s] == ("full", j)
                assert self.o_buf[s] is None or self.o_buf[s][0] == "read", f"PV buffer {s} overwritten before it was read"
                self.o_buf[s] = ("written", j)
                self.p_buf = ("consumed", j)
                self.kv[s] = ("free", j)
            self.async_q.append(run)
            self.async_q.append(lambda s=s: self.arrive(O_FULL + s))
            self.async_q.append(lambda: self.arrive(P_EMPTY))
            self.async_q.append(lambda s=s: self.arrive(KV_EMPTY + s))

        yield ("wait", Q_FULL, 0)
        yield from issue_s(0)
        for j in range(self.nblk):
            if j + 1 < self.nblk:
                yield from issue_s(j + 1)
            yield from issue_pv(j)

    def softmax(self):
        def accumulate(j):
            s, ph = j & 1, (j >> 1) & 1
            yield ("wait", O_FULL + s, ph)
            assert self.o_buf[s] == ("written", j), f"accumulate({j}) read PV buffer holding {self.o_buf[s]}"
            self.o_buf[s] = ("read", j)
            self.accumulated.append(j)
            self.arrive(O_EMPTY + s, 128)

        for j in range(self.nblk):
            s, ph = j & 1, (j >> 1) & 1
            yield ("wait", S_FULL + s, ph)
            assert self.s_buf[s] == ("written", j), f"softmax({j}) read S buffer holding {self.s_buf[s]}"
            yield ("wait", P_EMPTY, (j & 1) ^ 1)
            assert self.p_buf is None or self.p_buf[0] == "consumed", f"P overwritten while {self.p_buf}"
            self.p_buf = ("written", j)
            self.s_buf[s] = ("read", j)
            self.arrive(S_EMPTY + s, 128)
            self.arrive(P_FULL, 128)
            if j > 0:
                yield from accumulate(j - 1)
        yield from accumulate(self.nblk - 1)

    def run(self):
        roles = {"producer": self.producer(), "mma": self.mma(), "softmax": self.softmax()}
        blocked = {}
        steps = 0
        while roles or self.async_q:
            steps += 1
            assert steps < 100000, "livelock"
            choices = []
            for name in roles:
                if name not in blocked or self.passes(*blocked[name]):
                    choices.append(name)
            if self.async_q:
                choices.append("async")
            assert choices, f"DEADLOCK at nblk={self.nblk}: blocked={blocked}, phases={self.done_phases}"
            pick = self.rng.choice(choices)
            if pick == "async":
                self.async_q.pop(0)()   # asynchronous operations complete in issue order
                continue
            blocked.pop(pick, None)
            try:
                while True:
                    op = next(roles[pick])
                    if op[0] == "wait" and not self.passes(op[1], op[2]):
                        blocked[pick] = (op[1], op[2])
                        break
            except StopIteration:
                del roles[pick]
        return self.accumulated


@pytest.mark.parametrize("nblk", [1, 2, 3, 4, 5, 8, 33])
def test_protocol_has_no_deadlock_or_hazard(nblk):
    for seed in range(300):
        acc = Sim(nblk, random.Random(seed * 1000 + nblk)).run()
        assert acc == list(range(nblk))
