"""Discrete-event model of the shared-memory ring of the register-window gather
(csrc/exchange_win.cu): mbarrier phases / parities, the prefetch distance, and the two
copy schedules (variant 1: every warp copies a share of every row; variant 2: warp
r mod n_warps copies the whole row of record r).  Warps are stepped in random order;
copies land after random delays.  Checked: no deadlock, every warp reads record r from
stage r mod kStages while it holds record r, and no stage is overwritten before all
warps released it.  (The kernels themselves are tested on the GPU; this pins the
protocol the code implements, including the parity arithmetic.)"""
import random

import pytest

K_STAGES, K_AHEAD = 12, 8


class MBarrier:
    """mbarrier with a fixed arrival count; try_wait(parity) as in PTX: true once the
    phase of that parity has completed (the phase before the first counts as done)."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def try_wait(self, parity):
        return (self.phase & 1) != parity


class Warp:
    def __init__(self, sim, w):
        self.sim, self.w = sim, w
        self.r = 0                          # record being consumed
        self.state = "prologue"
        self.duty = w                       # variant 2: next record to copy
        self.p_rec = 0                      # variant 1: next record to issue
        self.done = False

    # -- copy side ----------------------------------------------------------
    def _copy_target(self):
        s = self.sim
        if s.variant == 1:
            return self.p_rec if self.p_rec < s.n_rec else None
        return self.duty if self.duty < s.n_rec else None

    def _try_copy(self):
        """One attempt to issue the pending copy; False if blocked on empty[]."""
        s = self.sim
        rec = self._copy_target()
        stage, parity = rec % K_STAGES, (rec // K_STAGES) & 1
        if not s.empty[stage].try_wait(parity ^ 1):
            return False
        # the stage must not hold unread data: everything older was released
        assert s.readers_left[stage] == 0, (rec, stage)
        s.in_flight.append([random.randint(0, 6), stage, rec, self.w])
        if s.variant == 1:
            self.p_rec += 1
        else:
            self.duty += s.n_warps
        return True

    def step(self):
        s = self.sim
        if self.done:
            return False
        if self.state == "prologue":
            tgt = self._copy_target()
            if tgt is not None and tgt < K_AHEAD:
                return self._try_copy()
            self.state = "copy"
            return True
        if self.state == "copy":
            tgt = self._copy_target()
            need = tgt is not None and (s.variant == 1 or tgt <= self.r + K_AHEAD)
            if need and not self._try_copy():
                return False
            self.state = "wait"
            return True
        if self.state == "wait":
            stage, parity = self.r % K_STAGES, (self.r // K_STAGES) & 1
            if not s.full[stage].try_wait(parity):
                return False
            assert s.content[stage] == self.r, (self.w, self.r, s.content[stage])
            s.readers_left[stage] -= 1
            s.empty[stage].arrive()
            self.r += 1
            if self.r == s.n_rec:
                self.done = True
            else:
                self.state = "copy"
            return True
        raise AssertionError(self.state)


class Sim:
    def __init__(self, variant, n_warps, n_rec, seed):
        random.seed(seed)
        self.variant, self.n_warps, self.n_rec = variant, n_warps, n_rec
        arrivals = n_warps if variant == 1 else 1     # per-thread counts scaled to warps
        self.full = [MBarrier(arrivals) for _ in range(K_STAGES)]
        self.empty = [MBarrier(n_warps) for _ in range(K_STAGES)]
        self.content = [None] * K_STAGES
        self.readers_left = [0] * K_STAGES
        self.partial = {}                              # (stage, rec) -> shares landed
        self.in_flight = []
        self.warps = [Warp(self, w) for w in range(n_warps)]

    def land_copies(self):
        for c in list(self.in_flight):
            c[0] -= 1
            if c[0] <= 0:
                _, stage, rec, _ = c
                self.in_flight.remove(c)
                key = (stage, rec)
                self.partial[key] = self.partial.get(key, 0) + 1
                if self.partial[key] == (self.n_warps if self.variant == 1 else 1):
                    self.content[stage] = rec
                    self.readers_left[stage] = self.n_warps
                self.full[stage].arrive()

    def run(self):
        idle = 0
        while not all(w.done for w in self.warps):
            self.land_copies()
            progressed = False
            for w in random.sample(self.warps, len(self.warps)):
                if random.random() < 0.7:
                    progressed |= w.step()
            idle = 0 if (progressed or self.in_flight) else idle + 1
            assert idle < 200, "deadlock: " + str([(w.state, w.r) for w in self.warps])


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("n_warps", [1, 2, 4, 5, 8])
@pytest.mark.parametrize("n_rec", [1, 7, 8, 9, 12, 13, 31, 100])
def test_ring_protocol(variant, n_warps, n_rec):
    for seed in range(3):
        Sim(variant, n_warps, n_rec, seed).run()
