"""CPU: a discrete model of the stage ring of the tensor-core kernels (csrc/glrm_dense_mma.cuh) — full / empty mbarriers with
phase parities, NST slots, no producer warp: a pass starts with NST - 1 fills issued by warp 0 and the fill of stage
st + NST - 1 is issued by warp (st + NST - 1) % 8 as it starts stage st.  Warps advance in random order; the model checks that
nobody waits forever, that a slot is never refilled before all 8 warps have released its previous content, and that every
warp reads, at stage st of pass p, exactly the content filled for (p, st) — with the kernels' own slot / parity formulas."""
import random

import pytest

W = 8


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0      # phase = number of completed phases

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def done(self, parity):                                           # try_wait.parity: has the phase with this parity completed?
        return (self.phase & 1) != parity


def run(nst, NST, passes, seed):
    rng = random.Random(seed)
    full = [MBar(1) for _ in range(NST)]
    empty = [MBar(W) for _ in range(NST)]
    content = [None] * NST                                            # (fill index) held by each slot
    readers_left = [0] * NST
    # per-warp program counters: (pass, stage, step) with step in {"duty", "wait_full", "release"}
    state = [dict(p=0, st=0, step="prologue" if w == 0 else "duty", cons=0, base=0, j=0) for w in range(W)]
    barrier_waiting = set()                                           # warps parked at the barrier that ends a pass
    reads = []

    def issue(base, st):                                              # kernel: issue(base, st)
        F = base + st
        slot = F % NST
        if not empty[slot].done(((F // NST) & 1) ^ 1):
            return False
        assert readers_left[slot] == 0, "refilled a slot that is still being read"
        content[slot] = F
        readers_left[slot] = W
        full[slot].arrive()                                           # (expect_tx arrive + the copies completing)
        return True

    steps = 0
    while len(barrier_waiting) < W or any(s["p"] < passes for s in state):
        steps += 1
        assert steps < 2_000_000, "no progress: deadlock"
        if len(barrier_waiting) == W:                                 # everybody reached the end of the pass
            barrier_waiting.clear()
            for w, s in enumerate(state):
                s["p"] += 1
                s["st"], s["base"], s["j"] = 0, s["cons"], 0
                s["step"] = "prologue" if w == 0 else "duty"
            if all(s["p"] >= passes for s in state):
                break
            continue
        w = rng.randrange(W)
        s = state[w]
        if w in barrier_waiting or s["p"] >= passes:
            continue
        if s["step"] == "prologue":                                   # pass_prologue: warp 0 issues the first NST - 1 fills
            if s["j"] < min(NST - 1, nst):
                if issue(s["base"], s["j"]):
                    s["j"] += 1
            else:
                s["step"] = "duty"
        elif s["step"] == "duty":
            ft = s["st"] + NST - 1
            if ft < nst and (ft & (W - 1)) == w:
                if not issue(s["base"], ft):
                    continue                                          # spinning on the empty barrier
            s["step"] = "wait_full"
        elif s["step"] == "wait_full":
            F = s["base"] + s["st"]
            slot = F % NST
            if full[slot].done((F // NST) & 1):
                assert content[slot] == F, f"warp {w} expected fill {F} in slot {slot}, found {content[slot]}"
                reads.append((w, F))
                s["step"] = "release"
        else:
            F = s["base"] + s["st"]
            slot = F % NST
            readers_left[slot] -= 1
            empty[slot].arrive()
            s["cons"] += 1
            s["st"] += 1
            if s["st"] == nst:
                barrier_waiting.add(w)
            else:
                s["step"] = "duty"
    return reads


@pytest.mark.parametrize("nst,NST", [(5, 4), (15, 4), (59, 3), (7, 2), (3, 2), (9, 4)])
def test_rotating_producer_ring_is_live_and_consistent(nst, NST):
    for seed in range(5):
        reads = run(nst, NST, passes=3, seed=seed)
        assert len(reads) == W * nst * 3
