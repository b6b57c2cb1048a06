"""Randomised simulation of the mbarrier protocol of experimental/linear_tcgen05.cu (producer warps,
MMA issuer, epilogue warps; 2 shared-memory stages, 2 TMEM accumulator stages) -- a phase-parity mistake
hangs the GPU, so the protocol is checked here first.  Model: an mbarrier has an arrival count and a phase
bit; try_wait.parity(p) succeeds iff the phase with parity p has completed, i.e. the current phase bit
differs from p (so waiting for parity 1 on a fresh barrier passes).  tcgen05.commit arrives on its barrier
some random time after the MMAs were issued.  Checked: no deadlock, a stage is never overwritten before the
MMAs that read it have completed, an accumulator is never overwritten before the epilogue has drained it,
the epilogue only reads completed accumulators, every tile is produced / computed / stored exactly once.

    python experimental/sim_barriers.py
"""
import random

STAGES, KBLOCKS, NPROD, NEPI = 2, 4, 4, 4   # producer / epilogue agents stand for whole warps here


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.pending = self.count
            self.phase ^= 1

    def done(self, parity):
        return self.phase != parity


def run(tiles, seed):
    rnd = random.Random(seed)
    a_full = [Bar(NPROD) for _ in range(STAGES)]
    a_empty = [Bar(1) for _ in range(STAGES)]
    t_full = [Bar(1) for _ in range(2)]
    t_empty = [Bar(NEPI) for _ in range(2)]
    stage_content = [None] * STAGES          # (tile, kb) currently valid in the stage, per producer-complete
    stage_writers = [set() for _ in range(STAGES)]
    stage_busy = [False] * STAGES            # MMAs reading the stage have not completed yet
    acc_content = [None, None]               # tile whose accumulation completed in the TMEM stage
    acc_busy = [False, False]                # MMAs still writing
    acc_undrained = [False, False]
    pending_commits = []                     # (time_left, fn)
    stored = []
    log = {"computed": []}

    def producer(pid):
        it = 0
        for tile in range(tiles):
            for kb in range(KBLOCKS):
                s, ph = it % STAGES, (it // STAGES) & 1
                while not a_empty[s].done(ph ^ 1):
                    yield
                assert not stage_busy[s], "producer overwrites a stage the tensor core is still reading"
                stage_writers[s].add(pid)
                if len(stage_writers[s]) == NPROD:
                    stage_content[s] = (tile, kb)
                    stage_writers[s] = set()
                a_full[s].arrive()
                it += 1
                yield

    def mma():
        it = 0
        for tcount in range(tiles):
            as_, aph = tcount & 1, (tcount >> 1) & 1
            while not t_empty[as_].done(aph ^ 1):
                yield
            assert not acc_undrained[as_], "MMA overwrites an accumulator the epilogue has not drained"
            for kb in range(KBLOCKS):
                s, ph = it % STAGES, (it // STAGES) & 1
                while not a_full[s].done(ph):
                    yield
                assert stage_content[s] == (tcount, kb), (stage_content[s], tcount, kb)
                stage_busy[s] = True
                acc_busy[as_] = True

                def free_stage(s=s):
                    stage_busy[s] = False
                    a_empty[s].arrive()
                pending_commits.append([rnd.randint(0, 6), free_stage])
                if kb == KBLOCKS - 1:
                    def acc_done(as_=as_, tcount=tcount):
                        acc_busy[as_] = False
                        acc_content[as_] = tcount
                        acc_undrained[as_] = True
                        t_full[as_].arrive()
                    pending_commits.append([rnd.randint(0, 6), acc_done])
                it += 1
                yield

    def epilogue(eid):
        for tcount in range(tiles):
            as_, aph = tcount & 1, (tcount >> 1) & 1
            while not t_full[as_].done(aph):
                yield
            assert acc_content[as_] == tcount and not acc_busy[as_], "epilogue reads an unfinished accumulator"
            yield
            if eid == 0:
                stored.append(tcount)
            t_empty[as_].arrive()
            if t_empty[as_].pending == t_empty[as_].count:   # last arrival of this phase
                acc_undrained[as_] = False
            yield

    agents = [producer(i) for i in range(NPROD)] + [mma()] + [epilogue(i) for i in range(NEPI)]
    alive = list(agents)
    idle = 0
    while alive:
        # commits complete in issue order (tcgen05 ops retire in order)
        if pending_commits:
            pending_commits[0][0] -= 1
            while pending_commits and pending_commits[0][0] <= 0:
                pending_commits.pop(0)[1]()
        ag = rnd.choice(alive)
        try:
            next(ag)
        except StopIteration:
            alive.remove(ag)
        idle += 1
        assert idle < 200000 * (tiles + 1), "deadlock"
    while pending_commits:
        pending_commits.pop(0)[1]()
    assert stored == list(range(tiles)), stored


if __name__ == "__main__":
    for tiles in (1, 2, 3, 5, 8):
        for seed in range(40):
            run(tiles, seed)
    print("barrier protocol ok")
