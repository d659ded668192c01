"""Discrete-event model of the mbarrier protocol of the row-tile DIN kernels (csrc/din_rt.cu, csrc/din_rt64.cu),
run on the CPU with WARP-level actors.

Both kernels share one protocol (ring of 4 slots, 3 gate-accumulator buffers, two consumers with two pooled
buffers each).  Every wait / arrive / commit is transcribed with the parity expression of the source; the actors
(5 gather warps, 2 builder warps + the weight-image loader, the MMA issuer, 2 x 4 consumer warps, the tensor pipe
that retires commits in issue order) run under random interleavings in which single warps may stall for a long
time.  Checked:

  * no deadlock: every actor finishes every group;
  * no parity aliasing: when a wait passes, the barrier has completed exactly the phase the code meant;
  * no mixed phase: all arrivals that complete a phase belong to the same tile (a warp that runs one tile ahead
    of its siblings must not complete their phase for them);
  * hazards: a ring slot / accumulator buffer / pooling-weight buffer is not rewritten while an MMA or a reader
    still uses it.

    python profiles/exp/rt_protocol_sim.py                 # the protocol as shipped (w_ready[q][u])
    python profiles/exp/rt_protocol_sim.py --single-wready # the round-1 protocol: one w_ready barrier per consumer
"""
import random
import sys

SLOTS = 4


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.done = name, count, count, 0
        self.tags = set()

    def arrive(self, n=1, tag=None):
        self.pending -= n
        assert self.pending >= 0, "over-arrival on %s" % self.name
        if tag is not None:
            self.tags.add(tag)
        if self.pending == 0:
            assert len(self.tags) <= 1, "%s: phase %d completed by arrivals of different tiles %s" % (
                self.name, self.done, sorted(self.tags))
            self.tags = set()
            self.done += 1
            self.pending = self.count

    def passed(self, parity):
        return (self.done & 1) != parity


class Deadlock(Exception):
    pass


class Sim:
    def __init__(self, groups, seed, single_wready, stall):
        self.rng = random.Random(seed)
        self.groups = groups                      # tiles per group of this CTA
        self.single = single_wready
        self.stall = stall
        b = Bar
        self.full = [b("full%d" % i, 160 + 64) for i in range(SLOTS)]
        self.empty = [b("empty%d" % i, 1) for i in range(SLOTS)]
        self.d1_full = [b("d1_full%d" % i, 1) for i in range(3)]
        if single_wready:
            self.w_ready = [[b("w_ready%d" % q, 128)] * 2 for q in range(2)]
        else:
            self.w_ready = [[b("w_ready%d%d" % (q, u), 128) for u in range(2)] for q in range(2)]
        self.d2_full = [[b("d2_full%d%d" % (q, u), 1) for u in range(2)] for q in range(2)]
        self.pipe = []
        self.busy = {}
        self.readers = {}                         # resource -> warps that have not read generation K yet
        self.group_sync = [0] * len(groups)       # arrivals at the end-of-group __syncthreads
        self.n_warps = 5 + 2 + 1 + 8
        self.ticks = 0

    # ---- primitives -------------------------------------------------------------------------------
    def wait(self, bar, parity, intended):
        while not bar.passed(parity):
            yield
        assert bar.done == intended + 1, "%s: wait(parity %d) meant completion #%d, barrier has completed %d" % (
            bar.name, parity, intended, bar.done)

    def issue_mma(self, res):
        for r in res:
            self.busy[r] = self.busy.get(r, 0) + 1
        self.pipe.append(("mma", res))

    def commit(self, bar, tag=None):
        self.pipe.append(("commit", bar, tag))

    def touch(self, res, what):
        assert self.busy.get(res, 0) == 0, "%s while an MMA still uses %s" % (what, res)

    def sync_group(self, j):
        self.group_sync[j] += 1
        while self.group_sync[j] < self.n_warps:
            yield

    def wr(self, q, K):
        u = (K >> 1) & 1
        if self.single:
            return self.w_ready[q][0], (K >> 1) & 1, K >> 1
        return self.w_ready[q][u], (K >> 2) & 1, K >> 2

    # ---- actors -------------------------------------------------------------------------------------
    def gather_warp(self, w):
        kbase = 0
        for j, n in enumerate(self.groups):
            def gather(k):
                K = kbase + k
                slot = K % SLOTS
                if K >= SLOTS:
                    yield from self.wait(self.empty[slot], ((K // SLOTS) + 1) & 1, K // SLOTS - 1)
                self.touch("A%d" % slot, "cp.async into tile %d" % K)
            for a in range(2):
                if a < n:
                    yield from gather(a)
                yield
            for k in range(n):
                self.full[(kbase + k) % SLOTS].arrive(32, kbase + k)
                yield
                if k + 2 < n:
                    yield from gather(k + 2)
                yield
            kbase += n
            yield from self.sync_group(j)

    def builder_warp(self, w):
        kbase = 0
        for j, n in enumerate(self.groups):
            for k in range(n):
                K = kbase + k
                slot = K % SLOTS
                if K >= SLOTS:
                    yield from self.wait(self.empty[slot], ((K // SLOTS) + 1) & 1, K // SLOTS - 1)
                self.touch("B%d" % slot, "W_r of tile %d" % K)
                self.full[slot].arrive(32, K)
                yield
            if w == 0:                                         # weight images land in the slots as they retire
                tail = min(n, SLOTS)
                for i in range(tail):
                    K = kbase + n - tail + i
                    yield from self.wait(self.empty[K % SLOTS], (K // SLOTS) & 1, K // SLOTS)
                    self.touch("A%d" % (K % SLOTS), "weight image over tile %d" % K)
                    yield
            kbase += n
            yield from self.sync_group(j)

    def issuer(self):
        kbase = 0
        for j, n in enumerate(self.groups):
            def mma1(k):
                K = kbase + k
                slot, db = K % SLOTS, K % 3
                yield from self.wait(self.full[slot], (K // SLOTS) & 1, K // SLOTS)
                assert not self.readers.get("D1_%d" % db), "gate accumulators %d rewritten by tile %d while %s still read them" % (
                    db, K, sorted(self.readers["D1_%d" % db]))
                self.issue_mma(["A%d" % slot, "B%d" % slot])
                self.readers["D1_%d" % db] = {(K & 1, w) for w in range(4)}
                self.commit(self.d1_full[db])
            for k in range(min(3, n)):
                yield from mma1(k)
                yield
            for k in range(n):
                K = kbase + k
                q, u = K & 1, (K >> 1) & 1
                bar, par, idx = self.wr(q, K)
                yield from self.wait(bar, par, idx)
                assert not self.readers.get("W2_%d%d" % (q, u)), "pooling weights of tile %d incomplete" % K
                self.issue_mma(["A%d" % (K % SLOTS), "W2_%d%d" % (q, u)])
                self.commit(self.d2_full[q][u])
                self.commit(self.empty[K % SLOTS])
                yield
                if k + 3 < n:
                    yield from mma1(k + 3)
                yield
            kbase += n
            yield from self.sync_group(j)

    def consumer_warp(self, q, w):
        kbase = 0
        for j, n in enumerate(self.groups):
            def pool_out(k):
                K = kbase + k
                u = (K >> 1) & 1
                yield from self.wait(self.d2_full[q][u], (K >> 2) & 1, K >> 2)
            first = (q - kbase) & 1
            ks = list(range(first, n, 2))
            for k in ks:
                K = kbase + k
                u = (K >> 1) & 1
                yield from self.wait(self.d1_full[K % 3], (K // 3) & 1, K // 3)
                for _ in range(self.stall(self.rng)):             # the gate pass of this warp (may be slow)
                    self.ticks += 1
                    yield
                self.readers["D1_%d" % (K % 3)].discard((q, w))
                self.touch("W2_%d%d" % (q, u), "pooling weights of tile %d" % K)
                bar, _, _ = self.wr(q, K)
                bar.arrive(32, K)
                yield
                if k - 2 >= 0:
                    yield from pool_out(k - 2)
                yield
            if ks:
                yield from pool_out(ks[-1])
            kbase += n
            yield from self.sync_group(j)

    # ---- scheduler ----------------------------------------------------------------------------------
    def run(self):
        actors = [self.gather_warp(w) for w in range(5)] + [self.builder_warp(w) for w in range(2)] + \
                 [self.issuer()] + [self.consumer_warp(q, w) for q in range(2) for w in range(4)]
        live = list(range(len(actors)))
        idle = 0
        while live:
            progressed = False
            if self.pipe and self.rng.random() < 0.5:         # the tensor pipe retires its oldest operation
                op = self.pipe.pop(0)
                if op[0] == "mma":
                    for r in op[1]:
                        self.busy[r] -= 1
                else:
                    op[1].arrive(1)
                progressed = True
            i = self.rng.choice(live)
            before = self.snapshot()
            try:
                next(actors[i])
            except StopIteration:
                live.remove(i)
                progressed = True
            if progressed or self.snapshot() != before:
                idle = 0
            else:
                idle += 1
                if idle > 20000 and not self.pipe:
                    raise Deadlock("no progress; live actors %s" % live)

    def snapshot(self):
        bars = self.full + self.empty + self.d1_full + [x for r in self.w_ready for x in r] + \
               [x for r in self.d2_full for x in r]
        return tuple((x.done, x.pending) for x in bars) + tuple(self.group_sync) + (len(self.pipe), self.ticks)


def check(single, runs=300, seed0=0, verbose=False, slots=4):
    """Returns the list of (seed, shape, message) failures.  slots: 4 = din_rt64's ring, 8 = din_rt's."""
    global SLOTS
    SLOTS = slots
    failures = []
    shapes = [[64], [56, 56], [14, 14, 14], [64, 64, 64], [2, 6, 2], [7, 5], [100]]
    stalls = [lambda r: 0,
              lambda r: 3 if r.random() < 0.5 else 0,
              lambda r: 4000 if r.random() < 0.02 else 1]       # now and then one warp is very late
    for s in range(runs):
        shape = shapes[s % len(shapes)]
        stall = stalls[(s // len(shapes)) % len(stalls)]
        try:
            Sim(shape, seed0 + s, single, stall).run()
        except (AssertionError, Deadlock) as e:
            failures.append((seed0 + s, shape, "%s: %s" % (type(e).__name__, e)))
            if verbose:
                print("seed %d shape %s -> %s" % failures[-1])
    return failures


if __name__ == "__main__":
    single = "--single-wready" in sys.argv
    bad = 0
    for slots in (4, 8):
        f = check(single, verbose=True, slots=slots)
        print("%d-slot ring, %s protocol: %d of 300 random schedules failed" % (
            slots, "single-w_ready" if single else "w_ready[q][u]", len(f)))
        bad += len(f)
    sys.exit(1 if bad and not single else 0)
