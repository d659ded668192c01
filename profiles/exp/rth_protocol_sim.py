"""Discrete-event model of din_rth_kernel's mbarrier protocol (csrc/din_rth.cu), run on the CPU.

The kernel could not be run on a GPU in the round it was written, and a wrong mbarrier parity
shows up there as a hang.  This script transcribes every wait / arrive / commit of the kernel -
with the parity expressions copied from the source - into cooperating actors (gather warps,
builder warp, issuer warp, consumer warps, the tensor pipe that retires commits in issue order,
the copy engine behind cp.async.bulk) and runs them under random interleavings.  It checks

  * no deadlock: every actor reaches the end of every group;
  * no parity aliasing: when a wait passes, the barrier has completed exactly the phase the code
    meant (the intended completion index is stated next to each wait);
  * operand hazards: a ring slot / weight buffer / accumulator is never rewritten while an MMA or
    a reader that uses it is outstanding.

    python profiles/exp/rth_protocol_sim.py        # exhaustive over small shapes, random schedules
"""
import random
import sys

SLOTS, AHEAD = 3, 2


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.done, self.tx = name, count, count, 0, 0

    def arrive(self, n=1):
        self.pending -= n
        assert self.pending >= 0, "over-arrival on %s" % self.name
        self._maybe()

    def expect(self, tx):
        self.tx += tx

    def complete_tx(self, tx):
        self.tx -= tx
        self._maybe()

    def _maybe(self):
        if self.pending == 0 and self.tx == 0:
            self.done += 1
            self.pending = self.count

    def passed(self, parity):
        return (self.done & 1) != parity


class Sim:
    def __init__(self, tiles_per_group, seed, builder_gathers=False):
        self.rng = random.Random(seed)
        self.groups = tiles_per_group
        self.bg = builder_gathers              # SRS_DIN_RTH_BG=1: the builder warp is the third gatherer
        b = lambda n, c: Bar(n, c)
        self.full = [b("full%d" % i, (96 if builder_gathers else 64) + 32) for i in range(SLOTS)]
        self.empty = [b("empty%d" % i, 1) for i in range(SLOTS)]
        self.d1_full, self.d1_free = b("d1_full", 1), b("d1_free", 128)
        self.w_ready = [b("w_ready%d" % i, 128) for i in range(2)]
        self.d2_full = [b("d2_full%d" % i, 1) for i in range(2)]
        self.pfull = [b("pfull%d" % i, 1) for i in range(3)]
        self.pfree = [b("pfree%d" % i, 1) for i in range(3)]
        self.cbar = b("cbar", 1)
        self.pipe = []            # tensor pipe: list of ("mma", reads, writes) / ("commit", bar)
        self.copies = []          # copy engine: (bar, bytes, buffer)
        self.sync_wait = {}       # __syncthreads rendezvous
        self.sync_gen = 0
        # hazard tracking: resources with outstanding MMA reads / writes
        self.busy = {}            # resource -> outstanding MMA count
        self.log = []

    # ---- primitives used by the actors (generators yield conditions) -------------------------
    def wait(self, bar, parity, intended):
        while not bar.passed(parity):
            yield
        assert bar.done == intended + 1, "%s: wait(parity %d) meant completion #%d, barrier has %d" % (
            bar.name, parity, intended, bar.done)

    def syncthreads(self, who):
        gen = self.sync_gen
        self.sync_wait.setdefault(gen, set()).add(who)
        while len(self.sync_wait[gen]) < 4:
            yield
        if self.sync_gen == gen:
            self.sync_gen += 1

    def issue_mma(self, reads, writes):
        for r in reads + writes:
            self.busy[r] = self.busy.get(r, 0) + 1
        self.pipe.append(("mma", reads, writes))

    def commit(self, bar):
        self.pipe.append(("commit", bar))

    def touch(self, res, what):
        assert self.busy.get(res, 0) == 0, "%s while an MMA still uses %s" % (what, res)

    # ---- actors ---------------------------------------------------------------------------------
    def gatherers(self):
        kbase = 0
        for n_tiles in self.groups:
            yield from self.syncthreads("g")                       # end of phase 0
            def gather(k):
                K = kbase + k
                slot = K % SLOTS
                if K >= SLOTS:
                    yield from self.wait(self.empty[slot], ((K // SLOTS) + 1) & 1, K // SLOTS - 1)
                self.touch("A%d" % slot, "cp.async into tile")
            for a in range(AHEAD):
                if a < n_tiles:
                    yield from gather(a)
            for k in range(n_tiles):
                self.full[(kbase + k) % SLOTS].arrive(64)
                yield
                if k + AHEAD < n_tiles:
                    yield from gather(k + AHEAD)
            kbase += n_tiles
            yield from self.syncthreads("g")                       # after phase 1
            yield from self.syncthreads("g")                       # X operand in place
            yield from self.wait(self.cbar, self.cph("g"), self.cidx("g"))
            yield from self.syncthreads("g")                       # H1 in place
            yield from self.wait(self.cbar, self.cph("g"), self.cidx("g"))
            for _ in range(4):
                yield from self.syncthreads("g")                   # red / zp / final / end of group

    # cbar phase bookkeeping per actor (cphase ^= 1 after every wait)
    def cph(self, who):
        self._c = getattr(self, "_c", {})
        v = self._c.get(who, 0)
        self._c[who] = v + 1
        self._last = getattr(self, "_last", {})
        self._last[who] = v
        return v & 1

    def cidx(self, who):
        return self._last[who]

    def builder(self):
        kbase = pbase = 0
        for n_tiles in self.groups:
            yield from self.syncthreads("b")

            def build_b(k):
                slot = (kbase + k) % SLOTS
                self.touch("B%d" % slot, "builder write")
                yield
                self.full[slot].arrive(32)

            def gather(k):
                K = kbase + k
                slot = K % SLOTS
                if K >= SLOTS:
                    yield from self.wait(self.empty[slot], ((K // SLOTS) + 1) & 1, K // SLOTS - 1)
                self.touch("A%d" % slot, "cp.async into tile (builder warp)")
            if not self.bg:
                for k in range(n_tiles):
                    K = kbase + k
                    slot = K % SLOTS
                    if K >= SLOTS:
                        yield from self.wait(self.empty[slot], ((K // SLOTS) + 1) & 1, K // SLOTS - 1)
                    yield from build_b(k)
            else:
                for a in range(AHEAD):
                    if a < n_tiles:
                        yield from gather(a)
                        yield from build_b(a)
                for k in range(n_tiles):
                    self.full[(kbase + k) % SLOTS].arrive(32)
                    yield
                    if k + AHEAD < n_tiles:
                        yield from gather(k + AHEAD)
                        yield from build_b(k + AHEAD)
            kbase += n_tiles
            yield from self.syncthreads("b")

            def stream(i0, i1):
                for i in range(i0, i1):
                    P = pbase + i
                    j = P % 3
                    if P >= 3:
                        yield from self.wait(self.pfree[j], ((P // 3) + 1) & 1, P // 3 - 1)
                    self.touch("W%d" % j, "bulk copy of piece %d" % i)
                    self.touch("ring_tiles", "bulk copy of piece %d" % i)
                    self.pfull[j].arrive(1)
                    self.pfull[j].expect(16384)
                    self.copies.append((self.pfull[j], 16384))
                    yield
            yield from stream(0, 3)
            yield from self.syncthreads("b")
            yield from stream(3, 8)
            pbase += 8
            yield from self.wait(self.cbar, self.cph("b"), self.cidx("b"))
            yield from self.syncthreads("b")
            yield from self.wait(self.cbar, self.cph("b"), self.cidx("b"))
            for _ in range(4):
                yield from self.syncthreads("b")

    def issuer(self):
        kbase = pbase = 0
        for n_tiles in self.groups:
            yield from self.syncthreads("i")

            def mma1(k):
                K = kbase + k
                slot = K % SLOTS
                yield from self.wait(self.full[slot], (K // SLOTS) & 1, K // SLOTS)
                if K >= 1:
                    yield from self.wait(self.d1_free, (K - 1) & 1, K - 1)
                self.touch("D1", "MMA1 overwrite of the accumulator")
                self.issue_mma(["A%d" % slot, "B%d" % slot, "ring_tiles"], ["D1"])
                self.commit(self.d1_full)
            if n_tiles > 0:
                yield from mma1(0)
            for k in range(n_tiles):
                K = kbase + k
                slot, u = K % SLOTS, K & 1
                if k + 1 < n_tiles:
                    yield from mma1(k + 1)
                yield from self.wait(self.w_ready[u], (K >> 1) & 1, K >> 1)
                self.issue_mma(["A%d" % slot, "b2s%d" % u, "ring_tiles"], ["D2_%d" % u])
                self.commit(self.d2_full[u])
                self.commit(self.empty[slot])
                yield
            kbase += n_tiles
            yield from self.syncthreads("i")
            yield from self.syncthreads("i")
            for i in range(6):
                P = pbase + i
                j = P % 3
                yield from self.wait(self.pfull[j], (P // 3) & 1, P // 3)
                self.issue_mma(["W%d" % j, "X"], ["TOP1"])
                self.commit(self.pfree[j])
                if i == 5:
                    self.commit(self.cbar)
                yield
            yield from self.wait(self.cbar, self.cph("i"), self.cidx("i"))
            self.touch("X", "layer-1 epilogue writes H1 over the X operand")
            yield from self.syncthreads("i")
            for i in range(6, 8):
                P = pbase + i
                j = P % 3
                yield from self.wait(self.pfull[j], (P // 3) & 1, P // 3)
                self.issue_mma(["W%d" % j, "X"], ["TOP2"])
                self.commit(self.pfree[j])
                if i == 7:
                    self.commit(self.cbar)
                yield
            pbase += 8
            yield from self.wait(self.cbar, self.cph("i"), self.cidx("i"))
            for _ in range(4):
                yield from self.syncthreads("i")

    def consumer(self):
        kbase = 0
        for n_tiles in self.groups:
            yield from self.syncthreads("c")

            def pool_out(k):
                K = kbase + k
                u = K & 1
                yield from self.wait(self.d2_full[u], (K >> 1) & 1, K >> 1)
                self.touch("D2_%d" % u, "pooled read-back")
            for k in range(n_tiles):
                K = kbase + k
                u = K & 1
                yield from self.wait(self.d1_full, K & 1, K)
                self.touch("D1", "gate read of the accumulator")
                yield
                self.d1_free.arrive(128)
                yield
                self.touch("b2s%d" % u, "pooling-weight write")
                self.w_ready[u].arrive(128)
                if k >= 1:
                    yield from pool_out(k - 1)
            if n_tiles > 0:
                yield from pool_out(n_tiles - 1)
            kbase += n_tiles
            yield from self.syncthreads("c")
            yield from self.syncthreads("c")
            yield from self.wait(self.cbar, self.cph("c"), self.cidx("c"))
            yield from self.syncthreads("c")
            yield from self.wait(self.cbar, self.cph("c"), self.cidx("c"))
            for _ in range(4):
                yield from self.syncthreads("c")

    # ---- asynchronous engines --------------------------------------------------------------------
    def step_engines(self):
        if self.pipe and self.rng.random() < 0.5:
            op = self.pipe.pop(0)
            if op[0] == "mma":
                for r in op[1] + op[2]:
                    self.busy[r] -= 1
            else:
                op[1].arrive(1)
        if self.copies and self.rng.random() < 0.5:
            bar, n = self.copies.pop(self.rng.randrange(len(self.copies)))
            bar.complete_tx(n)

    def run(self):
        actors = {"g": self.gatherers(), "b": self.builder(), "i": self.issuer(), "c": self.consumer()}
        idle = 0
        while actors:
            name = self.rng.choice(list(actors))
            before = (self.sync_gen, tuple(b.done for b in self.all_bars()), len(self.pipe), len(self.copies))
            try:
                next(actors[name])
            except StopIteration:
                del actors[name]
            self.step_engines()
            after = (self.sync_gen, tuple(b.done for b in self.all_bars()), len(self.pipe), len(self.copies))
            idle = idle + 1 if before == after else 0
            if idle > 20000:
                raise RuntimeError("deadlock; still running: %s; barriers: %s" % (
                    sorted(actors), {b.name: (b.done, b.pending, b.tx) for b in self.all_bars()}))
        while self.pipe or self.copies:
            self.step_engines()

    def all_bars(self):
        return self.full + self.empty + [self.d1_full, self.d1_free] + self.w_ready + self.d2_full + \
            self.pfull + self.pfree + [self.cbar]


def main():
    shapes = [[1], [2], [3], [4], [7], [14], [16], [1, 1, 1], [2, 3], [14, 14, 14], [16, 1, 5, 16], [3, 3, 3, 3, 3],
              [5, 4, 3, 2, 1, 2, 3]]
    runs = 0
    for shape in shapes:
        for seed in range(60):
            Sim(shape, seed).run()
            Sim(shape, seed, builder_gathers=True).run()
            runs += 2
    print("din_rth protocol model: %d runs over %d group shapes, no deadlock, no parity aliasing, "
          "no operand hazard" % (runs, len(shapes)))


if __name__ == "__main__":
    sys.exit(main())
