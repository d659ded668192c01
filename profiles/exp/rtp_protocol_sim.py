"""Discrete-event model of din_rtp_kernel's mbarrier protocol (csrc/din_rtp.cu), run on the CPU.

Every wait / arrive / commit of the kernel is transcribed - parity expressions copied from the
source - into cooperating actors (gatherers, builders, issuer, loader, two consumers, the top-MLP
warpgroup, the tensor pipe that retires commits in issue order) and run under random
interleavings.  Checked:

  * no deadlock: every actor finishes every group;
  * no parity aliasing: when a wait passes, the barrier has completed exactly the phase the code
    meant (the intended completion index is stated next to each wait);
  * hazards: a ring slot / weight operand / accumulator / staging buffer / pooled buffer is never
    rewritten while an MMA or a reader that uses it is outstanding.

    python profiles/exp/rtp_protocol_sim.py          # many shapes x random schedules
"""
import random
import sys

NSA, NSB, AHEAD = 6, 3, 2


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.done = name, count, count, 0

    def arrive(self, n=1):
        self.pending -= n
        assert self.pending >= 0, "over-arrival on %s" % self.name
        if self.pending == 0:
            self.done += 1
            self.pending = self.count

    def passed(self, parity):
        return (self.done & 1) != parity


class Sim:
    def __init__(self, groups, seed):
        self.rng = random.Random(seed)
        self.groups = groups                      # tiles per group of this CTA
        b = Bar
        self.a_full = [b("a_full%d" % i, 64) for i in range(NSA)]
        self.a_empty = [b("a_empty%d" % i, 1) for i in range(NSA)]
        self.b_full = [b("b_full%d" % i, 128) for i in range(NSB)]
        self.b_empty = [b("b_empty%d" % i, 1) for i in range(NSB)]
        self.d1_full = [b("d1_full%d" % i, 1) for i in range(2)]
        self.d1_free = [b("d1_free%d" % i, 128) for i in range(2)]
        self.w_ready = [[b("w_ready%d%d" % (q, u), 128) for u in range(2)] for q in range(2)]
        self.d2_full = [[b("d2_full%d%d" % (q, u), 1) for u in range(2)] for q in range(2)]
        self.staged = [b("staged%d" % i, 32) for i in range(2)]
        self.stage_free = [b("stage_free%d" % i, 640) for i in range(2)]
        self.pooled_ready = [b("pooled_ready%d" % i, 256) for i in range(2)]
        self.pooled_free = [b("pooled_free%d" % i, 128) for i in range(2)]
        self.cbar = b("cbar", 1)
        self.pipe = []                # tensor pipe: ("mma", resources) / ("commit", bar)
        self.busy = {}                # resource -> outstanding MMA count
        self.readers = {}             # resource -> set of roles still reading generation g
        self.stage_gen = [0, 1]       # group whose inputs staging buffer s holds (0 and 1 staged by the prologue)
        self.pooled_gen = [None, None]
        self.finished = set()

    # ---- primitives -------------------------------------------------------------------------------
    def wait(self, bar, parity, intended):
        while not bar.passed(parity):
            yield
        assert bar.done == intended + 1, "%s: wait(parity %d) meant completion #%d, barrier has %d" % (
            bar.name, parity, intended, bar.done)

    def issue_mma(self, res):
        for r in res:
            self.busy[r] = self.busy.get(r, 0) + 1
        self.pipe.append(("mma", res))

    def commit(self, bar):
        self.pipe.append(("commit", bar))

    def touch(self, res, what):
        assert self.busy.get(res, 0) == 0, "%s while an MMA still uses %s" % (what, res)

    def staged_wait(self, j):
        if j >= 2:
            yield from self.wait(self.staged[j & 1], ((j >> 1) - 1) & 1, (j >> 1) - 1)
        assert self.stage_gen[j & 1] == j, "staging buffer %d holds group %s, reader wants %d" % (
            j & 1, self.stage_gen[j & 1], j)

    # ---- actors -------------------------------------------------------------------------------------
    def gatherers(self, team):                   # team 0: even tiles (128 threads), team 1: odd tiles (96)
        n_thr = 64
        Kg = Dg = 0                              # own tiles issued / delivered
        slot_of = lambda n: (2 * n + team) % NSA
        kbase = 0
        for j, n_tiles in enumerate(self.groups):
            if j >= 2 and not self.staged[j & 1].passed(((j >> 1) - 1) & 1):
                while Dg < Kg:                                 # the loader is late: deliver what has landed first
                    self.a_full[slot_of(Dg)].arrive(n_thr)
                    Dg += 1
                    yield
            yield from self.staged_wait(j)
            for k in range((team - kbase) & 1, n_tiles, 2):
                K = kbase + k
                slot = K % NSA
                assert slot == slot_of(Kg)
                if K >= NSA:
                    par = ((K // NSA) + 1) & 1
                    while not self.a_empty[slot].passed(par):          # publish landed tiles before blocking
                        if Dg < Kg:
                            self.a_full[slot_of(Dg)].arrive(n_thr)
                            Dg += 1
                        yield
                    yield from self.wait(self.a_empty[slot], par, K // NSA - 1)
                self.touch("A%d" % slot, "cp.async into the tile")
                Kg += 1
                if Kg - Dg > AHEAD:
                    self.a_full[slot_of(Dg)].arrive(n_thr)
                    Dg += 1
                yield
            self.stage_free[j & 1].arrive(n_thr)
            kbase += n_tiles
            yield
        while Dg < Kg:
            self.a_full[slot_of(Dg)].arrive(n_thr)
            Dg += 1
            yield
        self.finished.add("g%d" % team)

    def builders(self):
        Kb = 0
        for j, n_tiles in enumerate(self.groups):
            yield from self.staged_wait(j)
            for k in range(n_tiles):
                slot = Kb % NSB
                if Kb >= NSB:
                    yield from self.wait(self.b_empty[slot], ((Kb // NSB) + 1) & 1, Kb // NSB - 1)
                self.touch("B%d" % slot, "builder writes W_r")
                assert self.stage_gen[j & 1] == j
                self.b_full[slot].arrive(128)
                Kb += 1
                yield
            self.stage_free[j & 1].arrive(128)
            yield
        self.finished.add("b")

    def issuer_mma1(self):                       # warp 4
        NT = sum(self.groups)
        sa = sb = pa = pb = 0
        for K in range(NT):
            q = K & 1
            yield from self.wait(self.a_full[sa], pa, K // NSA)
            yield from self.wait(self.b_full[sb], pb, K // NSB)
            if K >= 2:
                yield from self.wait(self.d1_free[q], ((K >> 1) - 1) & 1, (K >> 1) - 1)
            self.touch("D1_%d" % q, "MMA1 overwrites the accumulators")
            self.issue_mma(["A%d" % sa, "B%d" % sb, "D1_%d" % q])
            self.commit(self.d1_full[q])
            self.commit(self.b_empty[sb])
            yield
            sa += 1
            if sa == NSA:
                sa, pa = 0, pa ^ 1
            sb += 1
            if sb == NSB:
                sb, pb = 0, pb ^ 1
        self.finished.add("i1")

    def issuer_pool(self):                       # warp 5
        NT = sum(self.groups)
        sa = 0
        for K in range(NT):
            q, u = K & 1, (K >> 1) & 1
            yield from self.wait(self.w_ready[q][u], (K >> 2) & 1, K >> 2)
            self.touch("D2_%d%d" % (q, u), "pooling MMA overwrites accumulators")
            self.issue_mma(["A%d" % sa, "W%d%d" % (q, u), "D2_%d%d" % (q, u)])
            self.commit(self.d2_full[q][u])
            self.commit(self.a_empty[sa])
            yield
            sa = (sa + 1) % NSA
        self.finished.add("i2")

    def loader(self):
        for j in range(2, len(self.groups)):
            yield from self.wait(self.stage_free[j & 1], ((j >> 1) - 1) & 1, (j >> 1) - 1)
            self.stage_gen[j & 1] = j                  # overwrites the inputs of group j - 2
            yield
            self.staged[j & 1].arrive(32)
            yield
        self.finished.add("l")

    def consumer(self, q):
        pend = None
        pool_group = -1
        d1_reading = [False]

        def pooled_buffer_wait(j):
            nonlocal pool_group
            if pool_group != j:
                if j >= 2:
                    yield from self.wait(self.pooled_free[j & 1], ((j >> 1) - 1) & 1, (j >> 1) - 1)
                pool_group = j
                if self.pooled_gen[j & 1] != j:
                    self.pooled_gen[j & 1] = j

        def pool_out():
            nonlocal pend
            K, j, k, last = pend
            u = (K >> 1) & 1
            yield from self.wait(self.d2_full[q][u], (K >> 2) & 1, K >> 2)
            yield from pooled_buffer_wait(j)
            assert self.pooled_gen[j & 1] == j, "pooled buffer %d holds group %s" % (j & 1, self.pooled_gen[j & 1])
            self.busy["D2r_%d%d" % (q, u)] = 0
            if last:
                self.pooled_ready[j & 1].arrive(128)
            pend = None
            yield
        kbase = 0
        for j, n_tiles in enumerate(self.groups):
            yield from self.staged_wait(j)
            any_tile = False
            for k in range((q - kbase) & 1, n_tiles, 2):
                any_tile = True
                K = kbase + k
                u = (K >> 1) & 1
                assert self.stage_gen[j & 1] == j            # cst from the candidate rows
                yield
                yield from self.wait(self.d1_full[q], (K >> 1) & 1, K >> 1)
                yield
                self.d1_free[q].arrive(128)                  # accumulators in registers
                yield
                self.touch("W%d%d" % (q, u), "consumer writes pooling weights")
                self.w_ready[q][u].arrive(128)
                yield
                if pend:
                    yield from pool_out()
                pend = (K, j, k, k + 2 >= n_tiles)
            if not any_tile:
                if pend:
                    yield from pool_out()
                yield from pooled_buffer_wait(j)
                self.pooled_ready[j & 1].arrive(128)
            self.stage_free[j & 1].arrive(128)
            kbase += n_tiles
            yield
        if pend:
            yield from pool_out()
        self.finished.add("c%d" % q)

    def top(self):
        cph = 0
        cidx = 0
        for j, n_tiles in enumerate(self.groups):
            s = j & 1
            yield from self.staged_wait(j)
            yield from self.wait(self.pooled_ready[s], (j >> 1) & 1, j >> 1)
            assert self.pooled_gen[s] == j, "top MLP of group %d reads pooled rows of group %s" % (j, self.pooled_gen[s])
            assert self.stage_gen[s] == j
            self.touch("X", "X operand rebuilt")
            self.pooled_free[s].arrive(128)
            yield
            self.touch("TOP", "layer-1 MMA overwrites accumulators")
            self.issue_mma(["X", "TOP"])
            self.commit(self.cbar)
            yield from self.wait(self.cbar, cph, cidx)
            cph ^= 1
            cidx += 1
            assert self.stage_gen[s] == j                     # numerics / genre ids in the epilogue
            self.stage_free[s].arrive(128)
            self.touch("X", "H1 written over X")
            yield
            self.issue_mma(["X", "TOP"])
            self.commit(self.cbar)
            yield from self.wait(self.cbar, cph, cidx)
            cph ^= 1
            cidx += 1
            yield
        self.finished.add("t")

    def tensor_pipe(self):
        while True:
            if self.pipe and self.rng.random() < 0.6:
                op = self.pipe.pop(0)
                if op[0] == "mma":
                    for r in op[1]:
                        self.busy[r] -= 1
                else:
                    op[1].arrive(1)
            yield

    # ---- scheduler --------------------------------------------------------------------------------
    def run(self, max_steps=400000):
        actors = {"g0": self.gatherers(0), "g1": self.gatherers(1), "b": self.builders(), "i1": self.issuer_mma1(), "i2": self.issuer_pool(), "l": self.loader(),
                  "c0": self.consumer(0), "c1": self.consumer(1), "t": self.top()}
        pipe = self.tensor_pipe()
        names = list(actors)
        steps = 0
        while actors:
            steps += 1
            assert steps < max_steps, "deadlock / livelock: still running %s; pipe %d" % (sorted(actors), len(self.pipe))
            if self.rng.random() < 0.3:
                next(pipe)
                continue
            n = self.rng.choice(names)
            if n not in actors:
                continue
            try:
                next(actors[n])
            except StopIteration:
                del actors[n]
        # drain the pipe
        while self.pipe:
            next(pipe)
        return steps


def main():
    rng = random.Random(1)
    shapes = [[14], [14, 14], [14, 14, 14, 14], [1], [1, 1, 1, 1, 1], [2, 1, 3, 1, 1, 2], [16] * 5, [7, 14, 3, 16, 1, 2, 5],
              [3] * 9, [16, 1, 16, 1, 16], [5, 5, 5, 5, 5, 5, 5, 5]]
    for _ in range(30):
        shapes.append([rng.randint(1, 16) for _ in range(rng.randint(1, 8))])
    runs = 0
    for shape in shapes:
        for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
            Sim(shape, seed).run()
            runs += 1
    print("rtp protocol model: %d runs over %d shapes, no deadlock / aliasing / hazard" % (runs, len(shapes)))


if __name__ == "__main__":
    main()
