"""Model check of the W4A16 GEMM's barrier protocol (scalellm_b200/csrc/w4a16.cu) on the CPU.

The kernel's five roles (weight-blob producer, activation producer, dequant groups, MMA issuer with
its asynchronous tensor pipe, epilogue) are coroutines that talk through phase/parity mbarriers
exactly as the kernel does — same ring depths, same barrier indices and parities, same order of
waits, commits and arrivals — and a random scheduler interleaves them.  The model checks what the
hardware would silently get wrong: an MMA that reads a TMEM slot / activation stage holding another
tile, a producer that overwrites a slot a queued MMA still needs, an accumulator reused before
the epilogue drained it, and deadlock (which also catches a barrier that ran two phases ahead of a
waiter).  It restates the protocol by hand, so it guards the DESIGN of the default kernel and of
the B200_W4_VARIANT experiments, not their compiled code (the GPU parity tests do that).

  python tools/w4_protocol_sim.py            # all variants, many random partitions and schedules
"""
import random
import sys

RAW_STAGES, STAGES = 11, 6          # W4Cfg<MT<=64, NSUB=1>: weight ring; act ring == TMEM slot ring


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def done(self, parity):          # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, segments, var, rng):
        """segments: list of tile counts, one per accumulator segment of this CTA (NSUB = 1:
        a tile is one (n tile, k tile) unit); var: B200_W4_VARIANT bit mask."""
        self.var, self.rng = var, rng
        self.seg_of, self.first, self.last = [], [], []
        for s, n in enumerate(segments):
            for i in range(n):
                self.seg_of.append(s)
                self.first.append(i == 0)
                self.last.append(i == n - 1)
        self.total = len(self.seg_of)
        self.n_seg = len(segments)
        self.groups = 3 if var & 4 else 4
        self.raw_full = [Bar(1) for _ in range(RAW_STAGES)]
        self.raw_empty = [Bar(4) for _ in range(RAW_STAGES)]
        self.act_full = [Bar(1) for _ in range(STAGES)]
        self.act_empty = [Bar(1) for _ in range(STAGES)]
        # variant bit 8: the activation copy completes on the slot's deq_full barrier (4 dequant
        # arrivals + the producer's expect_tx arrival + the transaction itself)
        self.deq_full = [Bar(6 if var & 8 else 4) for _ in range(STAGES)]
        self.deq_empty = [Bar(1) for _ in range(STAGES)]
        self.tmem_full = [Bar(1), Bar(1)]
        self.tmem_empty = [Bar(4), Bar(4)]
        self.raw = [None] * RAW_STAGES       # tile id held by each ring entry
        self.act = [None] * STAGES
        self.slot = [None] * STAGES
        self.acc_seg = [None, None]          # segment accumulating in each buffer
        self.acc_tiles = [0, 0]
        self.drained = [True, True]
        self.pipe = []                       # in-order tensor pipe: ('mma', tile, slot, buf) / ('commit', bar)
        self.inflight = []                   # async copies: (kind, stage, tile)
        self.seg_done = 0

    # ---- roles (generators yield (bar, parity) to wait, None to give up the time slice) ----
    def slot_release_bar(self, cnt):
        if self.var & 2:
            pr = cnt >> 1
            return self.deq_empty[pr % 3], ((pr // 3) & 1) ^ 1
        return self.deq_empty[cnt % STAGES], ((cnt // STAGES) & 1) ^ 1

    def raw_producer(self):
        for cnt in range(self.total):
            rs, rph = cnt % RAW_STAGES, (cnt // RAW_STAGES) & 1
            yield self.raw_empty[rs], rph ^ 1
            self.inflight.append(("raw", rs, cnt))
            yield None

    def act_producer(self):
        for cnt in range(self.total):
            st, ph = cnt % STAGES, (cnt // STAGES) & 1
            if self.var & 3:
                yield self.slot_release_bar(cnt)
            else:
                yield self.act_empty[st], ph ^ 1
            assert not any(op[0] == "mma" and op[2] == st for op in self.pipe), "act stage overwritten under a queued MMA"
            if self.var & 8:
                self.deq_full[st].arrive()           # arrive.expect_tx
            self.inflight.append(("act", st, cnt))
            yield None

    def dequant(self, group):
        for cnt in range(self.total):
            if cnt % self.groups != group:
                continue
            rs, rph = cnt % RAW_STAGES, (cnt // RAW_STAGES) & 1
            st = cnt % STAGES
            yield self.raw_full[rs], rph
            assert self.raw[rs] == cnt, "dequant read a stale weight blob"
            yield None
            yield self.slot_release_bar(cnt)
            assert not any(op[0] == "mma" and op[2] == st for op in self.pipe), "TMEM slot overwritten under a queued MMA"
            self.slot[st] = cnt
            yield None
            for _ in range(4):               # the group's four warps arrive one by one
                self.deq_full[st].arrive()
                self.raw_empty[rs].arrive()

    def issue(self, cnt, seg):
        st = cnt % STAGES
        self.pipe.append(("mma", cnt, st, seg & 1))

    def mma_default(self):
        cnt = 0
        for seg in range(self.n_seg):
            buf, tph = seg & 1, (seg >> 1) & 1
            yield self.tmem_empty[buf], tph ^ 1
            while cnt < self.total and self.seg_of[cnt] == seg:
                st, ph = cnt % STAGES, (cnt // STAGES) & 1
                if not (self.var & 8):
                    yield self.act_full[st], ph
                yield self.deq_full[st], ph
                self.issue(cnt, seg)
                self.pipe.append(("commit", self.deq_empty[st]))
                if not (self.var & 3):
                    self.pipe.append(("commit", self.act_empty[st]))
                if self.last[cnt]:
                    self.pipe.append(("commit", self.tmem_full[buf]))
                cnt += 1
                yield None

    def mma_pairs(self):
        seg = -1
        for cnt in range(0, self.total, 2):
            n = min(2, self.total - cnt)
            for j in range(n):
                if self.first[cnt + j]:
                    seg += 1
                    assert self.seg_of[cnt + j] == seg
                    yield self.tmem_empty[seg & 1], ((seg >> 1) & 1) ^ 1
                st, ph = (cnt + j) % STAGES, ((cnt + j) // STAGES) & 1
                if not (self.var & 8):
                    yield self.act_full[st], ph
                yield self.deq_full[st], ph
            for j in range(n):
                s = self.seg_of[cnt + j]
                self.issue(cnt + j, s)
                if self.last[cnt + j]:
                    self.pipe.append(("commit", self.tmem_full[s & 1]))
            self.pipe.append(("commit", self.deq_empty[(cnt >> 1) % 3]))
            yield None

    def epilogue(self):
        for seg in range(self.n_seg):
            buf, tph = seg & 1, (seg >> 1) & 1
            yield self.tmem_full[buf], tph
            assert self.acc_seg[buf] == seg, "epilogue read another segment's accumulator"
            want = sum(1 for s in self.seg_of if s == seg)
            assert self.acc_tiles[buf] == want, "accumulator read before all its MMAs ran"
            yield None
            self.drained[buf] = True
            self.seg_done += 1
            for _ in range(4):
                self.tmem_empty[buf].arrive()

    # ---- asynchronous hardware ----
    def step_pipe(self):
        op = self.pipe.pop(0)
        if op[0] == "commit":
            op[1].arrive()
            return
        _, tile, st, buf = op
        assert self.slot[st] == tile, f"MMA of tile {tile} read TMEM slot {st} holding {self.slot[st]}"
        assert self.act[st] == tile, f"MMA of tile {tile} read activation stage {st} holding {self.act[st]}"
        if self.first[tile]:
            assert self.drained[buf], "accumulator overwritten before the epilogue drained it"
            self.acc_seg[buf], self.acc_tiles[buf], self.drained[buf] = self.seg_of[tile], 0, False
        assert self.acc_seg[buf] == self.seg_of[tile]
        self.acc_tiles[buf] += 1

    def step_copy(self):
        kind, st, tile = self.inflight.pop(self.rng.randrange(len(self.inflight)))
        if kind == "raw":
            self.raw[st] = tile
            self.raw_full[st].arrive()
        else:
            self.act[st] = tile
            (self.deq_full if self.var & 8 else self.act_full)[st].arrive()

    def run(self):
        roles = [self.raw_producer(), self.act_producer(), self.epilogue(),
                 self.mma_pairs() if self.var & 2 else self.mma_default()]
        roles += [self.dequant(g) for g in range(4)]
        waiting = [None] * len(roles)        # (bar, parity) a role is blocked on
        alive = [True] * len(roles)
        steps = 0
        while True:
            choices = [("role", i) for i in range(len(roles))
                       if alive[i] and (waiting[i] is None or waiting[i][0].done(waiting[i][1]))]
            if self.pipe:
                choices.append(("pipe", 0))
            if self.inflight:
                choices.append(("copy", 0))
            if not choices:
                break
            kind, i = self.rng.choice(choices)
            steps += 1
            assert steps < 200000, "livelock"
            if kind == "pipe":
                self.step_pipe()
            elif kind == "copy":
                self.step_copy()
            else:
                try:
                    waiting[i] = next(roles[i])
                except StopIteration:
                    alive[i] = False
        assert not any(alive), f"deadlock: roles {[i for i, a in enumerate(alive) if a]} blocked"
        assert self.seg_done == self.n_seg


def random_segments(rng):
    """A CTA's share of the stream-K unit list: a tail of one n tile, whole n tiles, a head."""
    kt = rng.choice([1, 2, 3, 8, 32, 112])
    share = rng.randint(1, 3 * kt + 5)
    start = rng.randrange(kt)
    segs, left, pos = [], share, start
    while left > 0:
        n = min(kt - pos, left)
        segs.append(n)
        left -= n
        pos = 0
    return segs


def check(variants=(0, 1, 2, 4, 6, 10, 14), trials=300, seed=0):
    rng = random.Random(seed)
    for var in variants:
        for _ in range(trials):
            Sim(random_segments(rng), var, rng).run()
    return True


if __name__ == "__main__":
    check(trials=int(sys.argv[1]) if len(sys.argv) > 1 else 1000)
    print("protocol model: no hazard, no deadlock (variants 0 1 2 4 6 10 14)")
