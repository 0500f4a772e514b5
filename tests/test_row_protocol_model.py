"""Host-side model of the mbarrier protocol of conv3x3_row_kernel (esrganplus_b200/csrc/conv3x3_row.cuh): TMA producer,
two MMA-issuer warps, three epilogue warpgroups, the tensor pipe and TMA as asynchronous agents, run under a randomised
scheduler that also FREEZES single agents for long stretches (a context switch, a driver stall).  No GPU involved.

What it checks, for every issuer protocol the kernel has (row_alt 0 = both warps split the taps of every row, 2 = the
warps alternate rows and the idle warp adds a third arrival on the block barrier) and for the first version of the
alternating protocol (1, no third arrival — kept here to show the hole it had):
  * liveness: every agent finishes (a waiter that missed a phase of a parity barrier waits for ever: "lapping");
  * safety:   a row buffer is only refilled after every MMA that reads it completed; a block is only read by the
              epilogue after the MMAs of its three input rows completed; MMAs only accumulate into a released block.
mbarrier semantics modelled: arrival count per phase, try_wait.parity(P) succeeds iff the phase currently in progress
has parity != P (so a waiter that is two phases late sees "not complete" again)."""
import random

import pytest


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def passed(self, parity):
        return (self.phase & 1) != parity


class Model:
    def __init__(self, mode, nblk, stages, segments, rng, chunks=1):
        """chunks > 1: one "landed" barrier per K-chunk tile of a row buffer (ConvKParams::chunk_bars, row-alternating modes
        with an even ring depth): both warps meet the first chunk's barrier of every row, the issuing thread the others."""
        self.mode, self.nblk, self.D, self.segs, self.rng = mode, nblk, stages, segments, rng
        self.nch = chunks
        self.full = [Bar(1) for _ in range(stages * chunks)]
        self.tile_row = [None] * (stages * chunks)   # global row whose chunk has LANDED in a tile
        self.blk_full = [Bar(3 if mode == 2 else 2) for _ in range(nblk)]
        self.blk_empty = [Bar(1) for _ in range(nblk)]      # the 4 warps of a warpgroup modelled as one arrival
        self.tok = [Bar(1), Bar(1)]
        self.queues = [[], []]            # per issuer thread: in-order async ops ('mma', row, blocks) / ('commit', bar)
        self.tma = []                     # loads in flight: (buffer, global row)
        self.mma_done = set()             # (thread, global row) whose MMAs completed
        self.rows_total = sum(segments)
        self.buf_row = [None] * stages    # global row currently held by a buffer
        self.blk_owner = [None] * nblk    # output sequence number occupying a block (None = released)
        self.issued_by = {}               # global row -> set of threads that issued MMAs for it
        self.errors = []

    def pos(self, o):
        return (self.nblk - 1) - (o % self.nblk)

    def use(self, o):
        return (o // self.nblk) & 1

    # ---- agents (generators yielding a predicate to wait for, or None to be rescheduled) ----
    def producer(self):
        buf_bar, buf_par = [0] * self.D, [0] * self.D
        I, O0, b = 0, 0, 0
        for ni in self.segs:
            for k in range(ni):
                if I >= self.D:
                    bb, pp = self.blk_full[buf_bar[b]], buf_par[b]
                    yield lambda bb=bb, pp=pp: bb.passed(pp)
                    old = self.buf_row[b]
                    for t in self.issued_by.get(old, ()):   # every MMA that reads the buffer must be complete
                        if (t, old) not in self.mma_done:
                            self.errors.append(f"buffer {b} refilled while MMAs of row {old} (thread {t}) are in flight")
                    if len(self.issued_by.get(old, ())) == 0:
                        self.errors.append(f"buffer {b} refilled before row {old} was issued")
                oc = O0 + k
                buf_bar[b], buf_par[b] = self.pos(oc), self.use(oc)
                self.buf_row[b] = I
                for c in range(self.nch):
                    self.tma.append((b * self.nch + c, I))
                yield None
                I, b = I + 1, (b + 1) % self.D
            O0 += ni + 2

    def issuer(self, mw):
        b, fph, tph, O0, I = 0, 0, 0, 0, 0
        for ni in self.segs:
            for k in range(ni):
                fb = self.full[b * self.nch]
                yield lambda fb=fb, fph=fph: fb.passed(fph)
                news = [O0 + k + 2] + ([O0, O0 + 1] if k == 0 else [])
                for o in news:
                    be, par = self.blk_empty[self.pos(o)], self.use(o) ^ 1
                    yield lambda be=be, par=par: be.passed(par)
                blocks = [O0 + k, O0 + k + 1, O0 + k + 2]
                last = k == ni - 1
                mine = self.mode == 0 or (I & 1) == mw
                if mine:
                    tk, par = self.tok[mw], tph ^ (1 if mw == 0 else 0)
                    yield lambda tk=tk, par=par: tk.passed(par)
                    for o in news:                                  # first touch: the block must be free, then it is owned
                        if mw == 0 or self.mode != 0:
                            if self.blk_owner[self.pos(o)] is not None:
                                self.errors.append(f"MMA into block {self.pos(o)} still owned by output {self.blk_owner[self.pos(o)]}")
                            self.blk_owner[self.pos(o)] = o
                    if self.buf_row[b] != I:
                        self.errors.append(f"thread {mw} issues row {I} from buffer {b} holding row {self.buf_row[b]}")
                    for c in range(self.nch):      # the issuing thread meets the later chunks' barriers inside its issue loop
                        if c > 0:
                            fc = self.full[b * self.nch + c]
                            yield lambda fc=fc, fph=fph: fc.passed(fph)
                        if self.tile_row[b * self.nch + c] != I:
                            self.errors.append(f"thread {mw} issues chunk {c} of row {I} from a tile holding row {self.tile_row[b * self.nch + c]}")
                    self.issued_by.setdefault(I, set()).add(mw)
                    q = self.queues[mw]
                    q.append(("mma", I))
                    self.tok[mw ^ 1].arrive()
                    if self.mode == 0:
                        q.append(("commit", self.blk_full[self.pos(O0 + k)]))
                        if last:
                            q.append(("commit", self.blk_full[self.pos(O0 + k + 1)]))
                            q.append(("commit", self.blk_full[self.pos(O0 + k + 2)]))
                    else:
                        q.append(("commit", self.blk_full[self.pos(O0 + k)]))
                        q.append(("commit", self.blk_full[self.pos(O0 + k + 1)]))
                        if k == 0:
                            q.append(("commit", self.blk_full[self.pos(O0)]))
                        if last:
                            q.append(("commit", self.blk_full[self.pos(O0 + k + 1)]))
                            q.append(("commit", self.blk_full[self.pos(O0 + k + 2)]))
                            q.append(("commit", self.blk_full[self.pos(O0 + k + 2)]))
                    tph ^= 1
                elif self.mode == 2:
                    self.blk_full[self.pos(O0 + k)].arrive()
                    if last:
                        # a COMMIT: the last output row's first contributor (row k-1) was issued by this thread
                        self.queues[mw].append(("commit", self.blk_full[self.pos(O0 + k + 1)]))
                        self.blk_full[self.pos(O0 + k + 2)].arrive()
                yield None
                I += 1
                b += 1
                if b == self.D:
                    b, fph = 0, fph ^ 1
            O0 += ni + 2

    def epilogue(self, wg):
        O0, turn, row0 = 0, 0, 0
        for ni in self.segs:
            for j in range(ni + 2):
                mine = turn == wg
                turn = (turn + 1) % 3
                if not mine:
                    continue
                o = O0 + j
                bf, par = self.blk_full[self.pos(o)], self.use(o)
                yield lambda bf=bf, par=par: bf.passed(par)
                for k in (j - 2, j - 1, j):                      # input rows (within the segment) feeding output j
                    if 0 <= k < ni:
                        g = row0 + k
                        ts = self.issued_by.get(g, set())
                        need = {0, 1} if self.mode == 0 else {g & 1}
                        if ts != need or any((t, g) not in self.mma_done for t in ts):
                            self.errors.append(f"epilogue reads output {o} before row {g} completed (issued by {ts})")
                if self.blk_owner[self.pos(o)] != o:
                    self.errors.append(f"epilogue of output {o} finds block owned by {self.blk_owner[self.pos(o)]}")
                self.blk_owner[self.pos(o)] = None
                yield None
                self.blk_empty[self.pos(o)].arrive()
            O0 += ni + 2
            row0 += ni

    def hw_step(self):
        """One asynchronous hardware event (a TMA load lands, or one queued tensor-pipe op completes); False if none."""
        choices = []
        if self.tma:
            choices.append("tma")
        for t in (0, 1):
            if self.queues[t]:
                choices.append(t)
        if not choices:
            return False
        c = self.rng.choice(choices)
        if c == "tma":
            t, row = self.tma.pop(self.rng.randrange(len(self.tma)))   # tiles land in any order
            self.tile_row[t] = row
            self.full[t].arrive()
        else:
            op = self.queues[c].pop(0)                            # per-thread order; no order across threads assumed
            if op[0] == "mma":
                self.mma_done.add((c, op[1]))
            else:
                op[1].arrive()
        return True


def run(mode, nblk, stages, segments, seed, freeze=True, chunks=1):
    rng = random.Random(seed)
    m = Model(mode, nblk, stages, segments, rng, chunks)
    agents = {"prod": m.producer(), "mma0": m.issuer(0), "mma1": m.issuer(1),
              "epi0": m.epilogue(0), "epi1": m.epilogue(1), "epi2": m.epilogue(2)}
    waiting = {k: None for k in agents}
    frozen, frozen_for = None, 0
    steps = 0
    while agents:
        steps += 1
        assert steps < 2_000_000, "model did not terminate"
        if freeze and frozen_for == 0 and rng.random() < 0.01:
            frozen, frozen_for = rng.choice(list(agents)), rng.randrange(50, 400)
        if frozen_for:
            frozen_for -= 1
        runnable = [k for k in agents if (waiting[k] is None or waiting[k]()) and not (frozen_for and k == frozen)]
        if rng.random() < 0.5 or not runnable:
            if m.hw_step():
                continue
            if not runnable:
                if frozen_for:                                    # only the frozen agent can move: thaw it
                    frozen_for = 0
                    continue
                return m, "deadlock: " + ", ".join(sorted(agents))
        if not runnable:
            continue
        k = rng.choice(runnable)
        try:
            waiting[k] = next(agents[k])
        except StopIteration:
            del agents[k]
            waiting.pop(k)
    while m.hw_step():
        pass
    return m, None


CONFIGS = [  # (TMEM blocks, row buffers, input rows per segment of one CTA)
    (16, 2, [16]), (16, 2, [28]), (16, 8, [16]), (8, 2, [15, 3]), (8, 6, [9, 7]), (16, 4, [1, 14, 2]), (8, 3, [2, 2, 2, 9]),
    (8, 2, [1]), (16, 2, [1, 1, 5]), (16, 14, [30]),
    # images of a few rows put many segments (each with two extra blocks) behind one another: the planner then keeps two
    # row buffers (plan_row.inl: h <= blocks - 3), otherwise the producer could be lapped on a block barrier
    (8, 2, [3, 1, 2, 1, 12]), (16, 2, [1] * 12), (8, 2, [2] * 9),
    # ring sizes of the shadow-block layout (conv3x3_row.cuh: NBLK = 14, 7 with the conv1x1; positions are O % NBLK)
    (14, 2, [16]), (14, 2, [30]), (14, 4, [16]), (14, 12, [30]), (14, 2, [1, 14, 2]), (14, 2, [1] * 12), (14, 6, [15, 3]),
    (7, 2, [16]), (7, 5, [16]), (7, 5, [9, 7]), (7, 3, [2, 2, 2, 9]), (7, 2, [1]), (7, 2, [2] * 9), (7, 2, [3, 1, 2, 1, 12]),
]


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("nblk,stages,segments", CONFIGS)
def test_protocol_is_live_and_safe_under_arbitrary_delays(mode, nblk, stages, segments):
    for seed in range(150):
        m, dead = run(mode, nblk, stages, segments, seed)
        assert dead is None, (mode, nblk, stages, segments, seed, dead)
        assert not m.errors, (mode, seed, m.errors[:3])


def test_first_alternating_protocol_could_miss_a_phase():
    """row_alt == 1 (no arrival from the idle warp): with two row buffers a delayed issuer can find full_bar two phases
    ahead and wait for ever — the hole DESIGN.md section 5 describes.  The model must be able to find it, otherwise the
    liveness test above proves nothing."""
    found = False
    for seed in range(400):
        _, dead = run(1, 16, 2, [28], seed)
        if dead:
            found = True
            break
    assert found


def _segments_of_cta(n, h, x_tiles, grid, cta):
    """Input rows per segment of one CTA, following SegWalk / the producer loop of conv3x3_row.cuh."""
    units = n * x_tiles * h
    u, u_end = units * cta // grid, units * (cta + 1) // grid
    segs = []
    while u < u_end:
        col = u // h
        ya = u - col * h
        cnt = min(u_end - u, h - ya)
        yb = ya + cnt
        segs.append(min(yb, h - 1) - max(ya - 1, 0) + 1)
        u += cnt
    return segs


def test_protocol_on_random_shapes_with_the_planner_rules():
    """Random batch / height / column-block counts, the CTA's segments as SegWalk cuts them, the number of row buffers as
    plan_row.inl bounds it (<= blocks - 2, <= 8, two for images of a few rows): both issuer protocols stay live and safe."""
    rng = random.Random(2024)
    for case in range(250):
        nblk = rng.choice([7, 8, 14])
        n, h, x_tiles = rng.randrange(1, 600), rng.choice([1, 2, 3, 4, 5, 6, 7, 9, 14, 33, 128]), rng.randrange(1, 3)
        units = n * x_tiles * h
        grid = min(units, rng.choice([148, 74]))
        segs = _segments_of_cta(n, h, x_tiles, grid, rng.randrange(grid))
        stages = rng.randrange(2, min(8, nblk - 2) + 1)
        if h <= nblk - 3:
            stages = 2
        for mode in (0, 2):
            m, dead = run(mode, nblk, stages, segs, seed=case)
            assert dead is None and not m.errors, (mode, nblk, stages, n, h, x_tiles, grid, segs, dead, m.errors[:2])


@pytest.mark.parametrize("nblk,stages,chunks,segments", [(16, 2, 3, [28]), (16, 2, 3, [1, 14, 2]), (16, 4, 2, [16]), (8, 4, 2, [15, 3]),
                                                         (16, 2, 2, [2, 2, 2, 9]), (8, 2, 3, [1]), (16, 2, 3, [30, 1, 1]),
                                                         (16, 3, 2, [40]), (8, 3, 2, [7, 9]),
                                                         (14, 2, 3, [30]), (14, 4, 2, [16]), (14, 2, 3, [1, 14, 2]), (7, 4, 2, [15, 3]),
                                                         (14, 2, 2, [2, 2, 2, 9]), (7, 2, 3, [1])])
def test_chunk_barriers_are_live_and_safe(nblk, stages, chunks, segments):
    """ConvKParams::chunk_bars (row-alternating issuers, stages * chunks <= 8): one landed-barrier per K-chunk tile, tiles
    landing in any order; a chunk's MMAs must only ever be issued from a tile that holds that row's chunk, and nobody may be
    lapped on a chunk barrier.  The later chunks' barriers are polled by the issuing warp only; with an odd ring depth a
    buffer alternates between the warps and each polls them every second phase — still safe, because the phase in between
    was met by the other warp earlier in turn order (the planner nevertheless keeps to even depths)."""
    for seed in range(120):
        m, dead = run(2, nblk, stages, segments, seed, chunks=chunks)
        assert dead is None, (nblk, stages, chunks, segments, seed, dead)
        assert not m.errors, (seed, m.errors[:3])


# ---------------------------------------------------------------------------------------------------------------------
# CTA pairs (conv3x3_row_kernel<..., PAIR = true>, tcgen05 cta_group::2): the same protocol over two CTAs.
# ---------------------------------------------------------------------------------------------------------------------
class PairModel(Model):
    """Leader (CTA 0) and peer (CTA 1) stream the same row sequence.  The issuers live in the leader only and run the
    row-alternating protocol (mode 2).  Barriers the issuers wait on live in the leader: `full` collects the leader
    producer's arrival (expect_tx of both tiles) plus the landing of BOTH CTAs' tiles; `blk_empty` one arrival per CTA (a
    warpgroup of each).  Every commit is multicast: it arrives on `blk_full` of both CTAs, as does the idle warp's plain
    arrival; each CTA's producer and epilogues wait on their own copy.  Besides liveness / safety in both CTAs the model
    checks that NOTHING asynchronous is in flight when the last agent finishes — the moment both CTAs pass the final cluster
    barrier and leave (a multicast arrival after that would hit a CTA that is gone)."""

    def __init__(self, nblk, stages, segments, rng, single=False):
        """single: ESRP_PAIR_SINGLE — one issuer thread issues every row and commits once per row; block barriers count 1."""
        super().__init__(2, nblk, stages, segments, rng, 1)
        self.single = single
        self.full = [Bar(3) for _ in range(stages)]                 # leader producer + tile of CTA 0 + tile of CTA 1
        if single:
            self.blk_full = [Bar(1) for _ in range(nblk)]
        self.blk_full2 = [self.blk_full, [Bar(1 if single else 3) for _ in range(nblk)]]
        self.blk_empty = [Bar(2) for _ in range(nblk)]
        self.buf_row2 = [[None] * stages, [None] * stages]
        self.tile_row2 = [[None] * stages, [None] * stages]
        self.blk_owner2 = [[None] * nblk, [None] * nblk]

    def producer2(self, cta):
        buf_bar, buf_par = [0] * self.D, [0] * self.D
        I, O0, b = 0, 0, 0
        for ni in self.segs:
            for k in range(ni):
                if I >= self.D:
                    bb, pp = self.blk_full2[cta][buf_bar[b]], buf_par[b]
                    yield lambda bb=bb, pp=pp: bb.passed(pp)
                    old = self.buf_row2[cta][b]
                    ts = self.issued_by.get(old, ())
                    if not ts or any((t, old) not in self.mma_done for t in ts):
                        self.errors.append(f"CTA {cta}: buffer {b} refilled while row {old} is not complete")
                oc = O0 + k
                buf_bar[b], buf_par[b] = self.pos(oc), self.use(oc)
                self.buf_row2[cta][b] = I
                if cta == 0:
                    self.full[b].arrive()                            # arrive.expect_tx(bytes of both tiles)
                self.tma.append((cta, b, I))
                yield None
                I, b = I + 1, (b + 1) % self.D
            O0 += ni + 2

    def issuer(self, mw):
        b, fph, tph, O0, I = 0, 0, 0, 0, 0
        both = lambda bars, o: [bars[0][self.pos(o)], bars[1][self.pos(o)]]
        for ni in self.segs:
            for k in range(ni):
                fb = self.full[b]
                yield lambda fb=fb, fph=fph: fb.passed(fph)
                news = [O0 + k + 2] + ([O0, O0 + 1] if k == 0 else [])
                for o in news:
                    be, par = self.blk_empty[self.pos(o)], self.use(o) ^ 1
                    yield lambda be=be, par=par: be.passed(par)
                last = k == ni - 1
                if self.single:
                    for o in news:
                        for c in (0, 1):
                            if self.blk_owner2[c][self.pos(o)] is not None:
                                self.errors.append(f"CTA {c}: MMA into block {self.pos(o)} still owned by output {self.blk_owner2[c][self.pos(o)]}")
                            self.blk_owner2[c][self.pos(o)] = o
                    for c in (0, 1):
                        if self.buf_row2[c][b] != I or self.tile_row2[c][b] != I:
                            self.errors.append(f"row {I} issued from buffer {b} of CTA {c} holding row {self.buf_row2[c][b]} / landed {self.tile_row2[c][b]}")
                    self.issued_by.setdefault(I, set()).add(0)
                    q = self.queues[0]
                    q.append(("mma", I))
                    for o in [O0 + k] + ([O0 + k + 1, O0 + k + 2] if last else []):
                        q.append(("commit", both(self.blk_full2, o)))
                elif (I & 1) == mw:
                    tk, par = self.tok[mw], tph ^ (1 if mw == 0 else 0)
                    yield lambda tk=tk, par=par: tk.passed(par)
                    for o in news:
                        for c in (0, 1):
                            if self.blk_owner2[c][self.pos(o)] is not None:
                                self.errors.append(f"CTA {c}: MMA into block {self.pos(o)} still owned by output {self.blk_owner2[c][self.pos(o)]}")
                            self.blk_owner2[c][self.pos(o)] = o
                    for c in (0, 1):
                        if self.buf_row2[c][b] != I or self.tile_row2[c][b] != I:
                            self.errors.append(f"row {I} issued from buffer {b} of CTA {c} holding row {self.buf_row2[c][b]} / landed {self.tile_row2[c][b]}")
                    self.issued_by.setdefault(I, set()).add(mw)
                    q = self.queues[mw]
                    q.append(("mma", I))
                    self.tok[mw ^ 1].arrive()
                    outs = [O0 + k, O0 + k + 1] + ([O0] if k == 0 else []) + ([O0 + k + 1, O0 + k + 2, O0 + k + 2] if last else [])
                    for o in outs:
                        q.append(("commit", both(self.blk_full2, o)))
                    tph ^= 1
                else:
                    for bar in both(self.blk_full2, O0 + k):
                        bar.arrive()
                    if last:
                        self.queues[mw].append(("commit", both(self.blk_full2, O0 + k + 1)))
                        for bar in both(self.blk_full2, O0 + k + 2):
                            bar.arrive()
                yield None
                I += 1
                b += 1
                if b == self.D:
                    b, fph = 0, fph ^ 1
            O0 += ni + 2

    def epilogue2(self, cta, wg):
        O0, turn, row0 = 0, 0, 0
        for ni in self.segs:
            for j in range(ni + 2):
                mine = turn == wg
                turn = (turn + 1) % 3
                if not mine:
                    continue
                o = O0 + j
                bf, par = self.blk_full2[cta][self.pos(o)], self.use(o)
                yield lambda bf=bf, par=par: bf.passed(par)
                for k in (j - 2, j - 1, j):
                    if 0 <= k < ni:
                        g = row0 + k
                        ts = self.issued_by.get(g, set())
                        if ts != ({0} if self.single else {g & 1}) or any((t, g) not in self.mma_done for t in ts):
                            self.errors.append(f"CTA {cta}: epilogue reads output {o} before row {g} completed (issued by {ts})")
                if self.blk_owner2[cta][self.pos(o)] != o:
                    self.errors.append(f"CTA {cta}: epilogue of output {o} finds block owned by {self.blk_owner2[cta][self.pos(o)]}")
                self.blk_owner2[cta][self.pos(o)] = None
                yield None
                self.blk_empty[self.pos(o)].arrive()                 # (the peer's arrival is a remote one on the leader's barrier)
            O0 += ni + 2
            row0 += ni

    def hw_step(self):
        choices = (["tma"] if self.tma else []) + [t for t in (0, 1) if self.queues[t]]
        if not choices:
            return False
        c = self.rng.choice(choices)
        if c == "tma":
            cta, b, row = self.tma.pop(self.rng.randrange(len(self.tma)))
            self.tile_row2[cta][b] = row
            self.full[b].arrive()                                    # complete_tx on the LEADER's barrier
        else:
            op = self.queues[c].pop(0)
            if op[0] == "mma":
                self.mma_done.add((c, op[1]))
            else:
                for bar in op[1]:                                    # multicast: not atomic across the two CTAs, but no agent
                    bar.arrive()                                     # can tell (each CTA only reads its own copy)
        return True


def run_pair(nblk, stages, segments, seed, single=False):
    rng = random.Random(seed)
    m = PairModel(nblk, stages, segments, rng, single)
    agents = {"prod0": m.producer2(0), "prod1": m.producer2(1), "mma0": m.issuer(0)}
    if not single:
        agents["mma1"] = m.issuer(1)
    for c in (0, 1):
        for wg in range(3):
            agents[f"epi{c}{wg}"] = m.epilogue2(c, wg)
    waiting = {k: None for k in agents}
    frozen, frozen_for, steps = None, 0, 0
    while agents:
        steps += 1
        assert steps < 4_000_000, "model did not terminate"
        if frozen_for == 0 and rng.random() < 0.01:
            frozen, frozen_for = rng.choice(list(agents)), rng.randrange(50, 400)
        if frozen_for:
            frozen_for -= 1
        runnable = [k for k in agents if (waiting[k] is None or waiting[k]()) and not (frozen_for and k == frozen)]
        if rng.random() < 0.5 or not runnable:
            if m.hw_step():
                continue
            if not runnable:
                if frozen_for:
                    frozen_for = 0
                    continue
                return m, "deadlock: " + ", ".join(sorted(agents))
        if not runnable:
            continue
        k = rng.choice(runnable)
        try:
            waiting[k] = next(agents[k])
        except StopIteration:
            del agents[k]
            waiting.pop(k)
    if m.tma or m.queues[0] or m.queues[1]:
        m.errors.append(f"asynchronous work in flight when the CTAs leave: {len(m.tma)} loads, {len(m.queues[0]) + len(m.queues[1])} tensor-pipe ops")
    return m, None


@pytest.mark.parametrize("nblk,stages,segments", [(14, 2, [16]), (14, 5, [16]), (14, 3, [30]), (14, 5, [1, 14, 2]), (14, 2, [1] * 12),
                                                  (14, 3, [3]), (14, 3, [1]), (14, 3, [4, 1]), (7, 5, [16]), (7, 2, [9, 7]), (7, 2, [1]),
                                                  (7, 3, [2, 2, 2, 9]), (7, 2, [2] * 9), (14, 12, [40])])
@pytest.mark.parametrize("single", [False, True])
def test_pair_protocol_is_live_safe_and_quiescent_at_exit(nblk, stages, segments, single):
    """The cta_group::2 variant (ring sizes of the shadow-block layout: 14, 7 with the conv1x1), with the row-alternating
    issuers and with ESRP_PAIR_SINGLE's one issuer thread / one commit per row: live and safe in both CTAs under arbitrary
    delays of any agent, and quiescent when the last agent finishes."""
    for seed in range(120):
        m, dead = run_pair(nblk, stages, segments, seed, single)
        assert dead is None, (nblk, stages, segments, single, seed, dead)
        assert not m.errors, (nblk, stages, segments, single, seed, m.errors[:3])
