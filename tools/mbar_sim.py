"""Randomised discrete-event model of the mbarrier protocols of the attention kernels.

Why: these kernels synchronise a TMA producer thread, an MMA-issuing thread and two softmax
warpgroups through ~20 mbarriers with 1-bit phase parities.  A protocol bug is a GPU hang (round 1
lost 11 GPU-minutes to one: a warpgroup arrived on `o_free` before the MMA thread had executed its
trivially-passing parity wait, the barrier ran a phase ahead and the parity aliased).  This model
runs the same wait / arrive / commit sequences under thousands of random interleavings on the CPU and
reports (a) deadlocks, (b) reads of a buffer version other than the expected one.

mbarrier semantics modelled: a barrier has an arrival count and a phase counter; the last arrival of
a phase flips the phase.  `wait(parity)` blocks while (phase & 1) == parity, i.e. it returns once the
phase with that parity has completed - and returns IMMEDIATELY for parity 1 on a fresh barrier.
tcgen05.commit is an arrival that happens some time after the MMAs issued before it complete (the
tensor pipe executes MMAs in order).

Usage: python tools/mbar_sim.py [variant ...]   variants: gemm gemm_unpaced_release fwd_v2 fwd_lazy_bad fwd_lazy fwd_pbuf2 fwd_pbuf2_badfinal dq dkdv dkdv_pbuf2_bad dkdv_pbuf2
(fwd_v2, dq, dkdv = the shipped kernels; *_bad = known-broken protocols kept as self-tests of the model)
"""
import random
import sys


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"too many arrivals on {self.name}"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def done(self, parity):
        return (self.phase & 1) != parity


class Sim:
    """Agents are generators yielding ('wait', bar, parity) | ('arrive', bar) | ('commit', [bars]) |
    ('write', buf, version) | ('read', buf, version) | ('mma', reads, writes).  'mma' and 'commit'
    are queued on a tensor pipe that executes in order with random delays: the operand reads /
    accumulator writes of an MMA happen when it EXECUTES (not when it is issued), and a commit
    arrives on its barriers once everything queued before it has executed."""

    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.pipe_speed = self.rng.choice([0.05, 0.3, 0.9])   # tensor pipe fast or slow vs. the threads
        self.agents = {}
        self.blocked = {}
        self.pipe = []      # pending commits: list of [delay, [bars]]
        self.bufs = {}
        self.errors = []

    def add(self, name, gen, weight=None):
        # every agent runs at a random relative speed: races need one side to be much slower
        self.agents[name] = (gen, weight if weight is not None else self.rng.choice([1, 1, 4, 16]))

    def run(self, max_steps=2_000_000):
        live = dict(self.agents)
        pend = {}
        for _ in range(max_steps):
            # deliver the head of the tensor pipe with some probability
            if self.pipe and self.rng.random() < self.pipe_speed:
                self.pipe[0][0] -= 1
                if self.pipe[0][0] <= 0:
                    _, kind, a, b = self.pipe.pop(0)
                    if kind == 'c':
                        for bar in a:
                            bar.arrive()
                    else:
                        for buf, ver in a:
                            if self.bufs.get(buf) != ver:
                                self.errors.append(f"tensor pipe: read {buf} version {self.bufs.get(buf)}, expected {ver}")
                                return False
                        for buf, ver in b:
                            self.bufs[buf] = ver
            if not live:
                if self.pipe:
                    continue
                return True
            runnable = []
            for n in live:
                op = pend.get(n)
                if op is not None and op[0] == 'wait' and not op[1].done(op[2]):
                    continue
                runnable.append(n)
            if not runnable:
                if self.pipe:
                    continue
                state = {n: (pend[n][1].name, pend[n][2]) for n in live if pend.get(n)}
                self.errors.append(f"DEADLOCK: {state}")
                return False
            n = self.rng.choices(runnable, weights=[live[m][1] for m in runnable])[0]
            gen = live[n][0]
            pend[n] = None
            try:
                op = next(gen)
            except StopIteration:
                del live[n]
                continue
            if op[0] == 'wait':
                pend[n] = op
            elif op[0] == 'arrive':
                op[1].arrive()
            elif op[0] == 'commit':
                self.pipe.append([self.rng.randint(1, 6), 'c', op[1], None])
            elif op[0] == 'mma':
                self.pipe.append([self.rng.randint(1, 6), 'm', op[1], op[2]])
            elif op[0] == 'write':
                self.bufs[op[1]] = op[2]
            elif op[0] == 'read':
                if self.bufs.get(op[1]) != op[2]:
                    self.errors.append(f"{n}: read {op[1]} version {self.bufs.get(op[1])}, expected {op[2]}")
                    return False
        self.errors.append("step limit")
        return False


# ---------------------------------------------------------------------------------------------
# forward attention v2 (attn_tc.cu attn_fwd_tc2_kernel): variants of the softmax / MMA protocol
# ---------------------------------------------------------------------------------------------
def fwd(sim, ntiles, variant, stages, badfinal=False):
    T = 2
    kv_full = [Bar(f"kv_full{i}", 1) for i in range(stages)]
    kv_free = [Bar(f"kv_free{i}", 1) for i in range(stages)]
    s_full = [Bar(f"s_full{t}", 1) for t in range(T)]
    s_free = [Bar(f"s_free{t}", 1) for t in range(T)]          # one agent stands for the 128 threads
    o_full = [Bar(f"o_full{t}", 1) for t in range(T)]
    nb = 2 if variant == "pbuf2" else 1
    p_full = [[Bar(f"p_full{t}{b}", 1) for b in range(nb)] for t in range(T)]
    p_free = [[Bar(f"p_free{t}{b}", 1) for b in range(nb)] for t in range(T)]
    o_free = [[Bar(f"o_free{t}{b}", 1) for b in range(nb)] for t in range(T)]
    rescale = lambda j: (j * 7 + 3) % 5 == 0   # pseudo-random rows need an O rescale

    def producer():
        st, ph = 0, 0
        for j in range(ntiles):
            yield ('wait', kv_free[st], ph ^ 1)
            yield ('write', f"K{st}", j)
            yield ('write', f"V{st}", j)
            yield ('arrive', kv_full[st])
            st += 1
            if st == stages:
                st, ph = 0, ph ^ 1

    def mma():
        st = stv = ph = 0
        for j in range(ntiles + 1):
            if j < ntiles:
                yield ('wait', kv_full[st], ph)
                for t in range(T):
                    yield ('wait', s_free[t], (j & 1) ^ 1)
                    yield ('mma', [(f"K{st}", j)], [(f"S{t}", j)])
                    yield ('commit', [s_full[t]])
            if j > 0:
                jj = j - 1
                for t in range(T):
                    if variant == "pbuf2":
                        bb, use = jj & 1, jj >> 1
                        yield ('wait', p_full[t][bb], use & 1)
                        if jj > 0:
                            yield ('wait', o_free[t][bb], (use if bb else use - 1) & 1)
                    else:
                        bb = 0
                        yield ('wait', p_full[t][0], jj & 1)
                        yield ('wait', o_free[t][0], (jj & 1) ^ 1)
                    if variant in ("lazy", "lazy_bad", "pbuf2"):
                        # O accumulates in TMEM: block jj adds to the sum of blocks < jj, which the
                        # softmax warpgroup may have rescaled in place just before
                        o_in = [] if jj == 0 else [(f"O{t}", ('r', jj) if rescale(jj) else jj - 1)]
                        yield ('mma', [(f"P{t}{bb}", jj), (f"V{stv}", jj)] + o_in, [(f"O{t}", jj)])
                    else:
                        yield ('mma', [(f"P{t}{bb}", jj), (f"V{stv}", jj)], [])
                    yield ('commit', [o_full[t], p_free[t][bb]])
                yield ('commit', [kv_free[stv]])
                stv = (stv + 1) % stages
            if j < ntiles:
                st += 1
                if st == stages:
                    st, ph = 0, ph ^ 1

    def softmax(t):
        for j in range(ntiles):
            yield ('wait', s_full[t], j & 1)
            yield ('read', f"S{t}", j)
            yield ('arrive', s_free[t])
            if variant == "v2":          # shipped kernel: O partial products read back every block
                if j > 0:
                    yield ('wait', p_free[t][0], (j - 1) & 1)
                yield ('write', f"P{t}0", j)
                yield ('arrive', p_full[t][0])
                if j > 0:
                    yield ('wait', o_full[t], (j - 1) & 1)
                    yield ('arrive', o_free[t][0])
            elif variant == "lazy_bad":  # first lazy-rescale attempt (deadlocked on the GPU)
                if j > 0:
                    if rescale(j):
                        yield ('wait', o_full[t], (j - 1) & 1)
                        yield ('read', f"O{t}", j - 1)
                        yield ('write', f"O{t}", ('r', j))
                    yield ('arrive', o_free[t][0])
                    yield ('wait', p_free[t][0], (j - 1) & 1)
                yield ('write', f"P{t}0", j)
                yield ('arrive', p_full[t][0])
            elif variant == "lazy":      # fixed: arrive on o_free only after p_free(j-1)
                if j > 0:
                    yield ('wait', p_free[t][0], (j - 1) & 1)
                    if rescale(j):
                        yield ('wait', o_full[t], (j - 1) & 1)
                        yield ('read', f"O{t}", j - 1)
                        yield ('write', f"O{t}", ('r', j))
                    yield ('arrive', o_free[t][0])
                yield ('write', f"P{t}0", j)
                yield ('arrive', p_full[t][0])
            elif variant == "pbuf2":     # double-buffered P, o_free per buffer
                bb = j & 1
                if j >= 2:
                    yield ('wait', p_free[t][bb], ((j >> 1) - 1) & 1)
                if j > 0:
                    if rescale(j):
                        yield ('wait', o_full[t], (j - 1) & 1)
                        yield ('read', f"O{t}", j - 1)
                        yield ('write', f"O{t}", ('r', j))
                    yield ('arrive', o_free[t][bb])
                yield ('write', f"P{t}{bb}", j)
                yield ('arrive', p_full[t][bb])
        if variant == "pbuf2" and badfinal and ntiles >= 2:
            yield ('wait', o_full[t], (ntiles - 2) & 1)   # aliases when o_full is already 2 phases on
        elif variant == "pbuf2":
            if ntiles >= 2:
                yield ('wait', p_free[t][(ntiles - 2) & 1], ((ntiles - 2) >> 1) & 1)
        yield ('wait', o_full[t], (ntiles - 1) & 1)
        if variant in ("lazy", "lazy_bad", "pbuf2"):
            yield ('read', f"O{t}", ntiles - 1)   # final read-out of the accumulated O

    sim.add("producer", producer())
    sim.add("mma", mma())
    for t in range(T):
        sim.add(f"softmax{t}", softmax(t))


# ---------------------------------------------------------------------------------------------
# backward dK/dV (attn_bwd_tc.cu attn_bwd_dkdv_tc_kernel): single- and double-buffered P^T / dS^T
# ---------------------------------------------------------------------------------------------
def dkdv(sim, ntiles, variant):
    ST = 2
    q_full = [Bar(f"q_full{i}", 1) for i in range(ST)]
    q_free = [Bar(f"q_free{i}", 1) for i in range(ST)]
    ld_full = [Bar(f"ld_full{i}", 1) for i in range(ST)]
    s_full, s_free = Bar("s_full", 1), Bar("s_free", 2)
    nb = 2 if variant.startswith("pbuf2") else 1
    p_full = [Bar(f"p_full{b}", 2) for b in range(nb)]
    p_free = [Bar(f"p_free{b}", 1) for b in range(nb)]
    acc_full = Bar("acc_full", 1)

    def producer():
        for j in range(ntiles):
            st = j % ST
            yield ('wait', q_free[st], ((j // ST) & 1) ^ 1)
            yield ('write', f"Q{st}", j)
            yield ('arrive', q_full[st])

    def mma():
        def scores(j):
            st = j % ST
            yield ('wait', q_full[st], (j // ST) & 1)
            yield ('wait', s_free, (j & 1) ^ 1)
            yield ('mma', [(f"Q{st}", j)], [("S", j)])
            yield ('commit', [s_full])
        yield from scores(0)
        for j in range(ntiles):
            st = j % ST
            if j + 1 < ntiles:
                yield from scores(j + 1)
            if variant.startswith("pbuf2"):
                yield ('wait', p_full[j & 1], (j >> 1) & 1)
                yield ('mma', [(f"P{j & 1}", j), (f"Q{st}", j)], [])
                yield ('commit', [p_free[j & 1], q_free[st]])
            else:
                yield ('wait', p_full[0], j & 1)
                yield ('mma', [("P0", j), (f"Q{st}", j)], [])
                yield ('commit', [p_free[0], q_free[st]])
        yield ('commit', [acc_full])

    def elem(g):
        for j in range(ntiles):
            st = j % ST
            if variant == "pbuf2_bad":   # statistics staged at the top of the iteration (as in the
                if g == 0:               # single-buffered kernel): races with a slow warpgroup
                    yield ('write', f"L{st}", j)
                    yield ('arrive', ld_full[st])
                yield ('wait', ld_full[st], (j // ST) & 1)
                yield ('wait', s_full, j & 1)
            elif variant == "pbuf2":     # staged only after S^T(j) exists: both warpgroups are past tile j-2
                yield ('wait', s_full, j & 1)
                if g == 0:
                    yield ('write', f"L{st}", j)
                    yield ('arrive', ld_full[st])
                yield ('wait', ld_full[st], (j // ST) & 1)
            else:
                if g == 0:
                    yield ('write', f"L{st}", j)
                    yield ('arrive', ld_full[st])
                yield ('wait', ld_full[st], (j // ST) & 1)
                yield ('wait', s_full, j & 1)
            yield ('read', "S", j)
            yield ('arrive', s_free)
            yield ('read', f"L{st}", j)
            if variant.startswith("pbuf2"):
                if j >= 2:
                    yield ('wait', p_free[j & 1], ((j >> 1) - 1) & 1)
                if g == 0:
                    yield ('write', f"P{j & 1}", j)
                yield ('arrive', p_full[j & 1])
            else:
                if j > 0:
                    yield ('wait', p_free[0], (j - 1) & 1)
                if g == 0:
                    yield ('write', "P0", j)
                yield ('arrive', p_full[0])
        yield ('wait', acc_full, 0)

    sim.add("producer", producer())
    sim.add("mma", mma())
    for g in range(2):
        sim.add(f"elem{g}", elem(g))


# ---------------------------------------------------------------------------------------------
# backward dQ (attn_bwd_tc.cu attn_bwd_dq_tc_kernel, shipped): double-buffered dS, 3 K/V stages
# ---------------------------------------------------------------------------------------------
def dq(sim, ntiles):
    ST = 3
    kv_full = [Bar(f"kv_full{i}", 1) for i in range(ST)]
    kv_free = [Bar(f"kv_free{i}", 1) for i in range(ST)]
    s_full, s_free = Bar("s_full", 1), Bar("s_free", 2)
    p_full = [Bar(f"p_full{b}", 2) for b in range(2)]
    p_free = [Bar(f"p_free{b}", 1) for b in range(2)]
    acc_full = Bar("acc_full", 1)

    def producer():
        for j in range(ntiles):
            st = j % ST
            yield ('wait', kv_free[st], ((j // ST) & 1) ^ 1)
            yield ('write', f"KV{st}", j)
            yield ('arrive', kv_full[st])

    def mma():
        def scores(j):
            st = j % ST
            yield ('wait', kv_full[st], (j // ST) & 1)
            yield ('wait', s_free, (j & 1) ^ 1)
            yield ('mma', [(f"KV{st}", j)], [("S", j)])
            yield ('commit', [s_full])
        yield from scores(0)
        for j in range(ntiles):
            st = j % ST
            if j + 1 < ntiles:
                yield from scores(j + 1)
            yield ('wait', p_full[j & 1], (j >> 1) & 1)
            yield ('mma', [(f"dS{j & 1}", j), (f"KV{st}", j)], [])
            yield ('commit', [p_free[j & 1], kv_free[st]])
        yield ('commit', [acc_full])

    def elem(g):
        for j in range(ntiles):
            yield ('wait', s_full, j & 1)
            yield ('read', "S", j)
            yield ('arrive', s_free)
            if j >= 2:
                yield ('wait', p_free[j & 1], ((j >> 1) - 1) & 1)
            if g == 0:
                yield ('write', f"dS{j & 1}", j)
            yield ('arrive', p_full[j & 1])
        yield ('wait', acc_full, 0)

    sim.add("producer", producer())
    sim.add("mma", mma())
    for g in range(2):
        sim.add(f"elem{g}", elem(g))


# ---------------------------------------------------------------------------------------------
# implicit GEMM (gemm_tc.cu): S-stage smem ring, two TMEM accumulators, 8 epilogue warps
# ---------------------------------------------------------------------------------------------
def gemm(sim, ntiles, stages=3, kblocks=4, early_release_unpaced=False):
    full = [Bar(f"full{i}", 1) for i in range(stages)]
    empty = [Bar(f"empty{i}", 1) for i in range(stages)]
    tfull = [Bar(f"tfull{i}", 1) for i in range(2)]
    tempty = [Bar(f"tempty{i}", 8) for i in range(2)]

    def producer():
        st = ph = 0
        for t in range(ntiles):
            for kb in range(kblocks):
                yield ('wait', empty[st], ph ^ 1)
                yield ('write', f"stage{st}", (t, kb))
                yield ('arrive', full[st])
                st += 1
                if st == stages:
                    st, ph = 0, ph ^ 1

    def mma():
        st = ph = acc = aph = 0
        for t in range(ntiles):
            yield ('wait', tempty[acc], aph ^ 1)
            for kb in range(kblocks):
                yield ('wait', full[st], ph)
                yield ('mma', [(f"stage{st}", (t, kb))], [(f"acc{acc}", (t, kb))])
                yield ('commit', [empty[st]])
                st += 1
                if st == stages:
                    st, ph = 0, ph ^ 1
            yield ('commit', [tfull[acc]])
            acc ^= 1
            if acc == 0:
                aph ^= 1

    def epi(w):
        acc = aph = 0
        idle = (w == 7)   # a warp whose group has no chunk in these tiles (block_n = 32)
        for t in range(ntiles):
            if not (idle and early_release_unpaced):
                yield ('wait', tfull[acc], aph)
            if not idle:
                yield ('read', f"acc{acc}", (t, kblocks - 1))
            yield ('arrive', tempty[acc])
            acc ^= 1
            if acc == 0:
                aph ^= 1

    sim.add("producer", producer())
    sim.add("mma", mma())
    for w in range(8):
        sim.add(f"epi{w}", epi(w))


VARIANTS = {
    "fwd_v2": lambda s, n: fwd(s, n, "v2", 3),
    "fwd_lazy_bad": lambda s, n: fwd(s, n, "lazy_bad", 3),
    "fwd_lazy": lambda s, n: fwd(s, n, "lazy", 3),
    "fwd_pbuf2": lambda s, n: fwd(s, n, "pbuf2", 2),
    "fwd_pbuf2_badfinal": lambda s, n: fwd(s, n, "pbuf2", 2, badfinal=True),
    "gemm": lambda s, n: gemm(s, n),
    "gemm_unpaced_release": lambda s, n: gemm(s, n, early_release_unpaced=True),
    "dq": lambda s, n: dq(s, n),
    "dkdv": lambda s, n: dkdv(s, n, "single"),
    "dkdv_pbuf2_bad": lambda s, n: dkdv(s, n, "pbuf2_bad"),
    "dkdv_pbuf2": lambda s, n: dkdv(s, n, "pbuf2"),
}


def check(variant, trials=1500):
    bad = None
    for seed in range(trials):
        for ntiles in (1, 2, 3, 4, 7, 8):
            sim = Sim(seed * 31 + ntiles)
            VARIANTS[variant](sim, ntiles)
            if not sim.run():
                bad = (seed, ntiles, sim.errors[-1])
                break
        if bad:
            break
    return bad


if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for v in names:
        bad = check(v)
        print(f"{v:14s}", "OK (no deadlock / stale read in 9000 random schedules)" if bad is None
              else f"FAIL seed={bad[0]} ntiles={bad[1]}: {bad[2]}")
