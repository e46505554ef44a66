"""Discrete-event check of the mbarrier protocol of attn_fwd_kernel (csrc/attn_fwd.cu) - no GPU needed.

The roles of one CTA - TMA producer, tcgen05 issuer, the 2 x 4 softmax / epilogue warps - are written here as Python
generators that perform the same sequence of barrier operations, with the same phase-parity expressions and ring
counters, as the CUDA code (HALF hand-off of P on, MUFU turn-taking off: the shipped configuration).  The tensor pipe
is a fourth actor: MMAs and commits are queued in issue order and execute later, one after the other, so the check
covers what the barriers are there for - an operand being overwritten while a queued MMA still needs it:

  * S_t / P_t alias in TMEM, per warp quadrant: free -> S (QK executed) -> P half 0 -> P half 1 -> free (PV executed);
  * O_t: a first PV (accumulate = false) may only execute after the previous owner's epilogue has read O_t out
    (o_free), the lazy rescale and the epilogue only touch O_t when no PV of that tile is still queued;
  * K / V ring stages and the Q buffers (which double as the staging tiles of the output) are only overwritten when
    no queued MMA and no TMA store still reads them, and are only read when they hold the tile the step wants.

Barriers follow mbarrier semantics: `arrive` decrements the pending count of the current phase, completing it (and
re-arming the count) at zero; `wait(parity)` passes iff the current phase parity differs from `parity` - so a waiter
two phases behind blocks forever.  That is the deadlock of the first multi-item run at head_dim 64 in round 2
(`round2_bug=True` restores that version): o_free completed a phase in EVERY item, and for a query tile without key
tiles - the ViT's last query block has an empty second tile - the issuer consumed the phase at the END of the item, by
which time the tile's epilogue warps (which wait for nothing but the Q load) could already have arrived for the
current item as well.  The scheduler picks runnable actors at random; a
run fails on deadlock or on any of the operand checks.  `skip` removes single waits, to show that each one is needed
and that the checks see its absence (tests/test_protocol_sim.py).  One is not: the o_free wait is implied by p_half -
the warps that read O_t out in the epilogue are the ones that write the next item's first P_t, so the first PV of the
next item cannot be issued before they are done - and stays in the kernel as an explicit guard.

    python tools/sim_fwd_protocol.py [--seeds 300]
"""
import argparse
import random


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than the barrier expects"
        if self.pending == 0:
            self.pending = self.count
            self.phase += 1

    def passed(self, parity):
        return (self.phase & 1) != parity


class Sim:
    """items: list of (n0, n1) = key tiles visible to the two 128-row query tiles of each work item of this CTA.
    NS: K / V ring stages (2 at head_dim 128, 4 at 64).  QB: Q buffers per query tile (LV_ATTN_QBUF64)."""

    def __init__(self, items, seed, NS=2, QB=1, skip=(), round2_bug=False):
        self.round2_bug = round2_bug
        self.rng = random.Random(seed)
        self.items, self.NS, self.QB = items, NS, QB
        self.skip = set(skip)        # waits left out on purpose: o_free | p_full | p_half | k_empty | v_empty | q_empty | o_full
        B = Bar
        self.b = {}
        for i in range(2 * QB):
            self.b[f"q_full{i}"] = B(f"q_full{i}", 1)
            self.b[f"q_empty{i}"] = B(f"q_empty{i}", 1)
        for t in range(2):
            self.b[f"s_full{t}"] = B(f"s_full{t}", 1)
            self.b[f"p_full{t}"] = B(f"p_full{t}", 4)
            self.b[f"p_half{t}"] = B(f"p_half{t}", 4)
            self.b[f"o_full{t}"] = B(f"o_full{t}", 1)
            self.b[f"o_free{t}"] = B(f"o_free{t}", 4)
            self.b[f"named{t}"] = B(f"named{t}", 4)          # bar.sync 1 + t, 128
        for i in range(NS):
            for n in ("k_full", "k_empty", "v_full", "v_empty"):
                self.b[f"{n}{i}"] = B(f"{n}{i}", 1)
        self.time = 0
        self.async_q = []            # TMA completions: (due time, barrier, callback)
        self.pipe = []               # tensor pipe, in issue order: ("qk" | "pv" | "commit", ...)
        self.pipe_ready = 0
        # operand state
        self.q_fill = [None] * (2 * QB)          # item whose Q tile sits in the buffer
        self.stage_busy = [False] * (2 * QB)     # output staged there and the TMA store has not read it yet
        self.k_fill = [None] * NS
        self.v_fill = [None] * NS
        self.sp = [["free"] * 4 for _ in range(2)]        # S_t / P_t per warp quadrant
        self.sp_id = [None, None]
        self.o_state = ["free", "free"]                   # free | acc | (epilogue reading: tracked by o_read)
        self.o_owner = [None, None]
        self.o_read = [set(), set()]
        self.pv_exec = [0, 0]                             # executed PV halves of the current owner

    # ---- tensor pipe ---------------------------------------------------------------------------------------------
    def pending(self, pred):
        return any(pred(op) for op in self.pipe)

    def pipe_step(self):
        if not self.pipe or self.time < self.pipe_ready:
            return False
        op = self.pipe.pop(0)
        self.pipe_ready = self.time + self.rng.randint(1, 4)
        kind = op[0]
        if kind == "commit":
            self.b[op[1]].arrive()
        elif kind == "qk":
            _, t, item, j, qbuf, kst = op
            assert self.q_fill[qbuf] == item and not self.stage_busy[qbuf], f"QK{t} of item {item} reads Q buffer {qbuf} = {self.q_fill[qbuf]}"
            assert self.k_fill[kst] == (item, j), f"QK{t}({item},{j}) reads K stage {kst} holding {self.k_fill[kst]}"
            assert all(s == "free" for s in self.sp[t]), f"QK{t}({item},{j}) overwrites S/P in state {self.sp[t]} of {self.sp_id[t]}"
            self.sp[t] = ["S"] * 4
            self.sp_id[t] = (item, j)
        else:
            _, t, item, j, half, vst, acc = op
            assert self.v_fill[vst] == (item, j), f"PV{t}({item},{j}) reads V stage {vst} holding {self.v_fill[vst]}"
            assert self.sp_id[t] == (item, j), f"PV{t}({item},{j}) reads P of {self.sp_id[t]}"
            need = ("P0", "P1") if half == 0 else ("P1",)
            assert all(s in need for s in self.sp[t]), f"PV{t}({item},{j}) half {half} before P was written: {self.sp[t]}"
            if half == 0 and not acc:
                assert self.o_state[t] == "free", f"first PV{t} of item {item} overwrites O of item {self.o_owner[t]} ({self.o_state[t]})"
                self.o_state[t], self.o_owner[t], self.pv_exec[t] = "acc", item, 0
            else:
                assert self.o_state[t] == "acc" and self.o_owner[t] == item, f"PV{t}({item},{j}) accumulates into O of {self.o_owner[t]}"
            self.pv_exec[t] += 1
            if half == 1:
                self.sp[t] = ["free"] * 4
        return True

    def later(self, bar, cb):
        self.async_q.append((self.time + self.rng.randint(1, 6), bar, cb))

    # ---- roles ---------------------------------------------------------------------------------------------------
    def producer(self):
        NS, QB = self.NS, self.QB
        kcnt = vcnt = 0
        for item_cnt, (n0, n1) in enumerate(self.items):
            nmax = max(n0, n1)
            qb, qpar = item_cnt % QB, (item_cnt // QB) & 1

            def load_kv(j, item=item_cnt):
                nonlocal kcnt, vcnt
                st = kcnt % NS
                if "k_empty" not in self.skip:
                    yield ("wait", f"k_empty{st}", ((kcnt // NS) & 1) ^ 1)
                assert not self.pending(lambda op: op[0] == "qk" and op[5] == st), f"K stage {st} overwritten under a queued QK"
                self.later(f"k_full{st}", lambda st=st: self.k_fill.__setitem__(st, (item, j)))
                kcnt += 1
                st = vcnt % NS
                if "v_empty" not in self.skip:
                    yield ("wait", f"v_empty{st}", ((vcnt // NS) & 1) ^ 1)
                assert not self.pending(lambda op: op[0] == "pv" and op[5] == st), f"V stage {st} overwritten under a queued PV"
                self.later(f"v_full{st}", lambda st=st: self.v_fill.__setitem__(st, (item, j)))
                vcnt += 1

            pre = min(NS, nmax) if QB == 1 else 0
            for j in range(pre):
                yield from load_kv(j)
            for t in range(2):
                buf = qb * 2 + t
                if "q_empty" not in self.skip:
                    yield ("wait", f"q_empty{buf}", qpar ^ 1)
                assert not self.stage_busy[buf], f"Q buffer {buf} overwritten while its TMA store is in flight"
                assert not self.pending(lambda op: op[0] == "qk" and op[4] == buf), f"Q buffer {buf} overwritten under a queued QK"
                self.later(f"q_full{buf}", lambda buf=buf, item=item_cnt: self.q_fill.__setitem__(buf, item))
            for j in range(pre, nmax):
                yield from load_kv(j)

    def of_counts(self, n0, n1):
        """Whether o_free[t] completes a phase in an item (both sides must agree): only if tile t has key tiles."""
        return n0 > 0, n1 > 0

    def issuer(self):
        NS, QB = self.NS, self.QB
        kcnt = vcnt_wait = vcnt_rel = 0
        pcnt = [0, 0]
        of_cnt = [0, 0]

        def issue_pv(t, item, j, vst, acc):
            self.pipe.append(("pv", t, item, j, 0, vst, acc))
            if "p_full" not in self.skip:
                yield ("wait", f"p_full{t}", (pcnt[t] - 1) & 1)
            self.pipe.append(("pv", t, item, j, 1, vst, True))

        for item_cnt, (n0, n1) in enumerate(self.items):
            nmax = max(n0, n1)
            qb, qpar = item_cnt % QB, (item_cnt // QB) & 1
            yield ("wait", f"q_full{qb * 2 + 0}", qpar)
            yield ("wait", f"q_full{qb * 2 + 1}", qpar)
            o_waited = [False, False] if not self.round2_bug else [item_cnt == 0, item_cnt == 0]
            vbase = vcnt_wait
            for j in range(nmax + 1):
                kst = 0
                if j < nmax:
                    kst = kcnt % NS
                    yield ("wait", f"k_full{kst}", (kcnt // NS) & 1)
                if j < n0:
                    self.pipe.append(("qk", 0, item_cnt, j, qb * 2 + 0, kst))
                    self.pipe.append(("commit", "s_full0"))
                if j >= 1:
                    if j - 1 < n1:
                        vc = vbase + j - 1
                        if vc == vcnt_wait:
                            yield ("wait", f"v_full{vc % NS}", (vc // NS) & 1)
                            vcnt_wait += 1
                        if "p_half" not in self.skip:
                            yield ("wait", "p_half1", pcnt[1] & 1)
                        pcnt[1] += 1
                        if not o_waited[1]:
                            if self.round2_bug:
                                yield ("wait", "o_free1", (item_cnt - 1) & 1)
                            elif of_cnt[1] > 0 and "o_free" not in self.skip:
                                yield ("wait", "o_free1", (of_cnt[1] - 1) & 1)
                            o_waited[1] = True
                        yield from issue_pv(1, item_cnt, j - 1, vc % NS, j - 1 > 0)
                        if j - 1 == n1 - 1:
                            self.pipe.append(("commit", "o_full1"))
                    self.pipe.append(("commit", f"v_empty{vcnt_rel % NS}"))
                    vcnt_rel += 1
                if j < n1:
                    self.pipe.append(("qk", 1, item_cnt, j, qb * 2 + 1, kst))
                    self.pipe.append(("commit", "s_full1"))
                if j < nmax:
                    self.pipe.append(("commit", f"k_empty{kst}"))
                    kcnt += 1
                if j < n0:
                    vc = vbase + j
                    if vc == vcnt_wait:
                        yield ("wait", f"v_full{vc % NS}", (vc // NS) & 1)
                        vcnt_wait += 1
                    if "p_half" not in self.skip:
                        yield ("wait", "p_half0", pcnt[0] & 1)
                    pcnt[0] += 1
                    if not o_waited[0]:
                        if self.round2_bug:
                            yield ("wait", "o_free0", (item_cnt - 1) & 1)
                        elif of_cnt[0] > 0 and "o_free" not in self.skip:
                            yield ("wait", "o_free0", (of_cnt[0] - 1) & 1)
                        o_waited[0] = True
                    yield from issue_pv(0, item_cnt, j, vc % NS, j > 0)
                    if j == n0 - 1:
                        self.pipe.append(("commit", "o_full0"))
            if self.round2_bug:      # "a tile without key tiles issued no P.V: consume its o_free phase all the same"
                for t in range(2):
                    if not o_waited[t]:
                        yield ("wait", f"o_free{t}", (item_cnt - 1) & 1)
            c0, c1 = self.of_counts(n0, n1)
            of_cnt[0] += int(c0)
            of_cnt[1] += int(c1)

    def softmax_arrives_o_free(self, n):
        return n > 0 or self.round2_bug

    def softmax_warp(self, t, quad):
        QB = self.QB
        scnt = ocnt = named_cnt = 0
        for item_cnt, ns in enumerate(self.items):
            n = ns[t]
            qb = item_cnt % QB
            buf = qb * 2 + t
            for j in range(n):
                yield ("wait", f"s_full{t}", scnt & 1)
                scnt += 1
                assert self.sp_id[t] == (item_cnt, j) and self.sp[t][quad] == "S", \
                    f"warp ({t},{quad}) reads S of {self.sp_id[t]} in state {self.sp[t][quad]}, wants {(item_cnt, j)}"
                yield ("work", self.rng.randint(1, 3))                      # tcgen05.ld, mask, row max
                if j > 0:                                                     # lazy rescale may touch O_t
                    assert not self.pending(lambda op: op[0] == "pv" and op[1] == t), f"rescale of O{t} under a queued PV"
                    assert self.o_state[t] == "acc" and self.o_owner[t] == item_cnt
                yield ("work", self.rng.randint(1, 4))                      # exponentials, first half
                self.sp[t][quad] = "P0"
                self.b[f"p_half{t}"].arrive()
                yield ("work", self.rng.randint(1, 4))
                self.sp[t][quad] = "P1"
                self.b[f"p_full{t}"].arrive()
            if n > 0:
                if "o_full" not in self.skip:
                    yield ("wait", f"o_full{t}", ocnt & 1)
                ocnt += 1
                assert self.o_state[t] == "acc" and self.o_owner[t] == item_cnt and self.pv_exec[t] == 2 * n, \
                    f"epilogue ({t},{quad}) of item {item_cnt} reads O of {self.o_owner[t]} after {self.pv_exec[t]} of {2 * n} PV halves"
                assert not self.pending(lambda op: op[0] == "pv" and op[1] == t)
            else:
                yield ("wait", f"q_full{buf}", (item_cnt // QB) & 1)
            yield ("work", self.rng.randint(1, 3))                          # tcgen05.ld of O_t
            if n > 0:
                self.o_read[t].add(quad)
                if len(self.o_read[t]) == 4:
                    self.o_read[t] = set()
                    self.o_state[t] = "free"
            if self.softmax_arrives_o_free(n):
                self.b[f"o_free{t}"].arrive()
            # O / l -> the item's Q_t tile (staging)
            assert self.q_fill[buf] == item_cnt, f"staging tile {buf} holds Q of item {self.q_fill[buf]}, not {item_cnt}"
            assert not self.pending(lambda op: op[0] == "qk" and op[4] == buf), f"staging tile {buf} written under a queued QK"
            self.stage_busy[buf] = True
            yield ("work", self.rng.randint(1, 3))
            self.b[f"named{t}"].arrive()
            yield ("wait", f"named{t}", named_cnt & 1)
            named_cnt += 1
            if quad == 0:
                yield ("work", self.rng.randint(1, 5))                      # TMA store, wait until the tile has been read
                self.stage_busy[buf] = False
                self.b[f"q_empty{buf}"].arrive()

    # ---- scheduler -----------------------------------------------------------------------------------------------
    def run(self):
        roles = {"producer": self.producer(), "issuer": self.issuer()}
        for t in range(2):
            for q in range(4):
                roles[f"wg{t}.{q}"] = self.softmax_warp(t, q)
        blocked, done = {}, set()
        while len(done) < len(roles) or self.pipe:
            self.time += 1
            due = [a for a in self.async_q if a[0] <= self.time]
            self.async_q = [a for a in self.async_q if a[0] > self.time]
            for _, bar, cb in sorted(due, key=lambda a: a[0]):
                cb()
                self.b[bar].arrive()
            progressed = self.pipe_step()
            names = [r for r in roles if r not in done]
            self.rng.shuffle(names)
            for r in names:
                op = blocked.get(r)
                if op and op[0] == "wait" and not self.b[op[1]].passed(op[2]):
                    continue
                if op and op[0] == "work" and self.time < op[1]:
                    progressed = True
                    continue
                try:
                    nxt = next(roles[r])
                    blocked[r] = ("work", self.time + nxt[1]) if nxt[0] == "work" else nxt
                    progressed = True
                except StopIteration:
                    done.add(r)
                    blocked.pop(r, None)
                    progressed = True
            if not progressed and not self.async_q and not (self.pipe and self.time < self.pipe_ready):
                raise RuntimeError(f"deadlock at t={self.time}: " + ", ".join(f"{r} on {blocked.get(r)}" for r in names)
                                   + " | phases " + str({k: v.phase for k, v in self.b.items() if v.phase}))
        assert all(s == "free" for t in range(2) for s in self.sp[t]) and self.o_state == ["free", "free"]
        return self.time


def random_items(rng):
    """Work lists like the kernel's: causal items (n0 <= n1), the ViT's short items, items whose second query tile is
    out of range (n1 = 0), and - for the protocol's sake - the shapes no launch produces (n0 > n1 > 0, n0 = 0)."""
    out = []
    for _ in range(rng.randint(1, 8)):
        kind = rng.random()
        if kind < 0.45:
            a = rng.randint(1, 9)
            out.append((a, a + rng.randint(0, 1)))
        elif kind < 0.7:
            out.append((rng.randint(1, 9), 0))
        elif kind < 0.85:
            a = rng.randint(1, 9)
            out.append((a, a))
        elif kind < 0.95:
            out.append((rng.randint(1, 9), rng.randint(1, 9)))
        else:
            out.append((0, rng.randint(0, 3)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=300)
    a = ap.parse_args()
    for seed in range(a.seeds):
        rng = random.Random(5000 + seed)
        items = random_items(rng)
        for NS, QB in ((2, 1), (4, 1), (4, 2)):
            Sim(items, seed, NS, QB).run()
    print(f"ok: {a.seeds} random work lists x 3 configurations, no deadlock, no operand hazard")


if __name__ == "__main__":
    main()
