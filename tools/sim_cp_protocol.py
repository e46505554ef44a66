"""Discrete-event check of the inter-GPU protocol of the fused context-parallel attention (csrc/attn_fwd.cu: cp_copier;
DESIGN.md section 5, "Buffer-reuse protocol") - no GPU needed.

Every rank runs, in stream order and once per layer ("epoch" e): write its K|V rows of epoch e into the peer-mapped
buffer e mod 2 (the QKV GEMM), then the attention kernel of epoch e.  Inside that kernel
  * copier 0 first stores e + 1 into its ready word at every peer (st.release.sys after a system fence);
  * every copier unit waits until the owner's ready word it sees is >= e + 1 (ld.acquire.sys), then reads the owner's
    buffer e mod 2 over NVLink;
  * copier 0 does not let the kernel retire before it has seen every peer's word >= e + 1.
There is no barrier between layers and no collective: the claim checked here is that with ranks running at arbitrary
relative speeds (a rank may be a whole kernel ahead of a peer) every read of a peer's buffer sees exactly the epoch it
wants - neither rows that are still being written nor rows already overwritten by the buffer's next use two epochs
later - and that nobody waits forever.  `variant` removes one element of the protocol to show that it is needed.
(The exit wait is needed only when a rank does not read from every peer: with the zig-zag layout every rank's second
chunk sees keys of all peers, so the copier waits already cover it; `reads_from` lets a test thin the read pattern out.)

    python tools/sim_cp_protocol.py [--seeds 300]
"""
import argparse
import random


class Rank:
    def __init__(self, r, cp):
        self.r = r
        self.ready = [0] * cp            # my_ready[q]: last epoch + 1 announced by peer q (own entry unused)
        self.buf = [None, None]          # per parity: epoch whose rows the buffer holds
        self.writing = [False, False]
        self.readers = [0, 0]            # peers currently reading the buffer


class Sim:
    def __init__(self, cp, epochs, seed, variant=""):
        self.cp, self.epochs, self.variant = cp, epochs, variant
        self.rng = random.Random(seed)
        self.ranks = [Rank(r, cp) for r in range(cp)]
        self.time = 0
        self.speed = [self.rng.choice([1, 1, 2, 5, 12]) for _ in range(cp)]     # some ranks much slower than others

    def reads_from(self, r):
        """Peers whose rows rank r pulls in a kernel.  Zig-zag sharding: all of them (chunk 2cp-1-r sees every first-half
        chunk)."""
        return [q for q in range(self.cp) if q != r]

    def dur(self, r, lo, hi):
        return self.rng.randint(lo, hi) * self.speed[r]

    # one rank's stream: generators yield ("work", ticks) or ("until", predicate, label)
    def stream(self, r):
        me = self.ranks[r]
        cp = self.cp
        for e in range(self.epochs):
            par = e & 1
            # ---- QKV GEMM of this layer: overwrites buffer e mod 2 ----
            assert me.readers[par] == 0, f"rank {r} overwrites buffer {par} for epoch {e} while {me.readers[par]} peer(s) still read epoch {me.buf[par]}"
            me.writing[par] = True
            if self.variant == "flag_before_write":
                for q in range(cp):
                    if q != r:
                        self.ranks[q].ready[r] = e + 1
            yield ("work", self.dur(r, 1, 4))
            me.writing[par] = False
            me.buf[par] = e
            # ---- attention kernel of epoch e ----
            if self.variant != "flag_before_write":
                for q in range(cp):
                    if q != r:
                        self.ranks[q].ready[r] = e + 1
            yield ("work", self.dur(r, 0, 2))
            # copier units in this rank's priority order; each reads one peer's rows
            owners = self.reads_from(r)
            self.rng.shuffle(owners)
            for q in owners:
                if self.variant != "no_ready_wait":
                    yield ("until", lambda q=q, e=e: me.ready[q] >= e + 1, f"ready word of rank {q} >= {e + 1}")
                peer = self.ranks[q]
                assert not peer.writing[par] and peer.buf[par] == e, \
                    f"rank {r} epoch {e} reads buffer {par} of rank {q} holding epoch {peer.buf[par]} (writing={peer.writing[par]})"
                peer.readers[par] += 1
                yield ("work", self.dur(r, 1, 5))
                assert not peer.writing[par] and peer.buf[par] == e, f"rank {q} overwrote buffer {par} under rank {r}'s read of epoch {e}"
                peer.readers[par] -= 1
            yield ("work", self.dur(r, 1, 6))                    # the attention itself
            if self.variant != "no_exit_wait":
                for q in range(cp):
                    if q != r:
                        yield ("until", lambda q=q, e=e: me.ready[q] >= e + 1, f"exit: word of rank {q} >= {e + 1}")

    def run(self):
        procs = {r: self.stream(r) for r in range(self.cp)}
        state = {}
        done = set()
        while len(done) < self.cp:
            self.time += 1
            progressed = False
            order = [r for r in procs if r not in done]
            self.rng.shuffle(order)
            for r in order:
                st = state.get(r)
                if st and st[0] == "work" and self.time < st[1]:
                    progressed = True
                    continue
                if st and st[0] == "until" and not st[1]():
                    continue
                try:
                    nxt = next(procs[r])
                    state[r] = ("work", self.time + nxt[1]) if nxt[0] == "work" else nxt
                    progressed = True
                except StopIteration:
                    done.add(r)
                    progressed = True
            if not progressed:
                raise RuntimeError("deadlock: " + ", ".join(f"rank {r} waits for {state[r][2]}" for r in order if state.get(r)))
        return self.time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=300)
    a = ap.parse_args()
    for seed in range(a.seeds):
        for cp in (2, 4, 8):
            Sim(cp, 6, seed).run()
    print(f"ok: {a.seeds} random speed assignments x cp 2 / 4 / 8, 6 layers: every read saw its epoch, nobody hung")


if __name__ == "__main__":
    main()
