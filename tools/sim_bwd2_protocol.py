"""Discrete-event check of the mbarrier protocol of attn_bwd2_kernel (csrc/attn_bwd.cu) - no GPU needed.

The four roles (TMA producer, tcgen05 issuer, two element-wise warpgroups) are written here as Python generators
that perform the same sequence of barrier operations, with the same phase-parity expressions, as the CUDA code.
Barriers follow mbarrier semantics: `arrive` decrements the pending count of the current phase, completing it (and
re-arming the count) at zero; `wait(parity)` passes iff the phase with that parity has completed, i.e. the current
phase parity differs - so a waiter that falls two phases behind blocks forever, exactly the bug class to catch.
Asynchronous completions (TMA transaction bytes, tcgen05.commit) are modelled as arrivals delivered after a random
delay, in issue order per role for commits.  The scheduler picks runnable roles at random; the run fails on
deadlock or if a role observes an operand that is not ready (e.g. MMA2 before the probabilities were written).

    python tools/sim_bwd2_protocol.py [--seeds 200]
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
    def __init__(self, items, seed):
        self.rng = random.Random(seed)
        self.items = items                      # list of nvis (visible 128-row streamed tiles) per work item
        B = Bar
        self.b = {"res_full": B("res_full", 1), "res_empty": B("res_empty", 2), "acc_full": B("acc_full", 1),
                  "acc_empty": B("acc_empty", 2)}          # warpgroups modelled as 1 arriving agent each
        for i in range(2):
            self.b[f"str_full{i}"] = B(f"str_full{i}", 1)
            self.b[f"str_empty{i}"] = B(f"str_empty{i}", 1)
            self.b[f"s_full{i}"] = B(f"s_full{i}", 1)
            self.b[f"p_full{i}"] = B(f"p_full{i}", 1)
        self.async_q = []                       # (due_time, barrier name, check callback)
        self.time = 0
        # data-readiness bookkeeping for assertions
        self.stage_fill = [-1, -1]              # which stream index currently sits in each smem stage
        self.s_ready = {}                       # (stream index n, half) -> S computed
        self.p_ready = {}                       # (n, half) -> P written
        self.resident_item = -1

    # -- helpers used by the role generators --
    def later(self, bar, lo=1, hi=6, cb=None):
        self.async_q.append((self.time + self.rng.randint(lo, hi), bar, cb))

    def producer(self):
        n_stream = 0
        for k, nvis in enumerate(self.items):
            yield ("wait", "res_empty", (k & 1) ^ 1)
            self.later("res_full", cb=lambda k=k: setattr(self, "resident_item", k))
            for _ in range(nvis):
                st = n_stream & 1
                yield ("wait", f"str_empty{st}", ((n_stream >> 1) & 1) ^ 1)
                self.later(f"str_full{st}", cb=lambda st=st, n=n_stream: self.stage_fill.__setitem__(st, n))
                n_stream += 1

    def issuer(self):
        n_stream = 0
        pcnt = [0, 0]
        commit_time = [0]

        def commit(bar):                          # commits complete in issue order
            commit_time[0] = max(commit_time[0], self.time) + self.rng.randint(1, 5)
            self.async_q.append((commit_time[0], bar, None))

        for k, nvis in enumerate(self.items):
            J = 2 * nvis
            yield ("wait", "res_full", k & 1)
            assert self.resident_item == k, "MMA on a stale resident tile"
            if J == 0:
                self.b["acc_full"].arrive()
                continue

            def mma1(j):
                n, h = n_stream + (j >> 1), j & 1
                st = n & 1
                if h == 0:
                    yield ("wait", f"str_full{st}", (n >> 1) & 1)
                assert self.stage_fill[st] == n, f"MMA1 reads stage {st} holding {self.stage_fill[st]}, wants {n}"
                self.s_ready[(n, h)] = True
                commit(f"s_full{h}")

            yield from mma1(0)
            for j in range(J):
                if j + 1 < J:
                    yield from mma1(j + 1)
                n, h = n_stream + (j >> 1), j & 1
                st = n & 1
                yield ("wait", f"p_full{h}", pcnt[h] & 1)
                pcnt[h] += 1
                if j == 0:
                    yield ("wait", "acc_empty", (k & 1) ^ 1)
                assert self.p_ready.get((n, h)), f"MMA2 before P of tile {n} half {h}"
                assert self.stage_fill[st] == n, "MMA2 reads an overwritten stage"
                if h == 1:
                    commit(f"str_empty{st}")
                if j == J - 1:
                    commit("acc_full")
            n_stream += nvis

    def warpgroup(self, b):
        n_stream, scnt = 0, 0
        for k, nvis in enumerate(self.items):
            for i in range(nvis):
                n = n_stream + i
                st = n & 1
                yield ("wait", f"str_full{st}", (n >> 1) & 1)
                yield ("wait", f"s_full{b}", scnt & 1)
                scnt += 1
                assert self.s_ready.get((n, b)), f"warpgroup {b} reads S of tile {n} before MMA1"
                assert self.stage_fill[st] == n, "column statistics of another tile"
                yield ("work", self.rng.randint(1, 8))
                self.p_ready[(n, b)] = True
                self.b[f"p_full{b}"].arrive()
            n_stream += nvis
            yield ("wait", "acc_full", k & 1)
            yield ("work", self.rng.randint(1, 4))
            self.b["acc_empty"].arrive()
            yield ("work", self.rng.randint(1, 4))   # TMA store reads the staging tile
            self.b["res_empty"].arrive()

    def run(self):
        roles = {"producer": self.producer(), "issuer": self.issuer(), "wg0": self.warpgroup(0), "wg1": self.warpgroup(1)}
        blocked = {}                                # role -> ("wait", bar, parity) | ("work", until)
        done = set()
        while len(done) < len(roles):
            self.time += 1
            due = [a for a in self.async_q if a[0] <= self.time]
            self.async_q = [a for a in self.async_q if a[0] > self.time]
            for _, bar, cb in sorted(due, key=lambda a: a[0]):
                if cb:
                    cb()
                self.b[bar].arrive()
            names = [r for r in roles if r not in done]
            self.rng.shuffle(names)
            progressed = False
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
            if not progressed and not self.async_q:
                raise RuntimeError(f"deadlock at t={self.time}: " + ", ".join(f"{r} on {blocked.get(r)}" for r in names)
                                   + " | phases " + str({k: v.phase for k, v in self.b.items()}))
        return self.time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=200)
    a = ap.parse_args()
    for seed in range(a.seeds):
        rng = random.Random(1000 + seed)
        items = [rng.choice([0, 1, 1, 2, 3, 5, 8]) for _ in range(rng.randint(1, 7))]
        Sim(items, seed).run()
    print(f"ok: {a.seeds} random schedules, no deadlock, no stale operand")


if __name__ == "__main__":
    main()
