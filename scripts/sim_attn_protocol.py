#!/usr/bin/env python
"""Discrete-event model of the barrier protocol of attn_tc_fwd_kernel (csrc/attention_tc.cu): the Q/K loader, the V loader, the
MMA warp and the two softmax warpgroups as coroutines over mbarriers with phase parity and (for the staggered variant)
the two named barriers.  It checks, for every number of work items per CTA and for the three variants (default, V2 = O in
its own TMEM columns, V2 + token-staggered groups), that

  * nobody waits forever (the kernel's bounded mbar_wait would trap),
  * no resource is overwritten while its previous contents are still needed:
      Q/K smem of a set  (written by TMA, read by the score MMA),
      V smem of a set    (written by TMA, read by the PV MMA),
      P smem of a set    (written by the softmax group, read by the PV MMA),
      S TMEM columns     (written by the score MMA, read by the softmax passes),
      O TMEM columns     (written by the PV MMA, read by the epilogue; in the default variant they alias S),
  * with the staggered variant, the two groups are never inside pass 2 at the same time.

It models ordering only (an async MMA "completes" some scheduler steps after it is issued, in issue order; TMA loads land
in any order), not time.  The
scheduler picks runnable coroutines in a seeded random order, so many interleavings are exercised.

    python scripts/sim_attn_protocol.py            (exit code 0 = all cases pass)
"""
import random
import sys


class Deadlock(Exception):
    pass


class MBar:
    """mbarrier with an arrival count; wait(parity) passes when the phase with that parity has completed"""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0  # phase = number of completed phases

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passed(self, parity):
        # try_wait.parity(P) is true iff the barrier's current (incomplete) phase has parity != P,
        # i.e. the last completed phase has parity P ... with a fresh barrier passing for P = 1
        return (self.phase & 1) != parity


class Named:
    """bar.sync / bar.arrive with a thread count of 2 groups: completes when one group arrived and one group syncs"""

    def __init__(self):
        self.arrived = 0  # arrivals not yet consumed by a sync

    def arrive(self):
        self.arrived += 1

    def can_sync(self):
        return self.arrived > 0

    def sync(self):
        self.arrived -= 1


def simulate(n_items, v2, stag, seed):
    rnd = random.Random(seed)
    # barriers per set: 0 qk_full, 1 v_free, 2 s_full, 3 p_full, 4 o_full, 5 tmem_free, 6 v_full
    bar = [[MBar(1), MBar(1), MBar(1), MBar(1), MBar(1), MBar(1), MBar(1)] for _ in range(2)]
    named = {1: Named(), 2: Named()}
    # resource state: which item currently owns it and whether its readers are done
    res = {}

    def write(name, item):
        st = res.get(name)
        assert st is None or st["free"], f"{name}: item {item} overwrites item {st['item']} before it was consumed"
        res[name] = {"item": item, "free": False, "landed": not name.startswith(("QK", "V"))}  # TMA data lands later

    def read(name, item):
        st = res.get(name)
        assert st is not None and st["item"] == item and st["landed"], f"{name}: item {item} reads {st}"

    def release(name, item):
        st = res[name]
        assert st["item"] == item
        st["free"] = True

    pending_async = []  # (kind, payload) completed later in issue order: models tcgen05.commit / TMA completion
    in_pass2 = set()

    def wait(b, parity):
        while not b.passed(parity):
            yield "blocked"

    def qk_loader():
        for i in range(n_items):
            s, k = i & 1, i >> 1
            if k > 0:
                yield from wait(bar[s][2], (k - 1) & 1)
            write(f"QK{s}", i)
            pending_async.append(("landed", (f"QK{s}", bar[s][0])))
            yield "step"

    def v_loader():
        for i in range(n_items):
            s, k = i & 1, i >> 1
            yield from wait(bar[s][1], (k & 1) ^ 1)
            write(f"V{s}", i)
            pending_async.append(("landed", (f"V{s}", bar[s][6])))
            yield "step"

    def mma():
        for i in range(n_items + 1):
            if i < n_items:
                s, k = i & 1, i >> 1
                yield from wait(bar[s][0], k & 1)
                if not v2:
                    yield from wait(bar[s][5], (k & 1) ^ 1)
                read(f"QK{s}", i)
                write(f"S{s}", i)  # in the default variant O aliases S: writing S needs O consumed too
                if not v2:
                    st = res.get(f"O{s}")
                    assert st is None or st["free"], f"S{s} (aliasing O{s}) written for item {i} before O of item {st['item']} was read"
                pending_async.append(("smma_done", (s, i)))
                yield "step"
            if i >= 1:
                j = i - 1
                s, k = j & 1, j >> 1
                yield from wait(bar[s][3], k & 1)
                yield from wait(bar[s][6], k & 1)
                if v2:
                    yield from wait(bar[s][5], (k & 1) ^ 1)
                read(f"P{s}", j)
                read(f"V{s}", j)
                write(f"O{s}", j)
                if not v2:  # O lands in the S columns: S of this item must have been consumed (it was: p_full)
                    assert res[f"S{s}"]["item"] == j and res[f"S{s}"]["free"]
                pending_async.append(("pv_done", (s, j)))
                yield "step"

    def softmax(g):
        n0, n1 = (n_items + 1) >> 1, n_items >> 1
        for i in range(g, n_items, 2):
            s, k = g, i >> 1
            yield from wait(bar[s][2], k & 1)
            read(f"S{s}", i)  # pass 1
            yield "step"
            if stag:
                nb = named[1] if g == 0 else named[2]
                if g == 1 or k > 0:
                    while not nb.can_sync():
                        yield "blocked"
                    nb.sync()
            in_pass2.add(g)
            assert not (stag and len(in_pass2) == 2), "both groups in pass 2 under the staggered variant"
            read(f"S{s}", i)  # pass 2
            write(f"P{s}", i)
            yield "step"
            in_pass2.discard(g)
            if stag:
                if g == 0:
                    if k < n1:
                        named[2].arrive()
                else:
                    if k + 1 < n0:
                        named[1].arrive()
            release(f"S{s}", i)
            bar[s][3].arrive()  # p_full (one arrival stands for the group's 128)
            yield from wait(bar[s][4], k & 1)
            read(f"O{s}", i)
            release(f"O{s}", i)
            bar[s][5].arrive()  # tmem_free
            yield "step"

    procs = {"qk": qk_loader(), "v": v_loader(), "mma": mma(), "sm0": softmax(0), "sm1": softmax(1)}
    alive = dict(procs)
    idle_rounds = 0
    while alive:
        progressed = False
        names = list(alive)
        rnd.shuffle(names)
        for nm in names:
            try:
                r = next(alive[nm])
            except StopIteration:
                del alive[nm]
                progressed = True
                continue
            if r == "step":
                progressed = True
        # complete an async operation now and then: tensor-pipe operations in issue order, TMA loads in any order
        if pending_async and (not progressed or rnd.random() < 0.5):
            first_mma = next((n for n, (kd, _) in enumerate(pending_async) if kd != "landed"), None)
            cand = [n for n, (kd, _) in enumerate(pending_async) if kd == "landed"] + ([first_mma] if first_mma is not None else [])
            kind, payload = pending_async.pop(rnd.choice(cand))
            progressed = True
            if kind == "landed":
                res[payload[0]]["landed"] = True
                payload[1].arrive()
            elif kind == "smma_done":
                s, i = payload
                release(f"QK{s}", i)
                bar[s][2].arrive()
            elif kind == "pv_done":
                s, j = payload
                release(f"P{s}", j)
                release(f"V{s}", j)
                bar[s][4].arrive()
                bar[s][1].arrive()
        idle_rounds = 0 if progressed else idle_rounds + 1
        if idle_rounds > 3:
            raise Deadlock(f"stuck with {sorted(alive)} alive")
    for nb in named.values():
        assert nb.arrived == 0, "dangling named-barrier arrival at kernel exit"


def main():
    bad = 0
    for v2, stag, name in ((False, False, "default"), (True, False, "V2"), (True, True, "V2+STAG")):
        for n in range(0, 12):
            for seed in range(40):
                try:
                    simulate(n, v2, stag, seed)
                except (AssertionError, Deadlock) as e:
                    print(f"FAIL {name} items={n} seed={seed}: {type(e).__name__}: {e}")
                    bad += 1
                    break
        print(f"{name}: checked 0..11 items per CTA x 40 interleavings")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
