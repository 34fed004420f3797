"""Model check of the ring window's flag protocol (engine.cu rotate_part / wait_part / the lane loop of train_impl).

The protocol orders copy-engine writes into a neighbour's memory by two monotonic flags per item sub-shard (`arrival`, `credit`).
Hardware with 3+ GPUs was scarce in round 2 (on 2 GPUs the member that sends to us and the member we send to are the same
process, which hides direction mistakes), so the rule itself is checked here: a discrete-event model of G members, each with its
lane streams and one copy stream, executes the exact operation sequence the engine enqueues under random interleavings and
random operation durations, and asserts
  * every launch of sub-epoch T on member g reads item group (g + T) mod G, slice by slice, in the version its previous
    holder left it (no stale slice, no slice from the future),
  * no hand-over writes into a buffer slice that a launch or an outgoing copy is still reading, or whose content has not been
    handed on yet,
  * the model neither deadlocks nor ends with a flag short of its final value (what train_impl's drain waits for).
Mutations of the rule (no credit wait; credit written to the wrong neighbour; arrival before the copy) must be caught, so the
model is known to be able to fail."""
import random

import pytest


class Violation(Exception):
    pass


def simulate(G, parts, sub_epochs, seed, lanes=2, mutate=None):
    rng = random.Random(seed)
    # memory: buf[g][b][p] = (group, version) held by slice p of Q buffer b of member g; version = sub-epochs trained on it
    buf = [[[None] * parts for _ in range(2)] for _ in range(G)]
    for g in range(G):
        for p in range(parts):
            buf[g][0][p] = (g, 0)                                   # member g starts holding group g in buffer 0
    readers = [[[0] * parts for _ in range(2)] for _ in range(G)]  # active readers of a slice (launches, outgoing copies)
    writers = [[[0] * parts for _ in range(2)] for _ in range(G)]
    sent_on = [[[True] * parts for _ in range(2)] for _ in range(G)]   # content already handed on (buffer 1 starts free)
    for g in range(G):
        for p in range(parts):
            sent_on[g][0][p] = False
    arrival = [[0] * parts for _ in range(G)]
    credit = [[0] * parts for _ in range(G)]
    done = [[[False] * parts for _ in range(sub_epochs)] for _ in range(G)]     # ev_part_done, per sub-epoch

    # ---- the operation lists the engine enqueues (engine.cu: lane loop 2231-2246, rotate_part 1771-1793) ----
    streams = {}
    for g in range(G):
        to, frm = (g - 1) % G, (g + 1) % G
        if mutate == "credit_to_wrong_neighbour":
            frm = to
        for lane in range(lanes):
            streams[(g, "lane", lane)] = []
        streams[(g, "copy", 0)] = []
        for T in range(sub_epochs):
            cur = T % 2
            for p in range(parts):
                ln = streams[(g, "lane", p % lanes)]
                if T > 0:
                    ln.append(("wait", "arrival", g, p, T))                     # wait_part: arrival[p] >= sent[p] (= T)
                ln.append(("launch", g, T, p, cur))
                cp = streams[(g, "copy", 0)]
                cp.append(("wait_done", g, T, p))
                if T > 0 and mutate != "no_credit_wait":
                    cp.append(("wait", "credit", g, p, T))                      # n > 1: credit[p] >= n - 1
                if mutate == "arrival_before_copy":
                    cp.append(("signal", "arrival", to, p, T + 1))
                cp.append(("copy", g, to, p, cur, cur ^ 1, T))
                if mutate != "arrival_before_copy":
                    cp.append(("signal", "arrival", to, p, T + 1))
                cp.append(("signal", "credit", frm, p, T + 1))
    head = {s: 0 for s in streams}
    busy = {}            # stream -> (end action) of an operation with a duration that has begun

    def runnable(s):
        if s in busy:
            return True
        ops = streams[s]
        if head[s] >= len(ops):
            return False
        op = ops[head[s]]
        if op[0] == "wait":
            flags = arrival if op[1] == "arrival" else credit
            return flags[op[2]][op[3]] >= op[4]
        if op[0] == "wait_done":
            return done[op[1]][op[2]][op[3]]
        return True

    steps = 0
    while True:
        ready = [s for s in streams if runnable(s)]
        if not ready:
            break
        s = rng.choice(ready)
        steps += 1
        if s in busy:                                   # the operation in flight on this stream ends
            busy.pop(s)()
            head[s] += 1
            continue
        op = streams[s][head[s]]
        if op[0] in ("wait", "wait_done"):
            head[s] += 1
        elif op[0] == "signal":
            flags = arrival if op[1] == "arrival" else credit
            if op[4] != flags[op[2]][op[3]] + 1:
                raise Violation("flag %s[%d][%d] jumps %d -> %d" % (op[1], op[2], op[3], flags[op[2]][op[3]], op[4]))
            flags[op[2]][op[3]] = op[4]
            head[s] += 1
        elif op[0] == "launch":
            _, g, T, p, b = op
            want = ((g + T) % G, T)
            if buf[g][b][p] != want:
                raise Violation("member %d sub-epoch %d slice %d reads %r, wants %r" % (g, T, p, buf[g][b][p], want))
            if writers[g][b][p]:
                raise Violation("member %d sub-epoch %d slice %d launched under an incoming copy" % (g, T, p))
            readers[g][b][p] += 1

            def end(g=g, T=T, p=p, b=b):
                readers[g][b][p] -= 1
                grp, ver = buf[g][b][p]
                buf[g][b][p] = (grp, ver + 1)
                done[g][T][p] = True
            busy[s] = end
        elif op[0] == "copy":
            _, g, to, p, sb, db, T = op
            if readers[to][db][p] or writers[to][db][p]:
                raise Violation("hand-over %d of slice %d: member %d writes member %d's buffer %d while it is in use" % (T + 1, p, g, to, db))
            if not sent_on[to][db][p]:
                raise Violation("hand-over %d of slice %d: member %d overwrites content member %d has not handed on" % (T + 1, p, g, to))
            readers[g][sb][p] += 1
            writers[to][db][p] += 1

            def end(g=g, to=to, p=p, sb=sb, db=db):
                readers[g][sb][p] -= 1
                writers[to][db][p] -= 1
                buf[to][db][p] = buf[g][sb][p]
                sent_on[to][db][p] = False
                sent_on[g][sb][p] = True
            busy[s] = end
        else:
            raise AssertionError(op)
    for s, ops in streams.items():
        if head[s] != len(ops):
            raise Violation("deadlock: stream %r stopped at %r" % (s, ops[head[s]]))
    for g in range(G):
        for p in range(parts):
            if arrival[g][p] != sub_epochs or credit[g][p] != sub_epochs:
                raise Violation("flags of member %d slice %d end at %d / %d" % (g, p, arrival[g][p], credit[g][p]))
            b = sub_epochs % 2
            if buf[g][b][p] != ((g + sub_epochs) % G, sub_epochs):
                raise Violation("member %d ends holding %r" % (g, buf[g][b][p]))
    return steps


@pytest.mark.parametrize("G", [2, 3, 4, 8])
@pytest.mark.parametrize("parts", [1, 2, 4])
def test_ring_window_protocol_holds_under_random_interleavings(G, parts):
    for seed in range(40):
        simulate(G, parts, sub_epochs=3 * G, seed=seed)                # three epochs: every group is home again three times


def test_after_whole_epochs_every_group_is_home():
    # (g + T) mod G == g when T is a multiple of G: what get_factors relies on between train calls
    simulate(8, 2, sub_epochs=16, seed=1)


@pytest.mark.parametrize("mutation", ["no_credit_wait", "credit_to_wrong_neighbour", "arrival_before_copy"])
def test_the_model_catches_broken_rules(mutation):
    """Each mutation must be caught on a ring of 4 (credit_to_wrong_neighbour is invisible on a ring of 2, where both
    neighbours are the same member -- the reason the rule is model-checked on larger rings)."""
    caught = 0
    for seed in range(60):
        try:
            simulate(4, 2, sub_epochs=12, seed=seed, mutate=mutation)
        except Violation:
            caught += 1
    assert caught > 0, mutation
    if mutation == "credit_to_wrong_neighbour":
        for seed in range(20):
            simulate(2, 2, sub_epochs=8, seed=seed, mutate=mutation)     # ... and indeed passes on 2
