"""CPU: exhaustive interleaving check of the buffer protocol behind the peer-memory collectives
(stochqn_b200/csrc/ext_impl.inc: push all-gather, pull reduce-scatter).

Every rank runs, per call k:   produce(k) into a buffer of parity k & 1   ->   rank barrier k   ->   consume(k)
  pull reduce-scatter: produce = the gradient product writes this rank's send vector; consume = the pull kernel reads
                       EVERY rank's send vector of that parity
  push all-gather    : produce = the push kernel writes into EVERY rank's gathered vector; consume = the next kernels of
                       this rank read its own gathered vector
The claim in the source: with two buffers (parity of the call count) a producer of call k+2 can never overlap a consumer
of call k, because a rank reaches produce(k+2) only after passing barrier k+1, which every rank reaches after its
consume(k) (stream order).  The model below explores ALL interleavings of the ranks' begin / end events and reports an
overlap of a write and a read of the same buffer; it must find none with two buffers and must find one with a single
buffer (which also shows that the checker can see the hazard)."""
import itertools

import pytest


def _program(calls):
    ops = []
    for k in range(calls):
        ops += [("pb", k), ("pe", k), ("arrive", k), ("pass", k), ("cb", k), ("ce", k)]
    return ops


def _explore(world, calls, nbuf, push):
    """DFS over the joint program counters.  Returns a description of the first hazard, or None."""
    prog = _program(calls)
    start = tuple([0] * world)
    seen = {start}
    stack = [start]

    def active(pcs):
        """(kind, rank, call) of every produce / consume that has begun and not ended"""
        out = []
        for r, pc in enumerate(pcs):
            if pc == 0:
                continue
            op, k = prog[pc - 1]
            if op == "pb":
                out.append(("p", r, k))
            elif op == "cb":
                out.append(("c", r, k))
        return out

    def touches(kind, r, k):
        """set of (owner rank, buffer index) the activity accesses"""
        b = k % nbuf
        if push:
            return {(q, b) for q in range(world)} if kind == "p" else {(r, b)}
        return {(r, b)} if kind == "p" else {(q, b) for q in range(world)}

    while stack:
        pcs = stack.pop()
        act = active(pcs)
        for (ka, ra, ca), (kb, rb, cb) in itertools.combinations(act, 2):
            if ka == kb == "c":
                continue                                   # two readers
            if ka == kb == "p" and ca == cb and not push:
                continue                                   # producers of the same call write their own vectors
            if ka == kb == "p" and push and ca == cb:
                continue                                   # push: every rank writes its OWN block of the peers' vectors
            if touches(ka, ra, ca) & touches(kb, rb, cb):
                if ca == cb and {ka, kb} == {"p", "c"}:
                    return "call %d consumed while it is still produced (ranks %d, %d)" % (ca, ra, rb)
                if ca != cb:
                    return "%s of call %d (rank %d) overlaps %s of call %d (rank %d)" % (ka, ca, ra, kb, cb, rb)
        for r in range(world):
            pc = pcs[r]
            if pc == len(prog):
                continue
            op, k = prog[pc]
            if op == "pass":                               # enabled once every rank has arrived at barrier k
                if not all(pcs[q] > pc - 1 for q in range(world)):      # prog[pc - 1] is ("arrive", k)
                    continue
            nxt = pcs[:r] + (pc + 1,) + pcs[r + 1:]
            if nxt not in seen:
                seen.add(nxt)
                stack.append(nxt)
    assert tuple([len(prog)] * world) in seen, "the model deadlocked"
    return None


@pytest.mark.parametrize("push", [False, True], ids=["pull_reduce_scatter", "push_all_gather"])
@pytest.mark.parametrize("world", [2, 3])
def test_two_buffers_never_overlap(world, push):
    assert _explore(world, 4, 2, push) is None


@pytest.mark.parametrize("push", [False, True], ids=["pull_reduce_scatter", "push_all_gather"])
def test_one_buffer_is_caught(push):
    hazard = _explore(2, 3, 1, push)
    assert hazard is not None and "overlaps" in hazard, hazard


# ---- the mailbox all-reduce itself (stochqn_b200/csrc/p2p.cuh) ------------------------------------------------------------------
# exchange k on every rank:  write my record into slot [parity][me] of EVERY rank's mailbox  ->  raise flag [parity][me] = k + 1
# in every mailbox (release)  ->  wait until all flags [parity][*] of MY mailbox are >= k + 1 (acquire)  ->  read the records
# [parity][*] of my mailbox.  Claim (p2p.cuh:13-14): a rank can only start exchange k + 2 (same parity as k) after every peer
# has posted its flag for k + 1, i.e. after every peer has finished reading k.
def _explore_mailbox(world, exchanges, nbuf):
    ops = []
    for k in range(exchanges):
        ops += [("wb", k), ("we", k), ("flag", k), ("wait", k), ("rb", k), ("re", k)]
    idx = {op: i for i, op in enumerate(ops)}
    start = tuple([0] * world)
    seen, stack = {start}, [start]

    def flag_value(writer_pc, parity):
        """what the writer has last stored into its flag of that parity (in any mailbox)"""
        v = 0
        for k in range(exchanges):
            if k % nbuf == parity and writer_pc > idx[("flag", k)]:
                v = k + 1
        return v

    while stack:
        pcs = stack.pop()
        for b, pc in enumerate(pcs):                      # is a rank reading exchange k while a writer is already past wb(k')?
            if pc == 0 or ops[pc - 1][0] != "rb":
                continue
            k = ops[pc - 1][1]
            for a in range(world):
                later = [j for j in range(k + 1, exchanges) if j % nbuf == k % nbuf]
                if later and pcs[a] > idx[("wb", later[0])]:
                    return "rank %d overwrites its slot for exchange %d while rank %d reads exchange %d" % (a, later[0], b, k)
                if pcs[a] <= idx[("we", k)]:
                    return "rank %d reads exchange %d before rank %d has written it" % (b, k, a)
        for r in range(world):
            pc = pcs[r]
            if pc == len(ops):
                continue
            op, k = ops[pc]
            if op == "wait" and not all(flag_value(pcs[a], k % nbuf) >= k + 1 for a in range(world)):
                continue
            nxt = pcs[:r] + (pc + 1,) + pcs[r + 1:]
            if nxt not in seen:
                seen.add(nxt)
                stack.append(nxt)
    assert tuple([len(ops)] * world) in seen, "the model deadlocked"
    return None


@pytest.mark.parametrize("world", [2, 3])
def test_mailbox_double_buffer_is_safe(world):
    assert _explore_mailbox(world, 5, 2) is None


def test_mailbox_single_buffer_is_caught():
    hazard = _explore_mailbox(2, 3, 1)
    assert hazard is not None and "overwrites" in hazard, hazard
