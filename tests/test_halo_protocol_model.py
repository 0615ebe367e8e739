"""CPU tier: explicit-state model check of the halo arrival protocol (quda_b200/csrc/kernels.cuh::pack_block /
wait_for_halo).  The host twin runs pack and Dslash one after the other, so it cannot see an ordering bug;
this model enumerates EVERY interleaving of the device-side events of two neighbouring ranks over several back-to-back
exchanges and checks that a boundary role never reads a ghost chunk that is not the neighbour's data of exactly this exchange
(no stale face, no face of a later exchange) and that nobody waits forever.

Modelled per rank and exchange k = 1, 2, ... (program order = stream order of the operator layer):
  pack(k)      n_src x C CTAs, concurrent: write own chunk into the NEIGHBOUR's ghost buffer (k & 1); then the ticket: the CTA
               that draws ticket n_src*C - 1 resets the tickets and publishes count = ((k + (k & 1)) / 2) * C into the neighbour's
               flag word of that buffer (one store -- the batched pack publishes the same value once for all sources)
  boundary(k)  waits until its own flag word of buffer (k & 1) has reached that count, then reads every chunk of that buffer
  pack(k + 1)  only after boundary(k) (stream order / the fork-join of the two-stream schedule)
Negative controls: one ghost buffer instead of two, and a batched pack that publishes after the first source, must both be
caught."""
import sys

import pytest


def explore(K, C, n_src=1, buffers=2, publish_after=None):
    """DFS over all interleavings.  Returns None if every execution is correct, else a description of the first violation."""
    n_cta = n_src * C
    publish_after = n_cta if publish_after is None else publish_after

    def uses(k):
        return (k + (k & 1)) // 2 if buffers == 2 else k

    def buf(k):
        return k & 1 if buffers == 2 else 0

    # state of a rank: (k, cta_stage tuple (0 = to write, 1 = written / to ticket, 2 = done), phase) with phase 0 = packing,
    # 1 = waiting, 2.. = reading chunk (phase - 2) of every source slab, then next exchange
    def initial():
        rank = (1, (0,) * n_cta, 0)
        ghost = tuple((0,) * n_cta for _ in range(buffers))   # per buffer: exchange number whose data sits in each chunk
        return (rank, rank, (ghost, ghost), ((0,) * buffers, (0,) * buffers), ((0,) * buffers, (0,) * buffers))

    seen = set()
    stack = [initial()]
    sys.setrecursionlimit(10000)
    while stack:
        st = stack.pop()
        if st in seen:
            continue
        seen.add(st)
        ranks, ghosts, flags, tickets = [st[0], st[1]], list(st[2]), list(st[3]), list(st[4])
        moves = 0
        finished = all(r[0] > K for r in ranks)
        if finished:
            continue
        for me in (0, 1):
            k, ctas, phase = ranks[me]
            if k > K:
                continue
            peer = 1 - me
            b = buf(k)

            def push(new_rank, new_ghost_peer=None, new_flags_peer=None, new_tickets=None):
                r2 = list(ranks)
                r2[me] = new_rank
                g2, f2, t2 = list(ghosts), list(flags), list(tickets)
                if new_ghost_peer is not None:
                    g2[peer] = new_ghost_peer
                if new_flags_peer is not None:
                    f2[peer] = new_flags_peer
                if new_tickets is not None:
                    t2[me] = new_tickets
                stack.append((r2[0], r2[1], tuple(g2), tuple(f2), tuple(t2)))

            if phase == 0:  # pack: any CTA may take its next step
                for c in range(n_cta):
                    if ctas[c] == 0:      # write the chunk into the neighbour's buffer
                        g = list(ghosts[peer])
                        gb = list(g[b])
                        gb[c] = k
                        g[b] = tuple(gb)
                        push((k, ctas[:c] + (1,) + ctas[c + 1:], 0), new_ghost_peer=tuple(g))
                        moves += 1
                    elif ctas[c] == 1:    # fence + ticket (+ publish by the last one)
                        t = list(tickets[me])
                        prev = t[b]
                        t[b] = prev + 1
                        nf = None
                        if prev == publish_after - 1:
                            f = list(flags[peer])
                            f[b] = uses(k) * C
                            nf = tuple(f)
                        if prev == n_cta - 1:
                            t[b] = 0
                        nc = ctas[:c] + (2,) + ctas[c + 1:]
                        push((k, nc, 1 if all(x == 2 for x in nc) else 0), new_flags_peer=nf, new_tickets=tuple(t))
                        moves += 1
            elif phase == 1:  # boundary: acquire the arrival counter
                if flags[me][b] >= uses(k) * C:
                    push((k, ctas, 2))
                    moves += 1
            else:  # boundary: read chunk (phase - 2)
                c = phase - 2
                if ghosts[me][b][c] != k:
                    return f"rank {me}, exchange {k}: chunk {c} of buffer {b} holds exchange {ghosts[me][b][c]}"
                if c + 1 < n_cta:
                    push((k, ctas, phase + 1))
                else:
                    push((k + 1, (0,) * n_cta, 0))
                moves += 1
        if moves == 0:
            return f"deadlock in state {st[:2]} flags {flags}"
    return None


@pytest.mark.parametrize("K,C,n_src", [(5, 2, 1), (4, 3, 1), (4, 1, 3), (3, 2, 2)])
def test_protocol_is_race_and_deadlock_free(K, C, n_src):
    assert explore(K, C, n_src) is None


def test_model_catches_a_single_ghost_buffer():
    """without double buffering a fast rank overwrites the face its neighbour is still reading"""
    v = explore(3, 2, buffers=1)
    assert v is not None and "holds exchange" in v


def test_model_catches_an_early_batch_signal():
    """a batched pack that publishes after the first source's CTAs lets the boundary role read slabs that have not landed"""
    v = explore(2, 2, n_src=2, publish_after=2)
    assert v is not None and "holds exchange" in v


# ------------------------------------------------------------------------------------------------------------------
# The in-kernel NVLink all-reduce of the solver's scalars (quda_b200/csrc/host/dirac.cu::finish_reduction): thread t of the
# finishing block of rank r stores its partial sum into slot [k & 1][r] of rank t's mailbox (value, then the sequence tag with
# release semantics), then spins on its OWN slot [k & 1][t] until the tag has reached k (acquire) and reads the value; the
# next reduction is a later launch, i.e. starts when all threads of this one are done.
def explore_allreduce(n_ranks, K, slots=2, tag_first=False):
    def slot(k):
        return k & 1 if slots == 2 else 0

    # rank state: (k, per-thread stage) with stage 0 = store value, 1 = store tag, 2 = wait for the tag, 3 = read, 4 = done
    init_rank = (1, (0,) * n_ranks)
    box0 = tuple(tuple((0, 0) for _ in range(n_ranks)) for _ in range(slots))  # [slot][source] = (value's reduction, tag)
    start = (tuple(init_rank for _ in range(n_ranks)), tuple(box0 for _ in range(n_ranks)))
    seen, stack = set(), [start]
    while stack:
        st = stack.pop()
        if st in seen:
            continue
        seen.add(st)
        ranks, boxes = st
        if all(r[0] > K for r in ranks):
            continue
        moves = 0
        for me in range(n_ranks):
            k, stages = ranks[me]
            if k > K:
                continue
            b = slot(k)
            for t in range(n_ranks):
                sg = stages[t]
                if sg == 4:
                    continue
                nb = boxes
                if sg in (0, 1):  # the two stores into rank t's box, slot [b][me]
                    write_value = (sg == 0) != tag_first
                    box = [list(s) for s in boxes[t]]
                    v, tag = box[b][me]
                    box[b][me] = (k, tag) if write_value else (v, k)
                    nb = boxes[:t] + (tuple(tuple(s) for s in box),) + boxes[t + 1:]
                elif sg == 2:
                    if boxes[me][b][t][1] < k:
                        continue  # still spinning
                else:  # read
                    if boxes[me][b][t][0] != k:
                        return f"rank {me}, reduction {k}: slot of rank {t} holds the value of reduction {boxes[me][b][t][0]}"
                ns = stages[:t] + (sg + 1,) + stages[t + 1:]
                nr = (k + 1, (0,) * n_ranks) if all(x == 4 for x in ns) else (k, ns)
                stack.append((ranks[:me] + (nr,) + ranks[me + 1:], nb))
                moves += 1
        if moves == 0:
            return f"deadlock: {ranks}"
    return None


@pytest.mark.parametrize("n_ranks,K", [(2, 6), (3, 1)])
def test_mailbox_allreduce_is_race_and_deadlock_free(n_ranks, K):
    assert explore_allreduce(n_ranks, K) is None


def test_model_catches_single_slot_mailboxes_and_tag_before_value():
    assert "holds the value" in (explore_allreduce(2, 3, slots=1) or "")       # a fast rank overwrites the sum being read
    assert "holds the value" in (explore_allreduce(2, 2, tag_first=True) or "")  # publishing the tag before the value
