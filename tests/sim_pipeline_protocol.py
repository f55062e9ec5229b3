"""Discrete-event model of the main kernels' barrier protocol (qlora_gemm.cuh): operand ring, UMMA issuer, asynchronous
tensor pipe with tcgen05.commit semantics, epilogue.  The roles are hand-ported from the kernel with the SAME index / phase arithmetic (stage `s`, phase
bit `ph`, accumulator phase `aph`); a random scheduler interleaves them and the
model checks, over many seeds and shapes:

  * no deadlock (every role runs to completion),
  * every parity wait is satisfied by exactly the phase it meant (no aliasing: the barrier is never a phase ahead),
  * an MMA reads the operand stage that holds ITS k-block of ITS tile,
  * the accumulator sub-tile an MMA writes has been drained of the previous tile, and the epilogue drains a sub-tile
    only after all of its MMAs have executed.

Also modelled, with the same checks: the double-buffered, two-epilogue-set variant (masked dX GEMM), the transform-group
ring of the LoRA-dropout kernels (4 groups over a 6-stage ring), and the decode kernels' operand + packed rings with two
decode groups and LoRA tail k-blocks (the `it` / `pit` position arithmetic).  Since round 2 TMA loads and
shared-memory reads complete asynchronously and out of order in the model, and run_all() ends with mutation checks: the
three protocols that stalled or corrupted on hardware in rounds 1 / 2 (transform groups without the "previous use
released" wait, packed ring without it at tile boundaries, packed slot released before its read returned) must fail.

Pure Python, no GPU: a design check, run by tests/test_host.py.  It does not model TMA, decode or TMEM timing.
"""
import random

STAGES = 4
MT = 2


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.completed = count, count, 0   # completed = number of finished phases

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier expects"
        if self.pending == 0:
            self.pending = self.count
            self.completed += 1

    def passes(self, parity):
        return (self.completed & 1) != parity   # try_wait.parity: the phase with this parity has completed


class Deadlock(Exception):
    pass


def simulate(num_tiles, kb_total, seed):
    rng = random.Random(seed)
    full = [Bar(2) for _ in range(STAGES)]      # producer + decode group
    empty = [Bar(1) for _ in range(STAGES)]     # tcgen05.commit
    tfull = Bar(1)
    tempty = [Bar(1) for _ in range(MT)]        # the epilogue as one agent
    stage_content = [None] * STAGES             # (tile, kb) the producer / decode wrote
    stage_ready = [0] * STAGES                  # writers done for the current content (2 = both)
    pipe = []                                   # tensor-pipe FIFO: ("mma", tile, kb, mt, stage) | ("commit", bar)
    acc_mmas = {}                               # (tile, mt) -> executed k-blocks
    acc_drained = {(-1, 0): True, (-1, 1): True}
    log = {"waits": 0, "issuer_done": False}

    def wait(bar, parity, intended):
        """Generator: block until the barrier's phase `intended` (0-based) has completed, via the parity the kernel uses."""
        while not bar.passes(parity):
            yield "blocked"
        log["waits"] += 1
        assert bar.completed == intended + 1, f"parity alias: wanted phase {intended}, barrier completed {bar.completed}"

    def writer(which):   # which = 0 producer (TMA), 1 decode group: both wait for the slot, fill it, arrive on full
        s, ph, use = 0, 0, [0] * STAGES
        for tile in range(num_tiles):
            for kb in range(kb_total):
                yield from wait(empty[s], ph ^ 1, use[s] - 1) if use[s] > 0 else iter(())
                if use[s] == 0:
                    assert empty[s].passes(ph ^ 1)          # fresh barrier: the kernel's first wait passes immediately
                if which == 0:
                    stage_content[s] = (tile, kb)
                stage_ready[s] += 1
                yield "step"
                full[s].arrive()
                use[s] += 1
                s += 1
                if s == STAGES:
                    s, ph = 0, ph ^ 1

    def issuer():
        s, ph, aph = 0, 0, 0
        use = [0] * STAGES                      # completed uses of each stage so far (for the intended-phase check)

        def issue(stage, mt, tile, kb):
            pipe.append(("mma", tile, kb, mt, stage))

        for tile in range(num_tiles):
            for kb in range(kb_total):
                yield from wait(full[s], ph, use[s])
                for mt in range(MT):
                    if kb == 0:
                        yield from wait(tempty[mt], aph ^ 1, tile - 1) if tile > 0 else iter(())
                    issue(s, mt, tile, kb)
                    yield "step"
                pipe.append(("commit", empty[s]))
                if kb == kb_total - 1:
                    pipe.append(("commit", tfull))
                use[s] += 1
                s += 1
                if s == STAGES:
                    s, ph = 0, ph ^ 1
            aph ^= 1
        log["issuer_done"] = True

    def epilogue():
        aph = 0
        for tile in range(num_tiles):
            yield from wait(tfull, aph, tile)
            for mt in range(MT):
                assert acc_mmas.get((tile, mt), 0) == kb_total, f"drain of tile {tile} sub-tile {mt} before its MMAs finished"
                yield "step"                      # the drain itself takes time
                acc_drained[(tile, mt)] = True
                tempty[mt].arrive()
            aph ^= 1

    def tensor_pipe():
        done = 0
        total = num_tiles * kb_total * MT
        while not log["issuer_done"] or pipe:
            if not pipe:
                yield "blocked"
                continue
            op = pipe.pop(0)
            if op[0] == "commit":
                op[1].arrive()                     # all earlier MMAs have executed (FIFO)
            else:
                _, tile, kb, mt, stage = op
                assert stage_content[stage] == (tile, kb), f"MMA of tile {tile} kb {kb} read stage holding {stage_content[stage]}"
                assert acc_drained.get((tile - 1, mt), False), f"MMA into sub-tile {mt} of tile {tile} before tile {tile - 1} was drained"
                assert acc_mmas.get((tile, mt), 0) == kb, "k-blocks of one accumulator out of order"
                acc_mmas[(tile, mt)] = kb + 1
                done += 1
            yield "step"
        assert done == total

    agents = {"producer": writer(0), "decode": writer(1), "issuer": issuer(), "epilogue": epilogue(), "pipe": tensor_pipe()}
    blocked_rounds = 0
    while agents:
        name = rng.choice(sorted(agents))
        try:
            r = next(agents[name])
        except StopIteration:
            del agents[name]
            blocked_rounds = 0
            continue
        blocked_rounds = blocked_rounds + 1 if r == "blocked" else 0
        if blocked_rounds > 20000:
            raise Deadlock(f"roles still alive: {sorted(agents)} (tiles {num_tiles}, kb {kb_total}, seed {seed})")
    assert all(acc_drained.get((t, m)) for t in range(num_tiles) for m in range(MT))
    return log["waits"]


def simulate_two_sets(num_tiles, kb_total, seed, stages=5):
    """The masked dX GEMM (GemmKNMask): MT = 1, double-buffered accumulator, TWO epilogue warp sets -- set e drains the
    tiles whose sequence number is e (mod 2) from accumulator stage e, with its own phase bit."""
    rng = random.Random(seed)
    full = [Bar(1) for _ in range(stages)]      # TMA producer only (both operands by TMA)
    empty = [Bar(1) for _ in range(stages)]
    tfull = [Bar(1), Bar(1)]
    tempty = [Bar(1), Bar(1)]
    stage_content = [None] * stages
    pipe = []
    acc_mmas, acc_drained = {}, {(-1, 0): True, (-2, 0): True}
    log = {"issuer_done": False}

    def wait(bar, parity, intended):
        while not bar.passes(parity):
            yield "blocked"
        assert bar.completed == intended + 1, f"parity alias: wanted phase {intended}, barrier completed {bar.completed}"

    def producer():
        s, ph, use = 0, 0, [0] * stages
        for tile in range(num_tiles):
            for kb in range(kb_total):
                if use[s] > 0:
                    yield from wait(empty[s], ph ^ 1, use[s] - 1)
                stage_content[s] = (tile, kb)
                yield "step"
                full[s].arrive()
                use[s] += 1
                s += 1
                if s == stages:
                    s, ph = 0, ph ^ 1

    def issuer():
        s, ph, acc, aph, use = 0, 0, 0, 0, [0] * stages
        for tile in range(num_tiles):
            for kb in range(kb_total):
                yield from wait(full[s], ph, use[s])
                if kb == 0 and tile >= 2:
                    yield from wait(tempty[acc], aph ^ 1, tile // 2 - 1)
                pipe.append(("mma", tile, kb, acc, s))
                yield "step"
                pipe.append(("commit", empty[s]))
                if kb == kb_total - 1:
                    pipe.append(("commit", tfull[acc]))
                use[s] += 1
                s += 1
                if s == stages:
                    s, ph = 0, ph ^ 1
            acc += 1
            if acc == 2:
                acc, aph = 0, aph ^ 1
        log["issuer_done"] = True

    def epilogue(eset):
        aph = 0
        for n, tile in enumerate(range(eset, num_tiles, 2)):
            yield from wait(tfull[eset], aph, n)
            assert acc_mmas.get(tile, 0) == kb_total, f"set {eset} drains tile {tile} before its MMAs finished"
            yield "step"
            acc_drained[(tile, 0)] = True
            tempty[eset].arrive()
            aph ^= 1

    def tensor_pipe():
        while not log["issuer_done"] or pipe:
            if not pipe:
                yield "blocked"
                continue
            op = pipe.pop(0)
            if op[0] == "commit":
                op[1].arrive()
            else:
                _, tile, kb, acc, stage = op
                assert stage_content[stage] == (tile, kb), f"MMA of tile {tile} kb {kb} read stage holding {stage_content[stage]}"
                assert acc_drained.get((tile - 2, 0), False), f"MMA into accumulator stage {acc} of tile {tile} before tile {tile - 2} was drained"
                assert acc_mmas.get(tile, 0) == kb
                acc_mmas[tile] = kb + 1
            yield "step"

    agents = {"producer": producer(), "issuer": issuer(), "epi0": epilogue(0), "epi1": epilogue(1), "pipe": tensor_pipe()}
    blocked_rounds = 0
    while agents:
        name = rng.choice(sorted(agents))
        try:
            r = next(agents[name])
        except StopIteration:
            del agents[name]
            blocked_rounds = 0
            continue
        blocked_rounds = blocked_rounds + 1 if r == "blocked" else 0
        if blocked_rounds > 20000:
            raise Deadlock(f"roles still alive: {sorted(agents)} (two sets, tiles {num_tiles}, kb {kb_total}, seed {seed})")
    assert all(acc_drained.get((t, 0)) for t in range(num_tiles))


def simulate_transform(num_tiles, kb_total, seed, stages=6, groups=4, guard=True):
    """The A-operand transform kernels (LoRA dropout in shared memory): `groups` transform groups take ring positions in
    turn (position si belongs to group si % groups), wait for the TMA data with the parity of (si / stages), mask the
    tile in place and arrive on the stage's "transformed" barrier, which is what the UMMA issuer waits for.

    TMA loads are asynchronous and complete in ANY order (round 2: the model of round 1 completed them in issue order
    and therefore could not see the stall it was written to exclude).  `guard=True` is the kernel as fixed in round 2: a
    group first waits for the previous use of the stage to be released (empty barrier), which makes its parity wait on
    the full barrier exact.  `guard=False` is the round-1 protocol; run_all() checks that the model catches it."""
    rng = random.Random(seed)
    full = [Bar(1) for _ in range(stages)]
    xf = [Bar(1) for _ in range(stages)]        # one arrive per warp of the owning group (modelled as one)
    empty = [Bar(1) for _ in range(stages)]
    stage_content, stage_masked = [None] * stages, [False] * stages
    pipe = []
    inflight = []                               # TMA loads issued and not yet landed: (stage, tile, kb)
    log = {"issuer_done": False, "mmas": 0, "producer_done": False}

    def wait(bar, parity, intended):
        while not bar.passes(parity):
            yield "blocked"
        assert bar.completed == intended + 1, f"parity alias: wanted phase {intended}, barrier completed {bar.completed}"

    def producer():
        s, ph, use = 0, 0, [0] * stages
        for tile in range(num_tiles):
            for kb in range(kb_total):
                if use[s] > 0:
                    yield from wait(empty[s], ph ^ 1, use[s] - 1)
                inflight.append((s, tile, kb))
                yield "step"
                use[s] += 1
                s += 1
                if s == stages:
                    s, ph = 0, ph ^ 1
        log["producer_done"] = True

    def tma():                                   # lands the in-flight loads in random order
        while not log["producer_done"] or inflight:
            if not inflight:
                yield "blocked"
                continue
            if rng.random() < 0.7:              # slow memory: loads stay in flight for a while
                yield "step"
                continue
            st, tile, kb = inflight.pop(rng.randrange(len(inflight)))
            stage_content[st], stage_masked[st] = (tile, kb), False
            full[st].arrive()
            yield "step"

    def group(g):
        it = 0
        for tile in range(num_tiles):
            kb = (g - it) & (groups - 1)
            while kb < kb_total:
                si = it + kb
                st, ph = si % stages, (si // stages) & 1
                if guard and si >= stages:
                    yield from wait(empty[st], ph ^ 1, si // stages - 1)
                yield from wait(full[st], ph, si // stages)
                assert stage_content[st] == (tile, kb) and not stage_masked[st]
                yield "step"
                stage_masked[st] = True
                xf[st].arrive()
                kb += groups
            it += kb_total

    def issuer():
        s, ph, use = 0, 0, [0] * stages
        for tile in range(num_tiles):
            for kb in range(kb_total):
                yield from wait(xf[s], ph, use[s])
                pipe.append(("mma", tile, kb, s))
                yield "step"
                pipe.append(("commit", empty[s]))
                use[s] += 1
                s += 1
                if s == stages:
                    s, ph = 0, ph ^ 1
        log["issuer_done"] = True

    def tensor_pipe():
        while not log["issuer_done"] or pipe:
            if not pipe:
                yield "blocked"
                continue
            op = pipe.pop(0)
            if op[0] == "commit":
                op[1].arrive()
            else:
                _, tile, kb, st = op
                assert stage_content[st] == (tile, kb) and stage_masked[st], "MMA read an unmasked or overwritten stage"
                log["mmas"] += 1
            yield "step"

    agents = {"producer": producer(), "tma": tma(), "issuer": issuer(), "pipe": tensor_pipe()}
    agents.update({f"xf{g}": group(g) for g in range(groups)})
    blocked_rounds = 0
    while agents:
        name = rng.choice(sorted(agents))
        try:
            r = next(agents[name])
        except StopIteration:
            del agents[name]
            blocked_rounds = 0
            continue
        blocked_rounds = blocked_rounds + 1 if r == "blocked" else 0
        if blocked_rounds > 40000:
            raise Deadlock(f"roles still alive: {sorted(agents)} (transform, tiles {num_tiles}, kb {kb_total}, seed {seed})")
    assert log["mmas"] == num_tiles * kb_total


def simulate_decode(num_tiles, kb_main, kb_tail, seed, stages=4, pst=4, ng=2, guard=True, fence_reads=True):
    """The decode kernels' main loop: operand ring (A by TMA; B decoded for main k-blocks, by TMA for the LoRA tail
    k-blocks), an independent packed-NF4 ring of depth `pst`, and `ng` decode groups -- group g owns the operand-ring
    positions si with si % ng == g, reads packed position pi = (main k-blocks so far) and releases it, waits for the
    operand slot, writes the decoded tile and arrives on the stage's full barrier; for tail k-blocks the owner only
    arrives.

    Round 2: packed-ring loads land in any order, and a group's READ of a packed slot is asynchronous too (issued, then
    returned some steps later).  `guard` = wait for the previous use of the packed slot to be released before the parity
    wait on its "landed" barrier (the reader of a slot alternates between the groups at tile boundaries when kb_tail is
    odd).  `fence_reads` = release the slot only after the read has returned.  Both False is the round-1 kernel."""
    rng = random.Random(seed)
    kb_total = kb_main + kb_tail
    full = [Bar(2) for _ in range(stages)]      # TMA producer + the owning decode group
    empty = [Bar(1) for _ in range(stages)]
    pk = [Bar(1) for _ in range(pst)]
    pk_empty = [Bar(1) for _ in range(pst)]
    a_content, b_content, p_content = [None] * stages, [None] * stages, [None] * pst
    pipe = []
    p_inflight = []                             # packed loads issued and not yet landed: (slot, tile, kb)
    log = {"issuer_done": False, "mmas": 0, "packed_done": False}

    def wait(bar, parity, intended):
        while not bar.passes(parity):
            yield "blocked"
        assert bar.completed == intended + 1, f"parity alias: wanted phase {intended}, barrier completed {bar.completed}"

    def producer():
        s, ph, use = 0, 0, [0] * stages
        for tile in range(num_tiles):
            for kb in range(kb_total):
                if use[s] > 0:
                    yield from wait(empty[s], ph ^ 1, use[s] - 1)
                a_content[s] = (tile, kb)
                if kb >= kb_main:
                    b_content[s] = (tile, kb)      # tail: B by TMA
                yield "step"
                full[s].arrive()
                use[s] += 1
                s += 1
                if s == stages:
                    s, ph = 0, ph ^ 1

    def packed_producer():
        ps, pph, use = 0, 0, [0] * pst
        for tile in range(num_tiles):
            for kb in range(kb_main):
                if use[ps] > 0:
                    yield from wait(pk_empty[ps], pph ^ 1, use[ps] - 1)
                p_inflight.append((ps, tile, kb))
                yield "step"
                use[ps] += 1
                ps += 1
                if ps == pst:
                    ps, pph = 0, pph ^ 1
        log["packed_done"] = True

    def packed_tma():
        while not log["packed_done"] or p_inflight:
            if not p_inflight:
                yield "blocked"
                continue
            if rng.random() < 0.7:              # slow memory: loads stay in flight for a while
                yield "step"
                continue
            ps, tile, kb = p_inflight.pop(rng.randrange(len(p_inflight)))
            p_content[ps] = (tile, kb)
            pk[ps].arrive()
            yield "step"

    def group(g):
        it = pit = 0
        for tile in range(num_tiles):
            kb = (g - it) & (ng - 1)
            while kb < kb_main:
                pi = pit + kb
                ps, pph = pi % pst, (pi // pst) & 1
                if guard and pi >= pst:
                    yield from wait(pk_empty[ps], pph ^ 1, pi // pst - 1)
                yield from wait(pk[ps], pph, pi // pst)
                if fence_reads:                 # the load returns (its value is consumed) before the slot is released
                    yield "step"
                    assert p_content[ps] == (tile, kb), "decode read a packed slot holding another k-block"
                    pk_empty[ps].arrive()
                else:                           # round 1: released as soon as the load is issued; it returns later
                    pk_empty[ps].arrive()
                    for _ in range(rng.randrange(0, 12)):
                        yield "step"
                    assert p_content[ps] == (tile, kb), "decode read a packed slot holding another k-block"
                si = it + kb
                st, ph = si % stages, (si // stages) & 1
                if si >= stages:
                    yield from wait(empty[st], ph ^ 1, si // stages - 1)
                b_content[st] = (tile, kb)
                yield "step"
                full[st].arrive()
                kb += ng
            for kb in range(kb_main, kb_total):
                si = it + kb
                if (si & (ng - 1)) != g:
                    continue
                st, ph = si % stages, (si // stages) & 1
                if si >= stages:
                    yield from wait(empty[st], ph ^ 1, si // stages - 1)
                full[st].arrive()
            it += kb_total
            pit += kb_main

    def issuer():
        s, ph, use = 0, 0, [0] * stages
        for tile in range(num_tiles):
            for kb in range(kb_total):
                yield from wait(full[s], ph, use[s])
                pipe.append(("mma", tile, kb, s))
                yield "step"
                pipe.append(("commit", empty[s]))
                use[s] += 1
                s += 1
                if s == stages:
                    s, ph = 0, ph ^ 1
        log["issuer_done"] = True

    def tensor_pipe():
        while not log["issuer_done"] or pipe:
            if not pipe:
                yield "blocked"
                continue
            op = pipe.pop(0)
            if op[0] == "commit":
                op[1].arrive()
            else:
                _, tile, kb, st = op
                assert a_content[st] == (tile, kb) and b_content[st] == (tile, kb), "MMA read a stage holding other data"
                log["mmas"] += 1
            yield "step"

    agents = {"producer": producer(), "packed": packed_producer(), "packed_tma": packed_tma(), "issuer": issuer(),
              "pipe": tensor_pipe()}
    agents.update({f"dec{g}": group(g) for g in range(ng)})
    blocked_rounds = 0
    while agents:
        name = rng.choice(sorted(agents))
        try:
            r = next(agents[name])
        except StopIteration:
            del agents[name]
            blocked_rounds = 0
            continue
        blocked_rounds = blocked_rounds + 1 if r == "blocked" else 0
        if blocked_rounds > 40000:
            raise Deadlock(f"roles still alive: {sorted(agents)} (decode, tiles {num_tiles}, kb {kb_main}+{kb_tail}, seed {seed})")
    assert log["mmas"] == num_tiles * kb_total


def run_all(seeds=12):
    n = 0
    for kb_total in (1, 2, 3, 4, 5, 6, 7, 9, 17):
        for tiles in (1, 2, 3, 5):
            for seed in range(seeds):
                simulate(tiles, kb_total, seed)
                n += 1
    for kb_total in (1, 2, 3):
        for tiles in (1, 2, 3, 4, 7):
            for seed in range(seeds):
                simulate_two_sets(tiles, kb_total, seed)
    for kb_total in (1, 2, 3, 5, 6, 7, 13, 64):
        for tiles in (1, 2, 3):
            for seed in range(max(1, seeds // 3)):
                simulate_transform(tiles, kb_total, seed)
    for kb_main in (1, 2, 4, 5, 8, 16, 64):
        for kb_tail in (0, 1, 2):
            for tiles in (1, 2, 3, 5):
                for seed in range(max(1, seeds // 3)):
                    simulate_decode(tiles, kb_main, kb_tail, seed)
    # mutation checks (round 2): the round-1 protocols must FAIL in the model now that loads land out of order
    def fails(fn, *a, **k):
        bad = 0
        for seed in range(300):
            try:
                fn(*a, seed, **k)
            except (AssertionError, Deadlock):
                bad += 1
        return bad
    assert fails(simulate_transform, 2, 13, guard=False) > 0, "model does not see the transform-group phase alias"
    assert fails(simulate_decode, 3, 5, 1, guard=False) > 0, "model does not see the packed-ring alias at tile boundaries"
    assert fails(simulate_decode, 2, 8, 0, fence_reads=False) > 0, "model does not see the early release of a packed slot"
    return n


if __name__ == "__main__":
    print("configurations simulated without deadlock or hazard:", run_all())
