"""Executable model of the cross-GPU epoch protocol of csrc/fed_comm.cuh (docs/PROTOCOL.md).

Stores to other GPUs / the host become visible late and out of order: every store goes into a pool and is
applied at a random later time; only a *release* store is held back until the writer's earlier stores
have landed (fence + st.release.sys), and a flag-in-data ("LL") word is its own unit.  Kernels of one
node run back to back (stream order), their CTAs interleave freely with everything else.  The model
checks, for random interleavings and several epochs, that

* every CTA of every node computes with exactly the theta of its epoch (never a mix of two epochs),
* the host receives, per epoch, exactly the rank-ordered sum of that epoch's node partials,
* nothing dead-locks — and that the checks have teeth: dropping the fence is detected.

`group > 0` models the two-level reduction of round 2 (`epilogue_t`: tickets per group of `kReduceGroup` CTAs,
the last CTA of a group writes the group partial, the last GROUP sums the group partials): every group partial
must be complete when it is read, exactly one CTA per node runs the final stage, and every ticket is back at
zero when the node's next kernel starts.
"""
import random

import pytest

N_THETA = 3


class World:
    def __init__(self, nodes, ctas, epochs, rng, ll=False, fence=True, group=0, reset_group_ticket=True,
                 expiry=0.0, unanimous=True, spec=False, all_or_nothing=True):
        self.n, self.g, self.epochs, self.rng, self.ll, self.fence = nodes, ctas, epochs, rng, ll, fence
        self.group, self.reset_group_ticket = group, reset_group_ticket
        n_groups = (ctas + group - 1) // group if group else 0
        self.group_ticket = [[0] * n_groups for _ in range(nodes)]
        self.group_partials = [[None] * n_groups for _ in range(nodes)]
        self.finalizers = {}
        # bounded waits: a peer CTA's wait for theta may run out (probability `expiry` per poll); the launch then
        # counts as idle and the serve loop re-arms the node with a kernel for the SAME epoch
        self.expiry, self.unanimous = expiry, unanimous
        self.any_expired = [False] * nodes
        self.idle = [False] * nodes
        self.relaunches = 0
        # speculative root launches: the root's kernels are in the stream BEFORE the client has written theta (as
        # tagged words in host memory); CTA 0 polls them, gives up after a while and releases the other CTAs of
        # its GPU through the abort word; the launch is idle and the next one takes the epoch over
        self.spec, self.all_or_nothing = spec, all_or_nothing
        self.host_words = [(0.0, 0)] * N_THETA
        self.abort = 0
        zero_word = (0.0, 0)
        self.mail = [{"theta": [zero_word if ll else 0.0] * N_THETA, "flag": 0} for _ in range(nodes)]
        self.slots = [zero_word if ll else 0.0 for _ in range(nodes)]
        self.slot_flags = [0] * nodes
        self.host_theta = [0.0] * N_THETA
        self.host_result = zero_word if ll else 0.0
        self.host_flag = 0
        self.partials = [[0.0] * ctas for _ in range(nodes)]
        self.ticket = [0] * nodes
        self.pool = []          # pending remote stores: [writer, seq, is_release, apply]
        self.seq = 0
        self.results = {}

    # -- memory ---------------------------------------------------------------------------------------
    def store(self, writer, apply, release=False):
        self.seq += 1
        self.pool.append([writer, self.seq, release, apply])

    def deliverable(self):
        out = []
        for item in self.pool:
            writer, seq, release, _ = item
            if release and self.fence and any(w == writer and s < seq for w, s, _, _ in self.pool):
                continue            # a release store waits for the writer's earlier stores
            out.append(item)
        return out

    @staticmethod
    def theta_of(epoch):
        return [epoch * 10.0 + i for i in range(N_THETA)]

    @staticmethod
    def partial_of(theta, node, cta):
        return sum(theta) * (node + 1) + cta

    # -- device code ------------------------------------------------------------------------------------
    def cta(self, node, c, epoch):
        root = node == 0
        writer = ("cta", node, c, epoch)
        aborted = False
        if root and c == 0 and self.spec:                     # speculative: wait for the client's tagged words
            theta = []
            for i in range(N_THETA):
                while not aborted and self.host_words[i][1] != epoch:
                    aborted = self.rng.random() < self.expiry
                    yield
                theta.append(self.host_words[i][0])
            if aborted and self.all_or_nothing:
                self.abort = epoch                            # nothing is broadcast
            elif aborted:                                     # broken variant: sends what it has
                aborted = False
        if root and c == 0 and not aborted:                   # broadcast
            if not self.spec:
                theta = list(self.host_theta)
            for peer in range(self.n):
                for i, v in enumerate(theta):
                    word = (v, epoch) if self.ll else v
                    self.store(writer, lambda p=peer, i=i, w=word: self.mail[p]["theta"].__setitem__(i, w))
                    yield
                if not self.ll:
                    self.store(writer, lambda p=peer: self.mail[p].__setitem__("flag", epoch), release=True)
        got = []
        expired = False
        gives_up = lambda: ((not root) and self.expiry > 0 and self.rng.random() < self.expiry) or \
            (root and self.spec and self.abort == epoch)
        if self.ll:                                           # acquire: poll the tagged words
            for i in range(N_THETA):
                while not expired and self.mail[node]["theta"][i][1] != epoch:
                    expired = gives_up()
                    yield
                got.append(self.mail[node]["theta"][i][0])
        else:                                                 # acquire: flag, then the data
            while not expired and self.mail[node]["flag"] < epoch:
                expired = gives_up()
                yield
            for i in range(N_THETA):
                got.append(self.mail[node]["theta"][i])
                yield
        if expired:
            self.any_expired[node] = True                     # atomicOr next to the tickets; the partial stays stale
        else:
            assert got == self.theta_of(epoch), f"node {node} CTA {c} computed epoch {epoch} with theta {got}"
            self.partials[node][c] = self.partial_of(got, node, c)
        yield
        if self.group:                                        # two-level, fixed shape
            grp, first = c // self.group, (c // self.group) * self.group
            size = min(self.group, self.g - first)
            n_groups = len(self.group_ticket[node])
            self.group_ticket[node][grp] += 1                 # atomic
            if self.group_ticket[node][grp] != size:
                return
            if self.reset_group_ticket:
                self.group_ticket[node][grp] = 0
            yield
            self.group_partials[node][grp] = (epoch, sum(self.partials[node][first:first + size]))
            yield
            self.ticket[node] += 1                            # atomic
            if self.ticket[node] != n_groups:
                return
            yield
            assert all(gp is not None and gp[0] == epoch for gp in self.group_partials[node]), \
                "the final stage read a group partial of another epoch"
            node_sum = sum(gp[1] for gp in self.group_partials[node])
            self.ticket[node] = 0
        else:
            self.ticket[node] += 1                            # atomic
            if self.ticket[node] != self.g:
                return
            self.ticket[node] = 0
            node_sum = sum(self.partials[node])               # fixed order
        # the final stage publishes only if EVERY CTA had theta (fed_comm.cuh: any_timed_out); the broken variant
        # looks at its own wait only
        if (self.any_expired[node] if self.unanimous else expired):
            self.any_expired[node] = False
            self.idle[node] = True
            if root:
                self.abort = 0                                # the next launch of this epoch starts clean
            return
        self.any_expired[node] = False
        self.finalizers[(node, epoch)] = self.finalizers.get((node, epoch), 0) + 1
        word = (node_sum, epoch) if self.ll else node_sum
        self.store(writer, lambda: self.slots.__setitem__(node, word))
        if not self.ll:
            self.store(writer, lambda: self.slot_flags.__setitem__(node, epoch), release=True)
        if not root:
            return
        total = 0.0
        for rank in range(self.n):                            # rank order
            if self.ll:
                while self.slots[rank][1] != epoch:
                    yield
                total += self.slots[rank][0]
            else:
                while self.slot_flags[rank] < epoch:
                    yield
                total += self.slots[rank]
            yield
        if self.ll:
            self.store(writer, lambda: setattr(self, "host_result", (total, epoch)))
        else:
            self.store(writer, lambda: setattr(self, "host_result", total))
            self.store(writer, lambda: setattr(self, "host_flag", epoch), release=True)

    def node_stream(self, node):
        """Kernels of one node in stream order; the root's are launched by the host, peers' are pre-enqueued."""
        for epoch in range(1, self.epochs + 1):
            if node == 0 and not self.spec:
                while self.launched < epoch:
                    yield
            while True:
                assert self.ticket[node] == 0 and not any(self.group_ticket[node]), "a ticket was left armed for the next launch"
                self.idle[node] = False
                ctas = [self.cta(node, c, epoch) for c in range(self.g)]
                while ctas:
                    c = self.rng.choice(ctas)
                    try:
                        next(c)
                    except StopIteration:
                        ctas.remove(c)
                    yield
                if not self.idle[node]:
                    break
                self.relaunches += 1                          # idle tick: same epoch again

    def host(self):
        for epoch in range(1, self.epochs + 1):
            if self.spec:                                     # the kernels are already waiting; the client takes its time
                for _ in range(self.rng.randint(0, 30)):
                    yield
                for i, v in enumerate(self.theta_of(epoch)):
                    self.host_words[i] = (v, epoch)
                    yield
            self.host_theta = self.theta_of(epoch)            # visible to the kernel launched afterwards
            self.launched = epoch
            if self.ll:
                while self.host_result[1] != epoch:
                    yield
                self.results[epoch] = self.host_result[0]
            else:
                while self.host_flag < epoch:
                    yield
                self.results[epoch] = self.host_result
            yield

    def run(self):
        self.launched = 0
        actors = [self.host()] + [self.node_stream(n) for n in range(self.n)]
        steps = 0
        while actors:
            steps += 1
            assert steps < 200_000, "dead-lock / live-lock"
            ready = self.deliverable()
            if ready and (self.rng.random() < 0.5):
                item = self.rng.choice(ready)
                self.pool.remove(item)
                item[3]()
                continue
            a = self.rng.choice(actors)
            try:
                next(a)
            except StopIteration:
                actors.remove(a)
        for item in list(self.pool):
            item[3]()
        return self

    def expected(self, epoch):
        theta = self.theta_of(epoch)
        return sum(self.partial_of(theta, n, c) for n in range(self.n) for c in range(self.g))


@pytest.mark.parametrize("ll", [False, True], ids=["fence+flag", "flag-in-data"])
def test_epochs_never_mix_and_results_are_exact(ll):
    rng = random.Random(99 + ll)
    for trial in range(60):
        nodes, ctas, epochs = rng.randint(1, 4), rng.randint(1, 3), rng.randint(1, 5)
        w = World(nodes, ctas, epochs, random.Random(rng.random()), ll=ll).run()
        assert w.results == {e: w.expected(e) for e in range(1, epochs + 1)}, (nodes, ctas, epochs)


def test_the_model_notices_a_missing_fence():
    """Without fence + release ordering the flag can overtake theta: some schedule reads a stale word."""
    rng = random.Random(5)
    with pytest.raises(AssertionError):
        for _ in range(300):
            World(3, 2, 4, random.Random(rng.random()), ll=False, fence=False).run()


@pytest.mark.parametrize("ll", [False, True], ids=["fence+flag", "flag-in-data"])
def test_two_level_reduce_is_exact_and_runs_one_final_stage_per_node(ll):
    rng = random.Random(7 + ll)
    for trial in range(60):
        nodes, ctas, epochs, group = rng.randint(1, 3), rng.randint(1, 7), rng.randint(1, 4), rng.randint(1, 3)
        w = World(nodes, ctas, epochs, random.Random(rng.random()), ll=ll, group=group).run()
        assert w.results == {e: w.expected(e) for e in range(1, epochs + 1)}, (nodes, ctas, epochs, group)
        assert w.finalizers == {(n, e): 1 for n in range(nodes) for e in range(1, epochs + 1)}


def test_the_model_notices_a_group_ticket_that_is_not_reset():
    with pytest.raises(AssertionError):
        World(2, 4, 3, random.Random(1), group=2, reset_group_ticket=False).run()


@pytest.mark.parametrize("ll", [False, True], ids=["fence+flag", "flag-in-data"])
@pytest.mark.parametrize("group", [0, 2])
def test_expired_waits_never_publish_a_stale_partial(ll, group):
    """Per-CTA bounded waits: theta arriving at the deadline leaves some CTAs with theta and others without.  The
    launch must count as idle unless EVERY CTA computed; the re-armed kernel then delivers the exact result."""
    rng = random.Random(31 + ll + group)
    relaunches = 0
    for trial in range(60):
        nodes, ctas, epochs = rng.randint(2, 4), rng.randint(2, 5), rng.randint(2, 4)
        w = World(nodes, ctas, epochs, random.Random(rng.random()), ll=ll, group=group, expiry=0.05).run()
        assert w.results == {e: w.expected(e) for e in range(1, epochs + 1)}, (nodes, ctas, epochs)
        assert all(v == 1 for v in w.finalizers.values())
        relaunches += w.relaunches
    assert relaunches > 20          # the scenario really occurs


def test_the_model_notices_a_final_stage_that_only_checks_its_own_wait():
    rng = random.Random(3)
    with pytest.raises(AssertionError):
        for _ in range(300):
            w = World(3, 4, 3, random.Random(rng.random()), ll=True, expiry=0.05, unanimous=False).run()
            assert w.results == {e: w.expected(e) for e in range(1, 4)}


@pytest.mark.parametrize("group", [0, 2])
def test_speculative_root_launches_pick_theta_up_or_give_up_cleanly(group):
    """The root's kernel is in the stream before theta exists.  Whatever the timing — theta in time, too late, or
    half written at the deadline — every epoch is computed exactly once with its own theta, by a launch in which
    every CTA of every node had it."""
    rng = random.Random(77 + group)
    relaunches = 0
    for trial in range(80):
        nodes, ctas, epochs = rng.randint(1, 3), rng.randint(1, 4), rng.randint(2, 4)
        w = World(nodes, ctas, epochs, random.Random(rng.random()), ll=True, group=group, spec=True, expiry=0.04).run()
        assert w.results == {e: w.expected(e) for e in range(1, epochs + 1)}, (nodes, ctas, epochs)
        assert all(v == 1 for v in w.finalizers.values())
        relaunches += w.relaunches
    assert relaunches > 20


def test_the_model_notices_a_root_that_broadcasts_a_half_read_theta():
    """CTA 0 gave up with some of the client's words still missing: broadcasting what it has would let the other
    CTAs (and GPUs) compute with a mixture of two epochs."""
    rng = random.Random(9)
    with pytest.raises(AssertionError):
        for _ in range(200):
            World(2, 3, 3, random.Random(rng.random()), ll=True, spec=True, expiry=0.1, all_or_nothing=False).run()
