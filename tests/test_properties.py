"""Property-based tests (the reference has none): pytree round trips, mailbox
rendezvous under arbitrary interleavings, sequence-id determinism, wire-frame round trips."""
import collections
import pickle

import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

from rayfed_b200.tree_util import tree_flatten, tree_map, tree_unflatten  # noqa: E402

pytestmark = pytest.mark.filterwarnings("ignore::hypothesis.errors.HypothesisWarning")

Point = collections.namedtuple("Point", ["x", "y"])

leaves = st.one_of(st.integers(), st.text(max_size=5), st.none(), st.floats(allow_nan=False), st.binary(max_size=4))


def containers(children):
    return st.one_of(
        st.lists(children, max_size=4),
        st.lists(children, max_size=4).map(tuple),
        st.dictionaries(st.text(max_size=3), children, max_size=4),
        st.tuples(children, children).map(lambda t: Point(*t)),
        st.lists(st.tuples(st.text(max_size=3), children), max_size=3, unique_by=lambda kv: kv[0]).map(
            collections.OrderedDict),
    )


trees = st.recursive(leaves, containers, max_leaves=25)


@settings(max_examples=150, deadline=None)
@given(trees)
def test_flatten_unflatten_is_identity(tree):
    flat, spec = tree_flatten(tree)
    rebuilt = tree_unflatten(flat, spec)
    assert rebuilt == tree and type(rebuilt) is type(tree)


@settings(max_examples=100, deadline=None)
@given(trees)
def test_leaf_replacement_keeps_structure(tree):
    """What resolve_dependencies does: swap leaves, keep the skeleton."""
    flat, spec = tree_flatten(tree)
    marked = tree_unflatten([("leaf", i) for i in range(len(flat))], spec)
    flat2, spec2 = tree_flatten(marked)
    # tuples are containers, so each ("leaf", i) marker flattens into two leaves
    assert flat2 == [v for i in range(len(flat)) for v in ("leaf", i)]
    assert tree_map(lambda v: v, tree) == tree
    assert spec2 is not None


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 30), st.integers(0, 30)), min_size=1, max_size=40, unique=True),
       st.randoms(use_true_random=False))
def test_mailbox_rendezvous_any_interleaving(keys, rnd):
    """Each key is claimed once and delivered once, in an arbitrary global order: every claim
    resolves to its own payload and the table ends empty."""
    from rayfed_b200.core import native

    mb = native().Mailbox()
    ops = [("claim", k) for k in keys] + [("deliver", k) for k in keys]
    rnd.shuffle(ops)
    refs = {}
    for op, (u, d) in ops:
        if op == "claim":
            refs[(u, d)] = mb.claim(str(u), str(d))
        else:
            mb.deliver_raw(str(u), str(d), pickle.dumps((u, d)), False)
    for (u, d), r in refs.items():
        assert r.done()
        got = r.result(1.0)
        got = pickle.loads(got) if isinstance(got, (bytes, bytearray)) else got
        assert tuple(got) == (u, d)
    assert mb.pending() == 0
    s = mb.stats()
    assert s["delivered"] == len(keys) and s["claimed"] == len(keys)


@settings(max_examples=50, deadline=None)
@given(st.lists(st.sampled_from(["task", "actor", "get"]), max_size=60))
def test_seq_ids_depend_only_on_the_call_sequence(calls):
    """Two controllers issuing the same calls mint the same ids: +1 task/method, +2 actor
    creation, +1 fed.get (fed/_private/global_context.py:45-47; fed_actor/api id arithmetic)."""
    from rayfed_b200.core import native

    def run():
        c = native().SeqCounter()
        ids = []
        for kind in calls:
            if kind == "actor":
                ids.append((c.next(), c.next()))
            else:
                ids.append(c.next())
        return ids, c.value()

    a, b = run(), run()
    assert a == b
    assert a[1] == sum(2 if k == "actor" else 1 for k in calls)


@settings(max_examples=80, deadline=None)
@given(st.dictionaries(st.text(max_size=4), st.one_of(st.integers(), st.text(max_size=3)), max_size=4))
def test_dict2tuple_round_trips_and_is_idempotent(d):
    from rayfed_b200.utils import dict2tuple

    t = dict2tuple(d)
    assert isinstance(t, tuple) and dict(t) == d and [k for k, _ in t] == list(d)
    assert dict2tuple(t) is t and dict2tuple(None) is None
    hash(t)  # usable as gRPC channel options / metadata
