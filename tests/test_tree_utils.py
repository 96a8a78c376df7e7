"""pytree behaviour (mirrors fed/tests/without_ray_tests/test_tree_utils.py)."""
from collections import OrderedDict, namedtuple

import pytest

import rayfed_b200.tree_util as tree_utils


def test_flatten_none():
    leaves, spec = tree_utils.tree_flatten(None)
    assert isinstance(leaves, list) and len(leaves) == 1
    assert tree_utils.tree_unflatten(leaves, spec) is None


@pytest.mark.parametrize("target", [1, 0.5, "hello", b"world"])
def test_flatten_single_primitive(target):
    leaves, spec = tree_utils.tree_flatten(target)
    assert leaves == [target]
    assert tree_utils.tree_unflatten(leaves, spec) == target


@pytest.mark.parametrize("target", [[1, 2, 3], (1, 2, 3), {"a": 1, "b": 2, "c": 3}])
def test_flatten_simple_containers(target):
    leaves, spec = tree_utils.tree_flatten(target)
    assert len(leaves) == len(target)
    assert tree_utils.tree_unflatten(leaves, spec) == target


def test_flatten_nested_with_set_leaf():
    o = [1, 2, (3, 4), [5, {"b", 6}, 7], 8]  # a set is a leaf
    leaves, spec = tree_utils.tree_flatten(o)
    assert len(leaves) == 8
    assert tree_utils.tree_unflatten(leaves, spec) == o


def test_flatten_and_replace_element():
    o = [1, 2, (3, 4), [5, {"b": 6}, 7], 8]
    leaves, spec = tree_utils.tree_flatten(o)
    leaves[0] = "hello"
    leaves[5] = b"world"
    res = tree_utils.tree_unflatten(leaves, spec)
    assert res == ["hello", 2, (3, 4), [5, {"b": b"world"}, 7], 8]


def test_namedtuple_and_ordereddict_roundtrip():
    Point = namedtuple("Point", "x y")
    o = {"p": Point(1, [2, 3]), "od": OrderedDict(a=1, b=(2, 3))}
    leaves, spec = tree_utils.tree_flatten(o)
    assert leaves == [1, 2, 3, 1, 2, 3]
    back = tree_utils.tree_unflatten([x * 10 for x in leaves], spec)
    assert isinstance(back["p"], Point) and back["p"] == Point(10, [20, 30])
    assert isinstance(back["od"], OrderedDict) and list(back["od"].items()) == [("a", 10), ("b", (20, 30))]


def test_unflatten_rejects_wrong_leaf_count():
    leaves, spec = tree_utils.tree_flatten([1, 2])
    with pytest.raises(ValueError):
        tree_utils.tree_unflatten(leaves + [3], spec)


def test_tree_map():
    assert tree_utils.tree_map(lambda v: v + 1, {"a": [1, 2], "b": (3,)}) == {"a": [2, 3], "b": (4,)}
