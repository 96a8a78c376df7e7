"""A small pytree: find the leaves of nested containers and put them back.

Role parity: fed/tree_util.py:180-231 (``tree_flatten`` / ``tree_unflatten``)
— the DAG layer uses it to discover ``FedObject``s buried inside list / tuple /
dict / namedtuple / OrderedDict arguments.  This is an independent
implementation: a node registry maps a container type to a pair of functions
``(children, aux) = to_children(obj)`` and ``obj = from_children(children,
aux)``; a ``TreeSpec`` is an immutable tree of ``(type, aux, child_specs)`` with
a cached leaf count, and unflattening consumes a single iterator (no slicing).
"""
from __future__ import annotations

from collections import OrderedDict, defaultdict, deque
from typing import Any, Callable, Dict, Iterable, List, NamedTuple, Optional, Tuple, Type


class _NodeDef(NamedTuple):
    to_children: Callable[[Any], Tuple[List[Any], Any]]
    from_children: Callable[[Iterable[Any], Any], Any]


_REGISTRY: Dict[Type, _NodeDef] = {}


def register_pytree_node(typ: Type, to_children, from_children) -> None:
    """Teach the pytree about a new container type."""
    _REGISTRY[typ] = _NodeDef(to_children, from_children)


def _is_namedtuple(obj) -> bool:
    t = type(obj)
    return isinstance(obj, tuple) and hasattr(t, "_fields") and hasattr(t, "_make")


register_pytree_node(list, lambda x: (list(x), None), lambda c, _: list(c))
register_pytree_node(tuple, lambda x: (list(x), None), lambda c, _: tuple(c))
register_pytree_node(dict, lambda d: (list(d.values()), list(d.keys())), lambda c, keys: dict(zip(keys, c)))
register_pytree_node(
    OrderedDict, lambda d: (list(d.values()), list(d.keys())), lambda c, keys: OrderedDict(zip(keys, c))
)
register_pytree_node(
    defaultdict,
    lambda d: (list(d.values()), (d.default_factory, list(d.keys()))),
    lambda c, aux: defaultdict(aux[0], zip(aux[1], c)),
)
register_pytree_node(deque, lambda d: (list(d), d.maxlen), lambda c, maxlen: deque(c, maxlen=maxlen))


class TreeSpec:
    """Structure of a flattened object. ``node_type is None`` marks a leaf."""

    __slots__ = ("node_type", "aux", "children", "num_leaves")

    def __init__(self, node_type: Optional[Type], aux: Any, children: Tuple["TreeSpec", ...]):
        self.node_type = node_type
        self.aux = aux
        self.children = children
        self.num_leaves = 1 if node_type is None else sum(c.num_leaves for c in children)

    def is_leaf(self) -> bool:
        return self.node_type is None

    def __eq__(self, other):
        return (
            isinstance(other, TreeSpec)
            and self.node_type is other.node_type
            and self.aux == other.aux
            and self.children == other.children
        )

    def __hash__(self):
        return hash((self.node_type, self.num_leaves, len(self.children)))

    def __repr__(self):
        if self.is_leaf():
            return "*"
        name = getattr(self.node_type, "__name__", str(self.node_type))
        return f"{name}({', '.join(map(repr, self.children))})"


LEAF_SPEC = TreeSpec(None, None, ())


_SCALAR_LEAVES = frozenset({type(None), bool, int, float, str, bytes, complex})


def has_containers(args, kwargs) -> bool:
    """Cheap pre-check used on the per-call hot path: is there anything to flatten?"""
    for a in args:
        t = type(a)
        if t in _SCALAR_LEAVES:
            continue
        if t in _REGISTRY or isinstance(a, tuple):
            return True
    if kwargs:
        for a in kwargs.values():
            t = type(a)
            if t in _SCALAR_LEAVES:
                continue
            if t in _REGISTRY or isinstance(a, tuple):
                return True
    return False


def _node_def(obj) -> Optional[_NodeDef]:
    if type(obj) in _SCALAR_LEAVES:
        return None
    nd = _REGISTRY.get(type(obj))
    if nd is not None:
        return nd
    if _is_namedtuple(obj):
        return _NodeDef(lambda x: (list(x), type(x)), lambda c, t: t(*c))
    return None


def tree_flatten(tree: Any) -> Tuple[List[Any], TreeSpec]:
    """Return ``(leaves, spec)``; anything that is not a registered container
    (including ``None``) is a leaf."""
    leaves: List[Any] = []

    def walk(obj) -> TreeSpec:
        nd = _node_def(obj)
        if nd is None:
            leaves.append(obj)
            return LEAF_SPEC
        children, aux = nd.to_children(obj)
        specs = tuple(walk(c) for c in children)
        node_type = type(obj)
        if _is_namedtuple(obj) and node_type not in _REGISTRY:
            # remember how to rebuild this particular namedtuple class
            aux = ("__namedtuple__", node_type)
            node_type = tuple
        return TreeSpec(node_type, aux, specs)

    return leaves, walk(tree)


def tree_unflatten(leaves: Iterable[Any], spec: TreeSpec) -> Any:
    """Inverse of :func:`tree_flatten` (leaves may have been replaced)."""
    leaves = list(leaves)
    if len(leaves) != spec.num_leaves:
        raise ValueError(f"tree_unflatten: spec expects {spec.num_leaves} leaves, got {len(leaves)}")
    it = iter(leaves)

    def build(s: TreeSpec):
        if s.is_leaf():
            return next(it)
        kids = [build(c) for c in s.children]
        if isinstance(s.aux, tuple) and len(s.aux) == 2 and s.aux[0] == "__namedtuple__":
            return s.aux[1](*kids)
        return _REGISTRY[s.node_type].from_children(kids, s.aux)

    return build(spec)


def tree_map(fn: Callable[[Any], Any], tree: Any) -> Any:
    leaves, spec = tree_flatten(tree)
    return tree_unflatten([fn(x) for x in leaves], spec)


def tree_leaves(tree: Any) -> List[Any]:
    return tree_flatten(tree)[0]
