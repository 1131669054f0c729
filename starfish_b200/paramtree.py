"""Nested parameter container with ``:``-joined flat keys.

The reference keeps ``SpectrumModel.params`` in a ``flatdict.FlatterDict`` (Starfish/models/
spectrum_model.py:166).  flatdict is not part of this image, so the drop-in model carries its own small
container with the behaviours that class relies on: nested and flat access (``p["global_cov:log_amp"]``,
``p["local_cov"]["1"]["mu"]``), lists stored as index-keyed children that ``as_dict()`` turns back into
lists, flat leaf keys in insertion order, group deletion and dict equality.
"""
from __future__ import annotations

SEP = ":"


class ParamTree:
    __slots__ = ("_kids", "_is_seq")

    def __init__(self, source=None):
        self._kids = {}
        self._is_seq = False
        if source is None:
            return
        if isinstance(source, ParamTree):
            source = source.as_dict()
        if isinstance(source, (list, tuple)):
            self._is_seq = True
            source = {str(i): v for i, v in enumerate(source)}
        for key, val in dict(source).items():
            self[key] = val

    # -- access ------------------------------------------------------------------------------------
    @staticmethod
    def _path(key):
        return str(key).split(SEP)

    def _descend(self, path, create=False):
        node = self
        for part in path:
            nxt = node._kids.get(part) if isinstance(node, ParamTree) else None
            if not isinstance(nxt, ParamTree):
                if not create or not isinstance(node, ParamTree):
                    raise KeyError(SEP.join(path))
                nxt = ParamTree()
                node._kids[part] = nxt
            node = nxt
        return node

    def __getitem__(self, key):
        *parents, leaf = self._path(key)
        node = self._descend(parents)
        if leaf not in node._kids:
            raise KeyError(key)
        return node._kids[leaf]

    def __setitem__(self, key, value):
        *parents, leaf = self._path(key)
        node = self._descend(parents, create=True)
        if isinstance(value, (dict, list, tuple)):
            value = ParamTree(value)
        node._kids[leaf] = value

    def __delitem__(self, key):
        *parents, leaf = self._path(key)
        node = self._descend(parents)
        if leaf not in node._kids:
            raise KeyError(key)
        del node._kids[leaf]

    def __contains__(self, key):
        try:
            self[key]
        except KeyError:
            return False
        return True

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    # -- flat views ----------------------------------------------------------------------------------
    def _walk(self, prefix=""):
        for name, val in self._kids.items():
            full = f"{prefix}{name}"
            if isinstance(val, ParamTree):
                yield from val._walk(full + SEP)
            else:
                yield full, val

    def keys(self):
        return [k for k, _ in self._walk()]

    def values(self):
        return [v for _, v in self._walk()]

    def items(self):
        return list(self._walk())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return sum(1 for _ in self._walk())

    def update(self, other):
        for k, v in (other.items() if hasattr(other, "items") else other):
            self[k] = v

    # -- conversion / comparison ----------------------------------------------------------------------
    def as_dict(self):
        plain = {k: (v.as_dict() if isinstance(v, ParamTree) else v) for k, v in self._kids.items()}
        if self._is_seq:
            return [plain[k] for k in sorted(plain, key=int)]
        return plain

    def __eq__(self, other):
        if isinstance(other, ParamTree):
            return self.as_dict() == other.as_dict()
        if isinstance(other, (dict, list, tuple)):
            return self.as_dict() == ParamTree(other).as_dict()
        return NotImplemented

    def __repr__(self):
        return f"ParamTree({self.as_dict()!r})"
