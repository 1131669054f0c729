"""Import shim (oracle scaffolding only): a minimal functional stand-in for ``flatdict`` 4.x.

The reference's ``SpectrumModel`` stores its parameters in ``flatdict.FlatterDict`` (delimiter ``:``);
flatdict is not installed in this image and cannot be fetched.  This class implements just the
behaviours SURVEY.md Appendix A lists, so the *unmodified* reference can be imported and run as the
parity oracle inside this container.  It is never imported by the product package.
"""


class FlatterDict:
    def __init__(self, value=None, delimiter=":"):
        self._delimiter = delimiter
        self._values = {}
        self.original_type = dict
        if value is None:
            return
        if isinstance(value, FlatterDict):
            self.original_type = value.original_type
            value = value.as_dict()
        if isinstance(value, (list, tuple)):
            self.original_type = type(value)
            value = {str(i): v for i, v in enumerate(value)}
        for k, v in value.items():
            self[k] = v

    # -- helpers -----------------------------------------------------------------
    def _wrap(self, v):
        if isinstance(v, (dict, list, tuple)) and not isinstance(v, FlatterDict):
            return FlatterDict(v, self._delimiter)
        return v

    def _split(self, key):
        key = str(key)
        if self._delimiter in key:
            return key.split(self._delimiter, 1)
        return key, None

    # -- mapping protocol --------------------------------------------------------
    def __getitem__(self, key):
        head, rest = self._split(key)
        if head not in self._values:
            raise KeyError(key)
        node = self._values[head]
        if rest is None:
            return node
        if not isinstance(node, FlatterDict):
            raise KeyError(key)
        return node[rest]

    def __setitem__(self, key, value):
        head, rest = self._split(key)
        if rest is None:
            self._values[head] = self._wrap(value)
            return
        node = self._values.get(head)
        if not isinstance(node, FlatterDict):
            node = FlatterDict(delimiter=self._delimiter)
            self._values[head] = node
        node[rest] = value

    def __delitem__(self, key):
        head, rest = self._split(key)
        if head not in self._values:
            raise KeyError(key)
        if rest is None:
            del self._values[head]
        else:
            del self._values[head][rest]

    def __contains__(self, key):
        try:
            self[key]
            return True
        except (KeyError, TypeError):
            return False

    def keys(self):
        out = []
        for k, v in self._values.items():
            if isinstance(v, FlatterDict):
                out.extend(f"{k}{self._delimiter}{sub}" for sub in v.keys())
            else:
                out.append(k)
        return out

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def update(self, other=None, **kw):
        for k, v in dict(other or {}, **kw).items():
            self[k] = v

    def as_dict(self):
        out = {}
        for k, v in self._values.items():
            out[k] = v.as_dict() if isinstance(v, FlatterDict) else v
        if self.original_type in (list, tuple):
            return self.original_type(out[k] for k in sorted(out, key=int))
        return out

    def __eq__(self, other):
        if isinstance(other, FlatterDict):
            return self.as_dict() == other.as_dict()
        if isinstance(other, dict):
            return self.as_dict() == FlatterDict(other, self._delimiter).as_dict()
        return NotImplemented

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    def __repr__(self):
        return f"<FlatterDict {dict(self.items())!r}>"


FlatDict = FlatterDict
