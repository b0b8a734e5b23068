"""Stand-in for torch_geometric.data.{Data,Batch}: a dict-like attribute bag. Test infrastructure."""


class Data:
    def __init__(self, **kw):
        self.__dict__["_store"] = dict(kw)

    def __getitem__(self, k):
        return self._store[k]

    def __setitem__(self, k, v):
        self._store[k] = v

    def __getattr__(self, k):
        try:
            return self.__dict__["_store"][k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self._store[k] = v

    def __contains__(self, k):
        return k in self._store

    def get(self, k, default=None):
        return self._store.get(k, default)

    def keys(self):
        return self._store.keys()

    def to(self, *a, **kw):
        return type(self)(**{k: (v.to(*a, **kw) if hasattr(v, "to") else v) for k, v in self._store.items()})

    def clone(self):
        return type(self)(**{k: (v.clone() if hasattr(v, "clone") else v) for k, v in self._store.items()})


class Batch(Data):
    pass
