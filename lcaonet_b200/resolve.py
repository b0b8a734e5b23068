"""String -> object resolvers with the reference's matching rules (lcaonet/utils/resolve.py:16-147):
names are compared lower-cased with '-', '_' and ' ' removed; cutoff names may omit the trailing
"cutoff", radial-basis names the trailing "radialbasis".  Unknown names raise ValueError."""
from __future__ import annotations

import math
from inspect import getmembers, isfunction

import torch
import torch.nn as nn


def _norm(s: str) -> str:
    return s.lower().replace("-", "").replace("_", "").replace(" ", "")


def glorot(value):
    stdv = math.sqrt(6.0 / (value.size(-2) + value.size(-1)))
    value.data.uniform_(-stdv, stdv)


def glorot_orthogonal(tensor, scale):
    """orthogonal_ then rescale so that var(W) = scale / (fan_in + fan_out)
    (torch_geometric.nn.inits.glorot_orthogonal, used by the reference Dense with scale=2.0)."""
    torch.nn.init.orthogonal_(tensor.data)
    scale = scale / ((tensor.size(-2) + tensor.size(-1)) * tensor.var())
    tensor.data *= scale.sqrt()


def init_resolver(query):
    funcs = [f for _, f in getmembers(torch.nn.init, isfunction) if "deprecated" not in str(f)] + [glorot, glorot_orthogonal]
    if callable(query):
        return query
    q = _norm(query)
    for f in funcs:
        if _norm(f.__name__) == q:
            return f
    raise ValueError(f"{query} not found")


def init_params(fn) -> tuple[str, ...]:
    return fn.__code__.co_varnames[: fn.__code__.co_argcount]


def activation_resolver(query) -> nn.Module:
    if isinstance(query, nn.Module):
        return query
    q = _norm(query)
    for name, cls in vars(torch.nn.modules.activation).items():
        if isinstance(cls, type) and issubclass(cls, nn.Module) and _norm(name) == q:
            return cls()
    raise ValueError(f"{query} not found")


def cutoff_kind(query) -> str:
    q = _norm(query) if isinstance(query, str) else _norm(getattr(query, "__name__", str(query)))
    if q.endswith("cutoff"):
        q = q[: -len("cutoff")]
    if q not in ("polynomial", "envelope", "cosine"):
        raise ValueError(f"{query} not found")
    return q


def rbf_kind(query) -> str:
    q = _norm(query) if isinstance(query, str) else _norm(getattr(query, "__name__", str(query)))
    if q.endswith("radialbasis"):
        q = q[: -len("radialbasis")]
    if q not in ("hydrogen", "sphericalbessel"):
        raise ValueError(f"{query} not found")
    return q
