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


class Swish(nn.Module):
    """x * sigmoid(beta x) with an optionally trainable beta (reference nn/activation.py:7-33)."""

    def __init__(self, beta: float = 1.0, train_beta: bool = True):
        super().__init__()
        self.train_beta = train_beta
        if train_beta:
            self.beta = nn.Parameter(torch.tensor(float(beta)))
        else:
            self.register_buffer("beta", torch.tensor(float(beta)))

    def extra_repr(self) -> str:
        return f"beta={self.beta.item():.2f}, trainable_beta:{self.train_beta}"

    def forward(self, x):
        return x * torch.sigmoid(self.beta * x)


class ShiftedSoftplus(nn.Module):
    """softplus(x) - ln 2 (reference nn/activation.py:36-65)."""

    def __init__(self, shift: float = math.log(2.0)):
        super().__init__()
        self.shift = shift

    def extra_repr(self) -> str:
        return f"shift={self.shift:.2f}"

    def forward(self, x):
        return torch.nn.functional.softplus(x) - self.shift


def activation_resolver(query) -> nn.Module:
    """torch.nn activations plus the reference's Swish and ShiftedSoftplus (utils/resolve.py:65-76)."""
    if isinstance(query, nn.Module):
        return query
    q = _norm(query)
    classes = dict(vars(torch.nn.modules.activation), Swish=Swish, ShiftedSoftplus=ShiftedSoftplus)
    for name, cls in classes.items():
        if isinstance(cls, type) and issubclass(cls, nn.Module) and _norm(name) == q:
            return cls()
    raise ValueError(f"{query} not found")


def activation_code(act: nn.Module) -> int:
    """LCAO_ACT_* code of an activation module, for the kinds the kernels fuse: parameter-free activations with their
    default hyper-parameters.  Anything else (Swish with a trainable or non-unit beta, PReLU, ...) raises."""
    from ._lib import ACT

    def default(**kw):
        return all(getattr(act, k) == v for k, v in kw.items())

    if isinstance(act, nn.SiLU):
        return ACT["silu"]
    if isinstance(act, Swish):  # x sigmoid(beta x) = SiLU(beta x) / beta: the callers scale the layer around the kernel
        return ACT["silu"]      # (`swish_beta`), so that beta stays an ordinary, trainable tensor
    if isinstance(act, ShiftedSoftplus) and abs(act.shift - math.log(2.0)) < 1e-12:
        return ACT["shiftedsoftplus"]
    if isinstance(act, nn.Softplus) and default(beta=1.0, threshold=20.0):
        return ACT["softplus"]
    if type(act) is nn.ReLU:
        return ACT["relu"]
    if type(act) is nn.Tanh:
        return ACT["tanh"]
    if type(act) is nn.Sigmoid:
        return ACT["sigmoid"]
    if isinstance(act, nn.GELU) and default(approximate="none"):
        return ACT["gelu"]
    if isinstance(act, nn.ELU) and default(alpha=1.0):
        return ACT["elu"]
    if isinstance(act, nn.LeakyReLU) and default(negative_slope=0.01):
        return ACT["leakyrelu"]
    raise NotImplementedError(f"activation {act!r}: the B200 kernels fuse SiLU (the reference default, lcaonet.py:345), "
                              "ShiftedSoftplus, Softplus, ReLU, Tanh, Sigmoid, GELU, ELU and LeakyReLU with default "
                              "hyper-parameters, and Swish (trainable beta included)")


def swish_beta(act: nn.Module):
    """beta of a Swish activation (a 0-d tensor, possibly a Parameter), or None for every other activation.
    Swish_beta(W x + b) = SiLU((beta W) x + beta b) / beta: a Dense layer followed by Swish is evaluated by the SiLU
    kernels on beta-scaled weights, and the 1/beta goes into whatever consumes the result — beta enters only through
    small torch ops on weights, so autograd delivers d loss / d beta without any kernel knowing about it."""
    return act.beta if isinstance(act, Swish) else None


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
