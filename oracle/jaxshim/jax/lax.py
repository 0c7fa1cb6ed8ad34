"""jax.lax stand-in: Python control flow."""
from . import numpy as _jnp
from ._tree import tree_stack


def scan(f, init, xs, length=None):
    carry = init
    ys = []
    n = length if xs is None else len(xs)
    for i in range(n):
        carry, y = f(carry, None if xs is None else xs[i])
        ys.append(y)
    return carry, tree_stack(ys)


def cond(pred, true_fun, false_fun, *operands, operand=None):
    args = operands if operands else (operand,)
    return true_fun(*args) if bool(pred) else false_fun(*args)
