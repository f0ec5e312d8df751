"""ORACLE (test infrastructure only — never imported by the product path).

Pippenger multi-exponentiation as bellman 0.14.0 `multiexp::multiexp` performs it (bellman is an
un-vendored crates.io dependency, /root/reference/Cargo.toml:27; algorithm restated from the
published source):
  * window c = 3 if n < 32 else ceil(ln n)
  * per window (lowest first): zero exponents skipped; exponent == 1 added straight to the
    accumulator in the first window only; otherwise bucket[(exp >> skip) mod 2^c - 1] += base
  * bucket sum by "summation by parts" (running sum from the top bucket down)
  * windows combined top-down: higher = 2^c * higher + this
The result is a group element, so any correct evaluation order yields the same affine bytes.
"""
import math
from . import curve as C
from .field import R_MOD

NUM_BITS = 255


def bellman_window(n: int) -> int:
    return 3 if n < 32 else int(math.ceil(math.log(n)))


def multiexp(F, bases, exps):
    n = len(exps)
    c = bellman_window(n)
    exps = [e % R_MOD for e in exps]

    def inner(skip, handle_trivial):
        acc = C.to_jac(F, None)
        buckets = [C.to_jac(F, None) for _ in range((1 << c) - 1)]
        for b, e in zip(bases, exps):
            if e == 0 or b is None:
                continue
            if e == 1:
                if handle_trivial:
                    acc = C._jadd(F, acc, C.to_jac(F, b))
                continue
            d = (e >> skip) % (1 << c)
            if d:
                buckets[d - 1] = C._jadd(F, buckets[d - 1], C.to_jac(F, b))
        run = C.to_jac(F, None)
        for bk in reversed(buckets):
            run = C._jadd(F, run, bk)
            acc = C._jadd(F, acc, run)
        skip += c
        if skip >= NUM_BITS:
            return acc
        higher = inner(skip, False)
        for _ in range(c):
            higher = C._jdbl(F, higher)
        return C._jadd(F, higher, acc)

    return C.from_jac(F, inner(0, True))
