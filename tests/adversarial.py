"""Structured worst-case token streams for the dynamic suffix automaton (shared by the CPU test of the one-thread
step logic and the GPU parity test).  What each one is for - the logic being exercised is the reference's
add_state, /root/reference/samd/sam/dyn_sam.py:41-67, and transfer_state, :69-78:

  a^k b            the suffix chain of the tail is k states long, and `b` gets an out-edge at every one of them
                   (k = 31, 32, 33: one pass of the lane-parallel chain insert; 63, 64, 65: its second pass and the
                   CHAIN_MAX fallback; 200: far beyond)
  (ab)^k c         the same with two interleaved chains
  Fibonacci word   maximal number of distinct repeats; a clone on almost every append
  de Bruijn        every k-gram once: no long match, every append ends at a different depth
  hub_then_clone   a non-root state collects more than five out-edges (overflow list), then gets split, so the clone
                   must copy the overflow edges in insertion order
  all_split        eight consecutive tokens every one of which splits a state
"""
import numpy as np


def a_k_b(k, a=3, b=4):
    return [a] * k + [b]


def ab_k_c(k, a=3, b=4, c=5):
    return [a, b] * k + [c]


def fibonacci_word(n, a=3, b=4):
    x, y = [a], [a, b]
    while len(y) < n:
        x, y = y, y + x
    return y[:n]


def de_bruijn(k, n):
    """de Bruijn sequence B(k, n) over tokens 3..3+k-1 (standard Lyndon-word construction)."""
    a = [0] * (k * n)
    seq = []

    def db(t, p):
        if t > n:
            if n % p == 0:
                seq.extend(a[1:p + 1])
        else:
            a[t] = a[t - p]
            db(t + 1, p)
            for j in range(a[t - p] + 1, k):
                a[t] = j
                db(t + 1, t)

    db(1, 1)
    return [3 + s for s in seq]


def hub_then_clone():
    """`x y` is followed by seven different tokens, so state("x y") - not the root - holds seven out-edges (two of them
    in the overflow table); then `y` alone is followed by one of them in a new left context, which splits the state:
    the clone inherits all seven in order.  More appends after that walk through the clone."""
    x, y = 3, 4
    s = []
    for t in range(10, 17):
        s += [x, y, t]
    s += [5, y, 12, 5, y, 13, x, y, 14, 5, y, 15, 7]
    for t in range(17, 20):
        s += [x, y, t]
    s += [5, y, 18, 6, y, 19]
    return s


def all_split(rounds=6):
    """Blocks `p q r s t u v w` seen once, then re-entered from ever new one-token left contexts: inside the copied block
    every token's target state is a proper-suffix match -> clone on (almost) every append."""
    block = list(range(10, 18))
    s = [30] + block
    for i in range(rounds):
        s += [40 + i] + block
    return s


def streams():
    out = {}
    for k in (1, 2, 5, 6, 7, 31, 32, 33, 63, 64, 65, 200):
        out[f"a^{k} b"] = a_k_b(k)
        out[f"a^{k} b a^{k} b b a"] = a_k_b(k) + a_k_b(k) + [4, 3]
    for k in (3, 16, 17, 33, 70):
        out[f"(ab)^{k} c"] = ab_k_c(k) + ab_k_c(k // 2 + 1)
    out["fib300"] = fibonacci_word(300)
    out["fib300+tail"] = fibonacci_word(300) + [5] + fibonacci_word(60) + [6, 5]
    out["debruijn(3,5)"] = de_bruijn(3, 5)
    out["debruijn(2,8)x2"] = de_bruijn(2, 8) * 2
    out["debruijn(6,3)+prefix"] = de_bruijn(6, 3) + de_bruijn(6, 3)[:40]
    out["hub_then_clone"] = hub_then_clone()
    out["all_split"] = all_split()
    rng = np.random.default_rng(77)
    out["a-runs"] = sum(([3] * int(rng.integers(1, 80)) + [int(rng.integers(4, 7))] for _ in range(12)), [])
    return out


def chunkings(n, seed):
    """How a stream of n tokens is fed: lists of chunk sizes (1..8 like the decode steps, all at once, one by one)."""
    rng = np.random.default_rng(seed)
    sizes = []
    left = n
    while left > 0:
        c = min(left, int(rng.integers(1, 9)))
        sizes.append(c)
        left -= c
    return {"steps1-8": sizes, "whole": [n], "single": [1] * n, "eights": [8] * (n // 8) + ([n % 8] if n % 8 else [])}
