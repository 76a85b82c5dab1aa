"""numpy restatement of the synthetic graph generator — TEST INFRASTRUCTURE ONLY.

No reference counterpart: the reference reads external datasets (PA4/handout/src/data.cu:3-66).
Parity with the reference is UNPINNED by construction; this independently written restatement
pins hpc_b200's generator (hpc_b200/csrc/graph.cpp) bit for bit. Definition:

  mix64       splitmix64 finaliser (adds the golden constant first)
  key(s)      mix64(seed ^ mix64(s));   hash2(key, a, b) = mix64(mix64(key + a) ^ b)
  unit(h)     ((h >> 12) + 1) * 2^-52                      in (0, 1]
  degrees     w_i = 0 with probability zero_ppm/1e6 (stream 2), else unit(mix64(key1 + i))^(-tail_k/4)
              via sqrt/mul/div; the heaviest row gets max_deg; the rest get
              min(max_deg, floor(w_i * s)) for the largest s (200-step bisection on [0, 2^40])
              whose sum stays <= nnz - max_deg; the remainder is handed out one per eligible row
              in row order, wrapping.
  columns     candidates k = 0, 1, ... of row r from h = hash2(key3, r, k), h2 = mix64(h):
              local (h2 % 1e6 < local_ppm): (r + (h2 >> 20) % (2*window+1) - window) mod M
              global: rank = floor(M * unit(h)^2), column (rank * mult + 12345) mod M
              first deg distinct candidates (at most 64*deg+64 draws, then the smallest unused
              ids), sorted ascending.
"""
from __future__ import annotations

import math

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(z):
    z = np.asarray(z, np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def key(seed, stream):
    return mix64(np.uint64(seed) ^ mix64(np.uint64(stream)))


def unit(h):
    return ((h >> np.uint64(12)) + np.uint64(1)).astype(np.float64) * (1.0 / 4503599627370496.0)


def tail_weight(u, k):
    s = np.sqrt(u)
    q = np.sqrt(s)
    return {1: 1.0 / q, 2: 1.0 / s, 3: 1.0 / (s * q), 4: 1.0 / u}[k]


def gen_degrees(m, nnz, max_deg, tail_k, zero_ppm, seed):
    i = np.arange(m, dtype=np.uint64)
    with np.errstate(over="ignore"):
        empty = (mix64(key(seed, 2) + i) % np.uint64(1000000)) < np.uint64(zero_ppm)
        w = tail_weight(unit(mix64(key(seed, 1) + i)), tail_k)
    w[empty] = 0.0
    top = int(np.argmax(w))  # first maximum
    elig = w != 0.0
    elig[top] = False
    want = nnz - max_deg

    def total(s):
        return int(np.minimum(np.floor(w[elig] * s), float(max_deg)).astype(np.int64).sum())

    lo, hi = 0.0, 1099511627776.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if mid == lo or mid == hi:
            break
        if total(mid) <= want:
            lo = mid
        else:
            hi = mid
    deg = np.zeros(m, np.int64)
    deg[elig] = np.minimum(np.floor(w[elig] * lo), float(max_deg)).astype(np.int64)
    rem = want - int(deg.sum())
    deg[top] = max_deg
    while rem > 0:
        cand = np.flatnonzero(elig & (deg < max_deg))
        if len(cand) == 0:
            raise ValueError("cannot place remainder")
        take = cand[:rem]
        deg[take] += 1
        rem -= len(take)
    return deg.astype(np.int32)


def scatter_mult(m):
    a = max(1, int(float(m) * 0.6180339887498949))
    while math.gcd(a, m) != 1:
        a += 1
    return a


def candidates(seed, m, local_ppm, window, r, k0, k1):
    k = np.arange(k0, k1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = mix64(mix64(key(seed, 3) + np.uint64(r)) ^ k)
        h2 = mix64(h)
        local = (h2 % np.uint64(1000000)) < np.uint64(local_ppm)
        off = ((h2 >> np.uint64(20)) % np.uint64(2 * window + 1)).astype(np.int64) - window
        cl = (r + off) % m
        v = unit(h)
        rank = np.minimum(np.floor(float(m) * (v * v)).astype(np.uint64), np.uint64(m - 1))
        cg = ((rank * np.uint64(scatter_mult(m)) + np.uint64(12345)) % np.uint64(m)).astype(np.int64)
    return np.where(local, cl, cg)


def gen_row(seed, m, local_ppm, window, r, d):
    if d == 0:
        return np.zeros(0, np.int32)
    max_draws = 64 * d + 64
    seen, out, k0 = set(), [], 0
    while len(out) < d and k0 < max_draws:
        k1 = min(max_draws, k0 + max(64, 2 * (d - len(out))))
        for c in candidates(seed, m, local_ppm, window, r, k0, k1).tolist():
            if c not in seen:
                seen.add(c)
                out.append(c)
                if len(out) == d:
                    break
        k0 = k1
    if len(out) < d:
        for c in range(m):
            if c not in seen:
                seen.add(c)
                out.append(c)
                if len(out) == d:
                    break
    return np.sort(np.asarray(out, np.int32))


def gen_graph(m, nnz, max_deg, tail_k, zero_ppm, local_ppm, window, seed, rows=None):
    """Full graph (rows=None) or only the listed rows' column lists (dict row -> array)."""
    deg = gen_degrees(m, nnz, max_deg, tail_k, zero_ppm, seed)
    ptr = np.zeros(m + 1, np.int32)
    np.cumsum(deg, out=ptr[1:])
    if rows is not None:
        return ptr, {int(r): gen_row(seed, m, local_ppm, window, int(r), int(deg[r])) for r in rows}
    idx = np.empty(nnz, np.int32)
    for r in range(m):
        idx[ptr[r]:ptr[r + 1]] = gen_row(seed, m, local_ppm, window, r, int(deg[r]))
    return ptr, idx
