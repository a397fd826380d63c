"""MinHashLSHEnsemble (containment search) with the reference's API (datasketch/lshensemble.py:61-265,
datasketch/lshensemble_partition.py), on top of this package's ``MinHashLSH``.

What runs where:

* parameter tables (``_optimal_param``: scipy ``quad`` over ten x/q ratios) and the size partitioner
  (``optimal_partitions``: dynamic programme over the distinct set sizes) are host-side control logic, float64,
  the same expressions in the same order as the reference so the chosen (b, r) and partition bounds are identical;
* ``index`` inserts each partition's sets into its ``MinHashLSH`` instances with ``insert_batch``: the band keys of
  the whole partition come from one ``dsk_band_keys`` launch per distinct r instead of one ``_H`` call per set
  and band (lshensemble.py:221-228 -> lsh.py:344);
* ``query`` probes the first b bands of each partition's index (``MinHashLSH._query_b``), as the reference does.
"""
from __future__ import annotations

import random
import string
import struct
from collections import Counter
from typing import Dict, Generator, Hashable, Iterable, List, Optional, Tuple

import numpy as np
from scipy.integrate import quad as integrate

from .lsh import MinHashLSH
from .minhash import MinHash


def _random_name(length: int) -> bytes:
    """Default storage basename: random lowercase letters, as bytes (storage.py:1052-1054)."""
    return "".join(random.choice(string.ascii_lowercase) for _ in range(length)).encode("utf8")


# ---- (b, r) for a containment threshold and a size ratio xq = x / q (lshensemble.py:17-58) -------------------
def _false_positive_probability(threshold: float, b: int, r: int, xq: float) -> float:
    """Area under the collision-probability curve below the threshold; containment t maps to the Jaccard
    similarity t / (1 + xq - t), and t cannot exceed xq."""
    prob = lambda t: 1 - (1 - (t / (1 + xq - t)) ** float(r)) ** float(b)  # noqa: E731
    area, _ = integrate(prob, 0.0, threshold if xq >= threshold else xq)
    return area


def _false_negative_probability(threshold: float, b: int, r: int, xq: float) -> float:
    prob = lambda t: 1 - (1 - (1 - (t / (1 + xq - t)) ** float(r)) ** float(b))  # noqa: E731
    if xq >= 1.0:
        return integrate(prob, threshold, 1.0)[0]
    if xq >= threshold:
        return integrate(prob, threshold, xq)[0]
    return 0.0


def _optimal_param(threshold: float, num_perm: int, max_r: int, xq: float, false_positive_weight: float,
                   false_negative_weight: float) -> Tuple[int, int]:
    """Grid search over b in [1, num_perm], r in [1, max_r], b*r <= num_perm; first strict minimum wins."""
    best, opt = float("inf"), (0, 0)
    for b in range(1, num_perm + 1):
        for r in range(1, max_r + 1):
            if b * r > num_perm:
                continue
            err = (_false_positive_probability(threshold, b, r, xq) * false_positive_weight
                   + _false_negative_probability(threshold, b, r, xq) * false_negative_weight)
            if err < best:
                best, opt = err, (b, r)
    return opt


# ---- partitioning of the set-size domain (lshensemble_partition.py) -------------------------------------------
def _interval_costs(sizes: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """cost[l, u] = expected false positives when every set with size in [sizes[l], sizes[u]] is treated as
    having size sizes[u] (lshensemble_partition.py:54-92): sum_i (sizes[u] - sizes[i]) / sizes[u] * counts[i]."""
    n = len(sizes)
    cost = np.zeros((n, n))
    for u in range(n):
        top = float(sizes[u])
        for l in range(u + 1):
            cost[l, u] = np.sum((top - sizes[l:u + 1]) / top * counts[l:u + 1])
    return cost


def optimal_partitions(sizes, counts, num_part: int) -> List[Tuple[int, int]]:
    """Cut the ascending domain of set sizes into ``num_part`` contiguous intervals ``(lower, upper)`` (inclusive)
    minimising the summed interval cost; ties go to the smaller cut index, as in the reference
    (lshensemble_partition.py:95-196)."""
    sizes = np.asarray(sizes)
    counts = np.asarray(counts)
    n = len(sizes)
    if num_part < 2:
        return [(sizes[0], sizes[-1])]
    if num_part >= n:
        return [(x, x) for x in sizes]
    cost = _interval_costs(sizes, counts)
    last = n - 1
    if num_part == 2:
        tot = cost[0, :last] + cost[1:, last]                      # first interval ends at u1, second is the rest
        u = int(np.argmin(tot))                                    # first minimum = smallest u1
        return [(sizes[0], sizes[u]), (sizes[u + 1], sizes[-1])]
    # tab[p][u] (p = 2 .. num_part-1): cheapest cover of sizes[0..u] by p intervals; entries with u < p-1 stay 0,
    # exactly like the reference's zero-initialised matrix (lshensemble_partition.py:130-146)
    tab: Dict[int, np.ndarray] = {}
    for p in range(2, num_part):
        t = np.zeros(n)
        prev = cost[0] if p == 2 else tab[p - 1]
        for u in range(p - 1, n):
            t[u] = np.min(prev[p - 2:u] + cost[p - 1:u + 1, u])    # u1 = p-2 .. u-1, next interval = u1+1 .. u
        tab[p] = t
    p = num_part
    u = p - 2 + int(np.argmin(tab[p - 1][p - 2:last] + cost[p - 1:last + 1, last]))
    bounds = [(sizes[u + 1], sizes[-1])]
    p -= 1
    while p > 1:
        # NOTE: the reference's back-tracking reads the p-interval table here, where the recurrence that filled the
        # tables used p-1 (lshensemble_partition.py:158-166).  Reproduced as is: the partition bounds have to be the
        # reference's, not the optimum of the recurrence.
        u1 = p - 2 + int(np.argmin(tab[p][p - 2:u] + cost[p - 1:u + 1, u]))
        bounds.insert(0, (sizes[u1 + 1], sizes[u]))
        u = u1
        p -= 1
    bounds.insert(0, (sizes[0], sizes[u]))
    return bounds


# ---- the index ------------------------------------------------------------------------------------------------
class MinHashLSHEnsemble:
    """LSH Ensemble for containment queries (constructor contract of lshensemble.py:108-153)."""

    def __init__(self, threshold: float = 0.9, num_perm: int = 128, num_part: int = 16, m: int = 8,
                 weights: Tuple[float, float] = (0.5, 0.5), storage_config: Optional[dict] = None,
                 prepickle: Optional[bool] = None) -> None:
        if threshold > 1.0 or threshold < 0.0:
            raise ValueError("threshold must be in [0.0, 1.0]")
        if num_perm < 2:
            raise ValueError("Too few permutation functions")
        if num_part < 1:
            raise ValueError("num_part must be at least 1")
        if m < 2 or m > num_perm:
            raise ValueError("m must be in the range of [2, num_perm]")
        if any(w < 0.0 or w > 1.0 for w in weights):
            raise ValueError("Weight must be in [0.0, 1.0]")
        if sum(weights) != 1.0:
            raise ValueError("Weights must sum to 1.0")
        self.threshold = threshold
        self.h = num_perm
        self.m = m
        rs = self._init_optimal_params(weights)
        storage_config = storage_config if storage_config else {"type": "dict"}
        basename = storage_config.get("basename", _random_name(11))
        self.indexes = [
            {r: MinHashLSH(num_perm=self.h, params=(int(self.h / r), r),
                           storage_config=self._get_storage_config(basename, storage_config, part, r),
                           prepickle=prepickle) for r in rs}
            for part in range(num_part)]
        self.lowers = [None for _ in self.indexes]
        self.uppers = [None for _ in self.indexes]

    def _init_optimal_params(self, weights):
        fpw, fnw = weights
        self.xqs = np.exp(np.linspace(-5, 5, 10))
        self.params = np.array([_optimal_param(self.threshold, self.h, self.m, xq, fpw, fnw) for xq in self.xqs],
                               dtype=int)
        return {r for _, r in self.params}

    def _get_optimal_param(self, x, q):
        i = np.searchsorted(self.xqs, float(x) / float(q), side="left")
        if i == len(self.params):
            i -= 1
        return self.params[i]

    def _get_storage_config(self, basename, base_config, partition, r):
        config = dict(base_config)
        config["basename"] = b"-".join([basename, struct.pack(">H", partition), struct.pack(">H", r)])
        return config

    def index(self, entries: Iterable[Tuple[Hashable, MinHash, int]]) -> None:
        """Index all sets given as ``(key, minhash, size)``; callable once (lshensemble.py:189-228)."""
        if not self.is_empty():
            raise ValueError("Cannot call index again on a non-empty index")
        if not isinstance(entries, list):
            checked = []
            for key, minhash, size in entries:
                if size <= 0:
                    raise ValueError("Set size must be positive")
                checked.append((key, minhash, size))
            entries = checked
        if len(entries) == 0:
            raise ValueError("entries is empty")
        sizes, counts = np.array(sorted(Counter(e[2] for e in entries).most_common())).T
        partitions = optimal_partitions(sizes, counts, len(self.indexes))
        for i, (lower, upper) in enumerate(partitions):
            self.lowers[i], self.uppers[i] = lower, upper
        entries.sort(key=lambda e: e[2])
        # same walk as the reference (one step to the next partition when a size exceeds the current upper bound),
        # but the partition's sets are inserted together so their band keys come from one kernel launch per r
        groups: List[list] = [[] for _ in self.indexes]
        curr = 0
        for e in entries:
            if e[2] > self.uppers[curr]:
                curr += 1
            groups[curr].append(e)
        for part, members in enumerate(groups):
            if not members:
                continue
            keys = [e[0] for e in members]
            mhs = [e[1] for e in members]
            batchable = all(np.ndim(mh.hashvalues) == 1 for mh in mhs)
            for r, lsh in self.indexes[part].items():
                if batchable:
                    lsh.insert_batch(keys, mhs)
                else:  # WeightedMinHash rows are (k, t) pairs: per-object path
                    for k, mh in zip(keys, mhs):
                        lsh.insert(k, mh)

    def query(self, minhash, size: int) -> Generator[Hashable, None, None]:
        """Keys of sets whose containment of the query exceeds the threshold (lshensemble.py:230-249)."""
        for i, index in enumerate(self.indexes):
            u = self.uppers[i]
            if u is None:
                continue
            b, r = self._get_optimal_param(u, size)
            for key in index[r]._query_b(minhash, b):
                yield key

    def __contains__(self, key: Hashable) -> bool:
        return any(any(key in index[r] for r in index) for index in self.indexes)

    def is_empty(self) -> bool:
        return all(all(index[r].is_empty() for r in index) for index in self.indexes)
