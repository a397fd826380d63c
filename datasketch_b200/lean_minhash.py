"""LeanMinHash with the reference's API (datasketch/lean_minhash.py:8-253).

A frozen MinHash that keeps only ``seed`` and ``hashvalues``; binary layout of
``serialize`` / pickling is the reference's ``struct`` format ``<bo>q i {K}I``
(seed int64, count int32, K x uint32 -- lean_minhash.py:174-175).  Single
objects are packed on the host exactly like the reference; *batches* of
signatures are packed / unpacked by the codec kernels in ``libdsk_b200.so``
(``datasketch_b200.codec.lean_pack`` / ``lean_unpack``).
"""
from __future__ import annotations

import struct
from typing import Iterable, Optional

import numpy as np

from .minhash import MinHash


class LeanMinHash(MinHash):
    """Lean version of MinHash (lean_minhash.py:8-66)."""

    __slots__ = ("seed", "_state")

    def _initialize_slots(self, seed, hashvalues) -> None:
        self.seed = seed
        self._state = self._parse_hashvalues(hashvalues)

    def __init__(self, minhash: Optional[MinHash] = None, seed: Optional[int] = None,
                 hashvalues: Optional[Iterable] = None) -> None:
        if minhash is not None:
            self._initialize_slots(minhash.seed, minhash.hashvalues)
        elif hashvalues is not None and seed is not None:
            self._initialize_slots(seed, hashvalues)
        else:
            raise ValueError("Init parameters cannot be None: make sure to set either minhash or both of "
                             "hash values and seed")

    # a LeanMinHash has no pending queue: its state is always materialised
    @property
    def hashvalues(self) -> np.ndarray:
        return self._state

    @hashvalues.setter
    def hashvalues(self, value) -> None:
        self._state = value

    def _flush(self) -> None:  # nothing is ever queued
        return

    def update(self, b) -> None:
        """Not available on a LeanMinHash (lean_minhash.py:93-97)."""
        raise TypeError("Cannot update a LeanMinHash")

    def update_batch(self, b) -> None:
        raise TypeError("Cannot update a LeanMinHash")

    def copy(self) -> "LeanMinHash":
        # the reference's copy() passes the slot *names* and raises (lean_minhash.py:99-102, untested there);
        # this is the evident intent: an independent copy of seed + hashvalues
        lmh = object.__new__(LeanMinHash)
        lmh._initialize_slots(self.seed, self._state)
        return lmh

    def bytesize(self, byteorder: str = "@") -> int:
        """lean_minhash.py:104-124."""
        return (struct.calcsize(byteorder + "q") + struct.calcsize(byteorder + "i")
                + len(self) * struct.calcsize(byteorder + "I"))

    def serialize(self, buf, byteorder: str = "@") -> None:
        """lean_minhash.py:126-175."""
        if len(buf) < self.bytesize():
            raise ValueError("The buffer does not have enough space for holding this MinHash.")
        fmt = "%sqi%dI" % (byteorder, len(self))
        struct.pack_into(fmt, buf, 0, self.seed, len(self), *self._state)

    @classmethod
    def deserialize(cls, buf, byteorder: str = "@") -> "LeanMinHash":
        """lean_minhash.py:177-214."""
        fmt_head = "%sqi" % byteorder
        try:
            seed, num_perm = struct.unpack_from(fmt_head, buf, 0)
        except TypeError:
            buf = memoryview(buf)
            seed, num_perm = struct.unpack_from(fmt_head, buf, 0)
        offset = struct.calcsize(fmt_head)
        values = struct.unpack_from(byteorder + "%dI" % num_perm, buf, offset)
        lmh = object.__new__(LeanMinHash)
        lmh._initialize_slots(seed, values)
        return lmh

    def __getstate__(self):
        buf = bytearray(self.bytesize())
        struct.pack_into("qi%dI" % len(self), buf, 0, self.seed, len(self), *self._state)
        return buf

    def __setstate__(self, buf) -> None:
        try:
            seed, num_perm = struct.unpack_from("qi", buf, 0)
        except TypeError:
            buf = memoryview(buf)
            seed, num_perm = struct.unpack_from("qi", buf, 0)
        values = struct.unpack_from("%dI" % num_perm, buf, struct.calcsize("qi"))
        self._initialize_slots(seed, values)

    def __hash__(self) -> int:
        return hash((self.seed, tuple(self._state)))

    def __eq__(self, other) -> bool:
        return (type(self) is type(other) and self.seed == other.seed
                and np.array_equal(self._state, other._state))

    def __len__(self) -> int:
        return len(self._state)

    @classmethod
    def union(cls, *lmhs: "LeanMinHash") -> "LeanMinHash":
        """lean_minhash.py:237-253."""
        if len(lmhs) < 2:
            raise ValueError("Cannot union less than 2 MinHash")
        num_perm = len(lmhs[0])
        seed = lmhs[0].seed
        if any((seed != m.seed or num_perm != len(m)) for m in lmhs):
            raise ValueError("The unioning MinHash must have the same seed, number of permutation functions.")
        lmh = object.__new__(LeanMinHash)
        lmh._initialize_slots(seed, np.minimum.reduce([m.hashvalues for m in lmhs]))
        return lmh
