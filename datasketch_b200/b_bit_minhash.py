"""b-bit MinHash with the reference's API (datasketch/b_bit_minhash.py:13-172) -- a "next" row of
SURVEY.md section 8(f): a pure codec over a finished signature.

Single objects are handled on the host exactly like the reference (same pickle bytes); whole
signature matrices are masked and packed into the same 64-bit blocks by ``dsk_bbit_pack``
(``datasketch_b200.codec.bbit_pack``).
"""
from __future__ import annotations

import struct

import numpy as np


class bBitMinHash:
    """Keeps the b lowest bits of every hash value of a MinHash (b_bit_minhash.py:13-40)."""

    __slots__ = ("b", "hashvalues", "r", "seed")

    _serial_fmt_params = "<qBdi"
    _serial_fmt_block = "Q"

    def __init__(self, minhash, b: int = 1, r: float = 0.0):
        b = int(b)
        r = float(r)
        if b > 32 or b < 0:
            raise ValueError("b must be an integer in [0, 32]")
        if r > 1.0:
            raise ValueError("r must be a float in [0.0, 1.0]")
        bmask = (1 << b) - 1
        self.hashvalues = np.bitwise_and(minhash.hashvalues, bmask).astype(np.uint32)
        self.seed = minhash.seed
        self.b = b
        self.r = r

    def __eq__(self, other) -> bool:
        return (type(self) is type(other) and self.seed == other.seed and self.b == other.b and self.r == other.r
                and np.array_equal(self.hashvalues, other.hashvalues))

    __hash__ = None

    def jaccard(self, other: "bBitMinHash") -> float:
        """b_bit_minhash.py:53-72: collision-corrected estimate."""
        if self.b != other.b:
            raise ValueError("Cannot compare two b-bit MinHashes with different b values")
        if self.seed != other.seed:
            raise ValueError("Cannot compare two b-bit MinHashes with different set of permutations")
        intersection = np.count_nonzero(self.hashvalues == other.hashvalues)
        raw_est = float(intersection) / float(self.hashvalues.size)
        a1 = self._calc_a(self.r, self.b)
        a2 = self._calc_a(other.r, other.b)
        c1, c2 = self._calc_c(a1, a2, self.r, other.r)
        return (raw_est - c1) / (1 - c2)

    def bytesize(self) -> int:
        return self._bytesize()[-1]

    def __getstate__(self):
        """Header '<qBdi' + 64-bit blocks, slot j of a block at bit (n-1-j)*slot_size (b_bit_minhash.py:78-92)."""
        slot_size, n, num_blocks, total = self._bytesize()
        buffer = bytearray(total)
        blocks = [0] * num_blocks
        hv = [int(x) for x in self.hashvalues]
        for i in range(num_blocks):
            blk = 0
            for j, v in enumerate(hv[i * n:(i + 1) * n]):
                blk |= v << (n - 1 - j) * slot_size
            blocks[i] = blk
        fmt = self._serial_fmt_params + "%d%s" % (num_blocks, self._serial_fmt_block)
        struct.pack_into(fmt, buffer, 0, self.seed, self.b, self.r, self.hashvalues.size, *blocks)
        return buffer

    def __setstate__(self, buf) -> None:
        try:
            self.seed, self.b, self.r, num_perm = struct.unpack_from(self._serial_fmt_params, buf, 0)
        except TypeError:
            buf = memoryview(buf)
            self.seed, self.b, self.r, num_perm = struct.unpack_from(self._serial_fmt_params, buf, 0)
        offset = struct.calcsize(self._serial_fmt_params)
        self.hashvalues = np.zeros((num_perm,), dtype=np.uint32)
        slot_size, n, num_blocks, _total = self._bytesize()
        blocks = struct.unpack_from("%d%s" % (num_blocks, self._serial_fmt_block), buf, offset)
        mask = (1 << slot_size) - 1
        for i in range(num_blocks):
            for j in range(min(n, num_perm - i * n)):
                self.hashvalues[i * n + j] = np.uint32((blocks[i] >> (n - 1 - j) * slot_size) & mask)

    def _calc_a(self, r: float, b: int) -> float:
        if r == 0.0:
            return 1.0 / (1 << b)
        return r * (1 - r) ** (2 ** b - 1) / (1 - (1 - r) ** (2 * b))

    def _calc_c(self, a1: float, a2: float, r1: float, r2: float):
        if r1 == 0.0 and r2 == 0.0:
            return a1, a2
        div = 1 / (r1 + r2)
        return (a1 * r2 + a2 * r1) * div, (a1 * r1 + a2 * r2) * div

    @staticmethod
    def _find_slot_size(b: int) -> int:
        for s in (1, 2, 4, 8, 16, 32):
            if b <= s and not (s == 1 and b != 1) and not (s == 2 and b != 2):
                return s
        raise ValueError("Incorrect value of b")

    def _bytesize(self):
        block_size = struct.calcsize(self._serial_fmt_block)
        slot_size = self._find_slot_size(self.b)
        n = int(block_size * 8 / slot_size)
        num_blocks = int(np.ceil(float(self.hashvalues.size) / n))
        total = struct.calcsize(self._serial_fmt_params + "%d%s" % (num_blocks, self._serial_fmt_block))
        return slot_size, n, num_blocks, total
