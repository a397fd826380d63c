"""b-bit MinHash with the reference's API (datasketch/b_bit_minhash.py:13-172) -- a "next" row of
SURVEY.md section 8(f): a pure codec over a finished signature.

A single object is packed on the host with vectorised numpy shifts (same pickle bytes as the reference,
pinned by tests/golden/bbit.npz); whole signature matrices are masked and packed into the same 64-bit
blocks on the device by ``dsk_bbit_pack`` (``datasketch_b200.codec.bbit_pack``).

Wire format (b_bit_minhash.py:22-26, :78-92): header ``<qBdi`` = seed, b, r, num_perm, then ``ceil(K / n)``
little-endian uint64 blocks, ``n = 64 / slot`` values per block, value j of a block at bit ``(n-1-j)*slot``;
slot = b rounded up to 4, 8, 16 or 32 (1 and 2 stay).
"""
from __future__ import annotations

import struct

import numpy as np

_HEADER = struct.Struct("<qBdi")


def _slot_bits(b: int) -> int:
    """Storage width of one b-bit value."""
    if b in (1, 2):
        return b
    for width in (4, 8, 16, 32):
        if b <= width:
            return width
    raise ValueError("Incorrect value of b")


class bBitMinHash:
    """Keeps the b lowest bits of every hash value of a MinHash (b_bit_minhash.py:13-40)."""

    __slots__ = ("b", "hashvalues", "r", "seed")

    def __init__(self, minhash, b: int = 1, r: float = 0.0):
        b, r = int(b), float(r)
        if not 0 <= b <= 32:
            raise ValueError("b must be an integer in [0, 32]")
        if r > 1.0:
            raise ValueError("r must be a float in [0.0, 1.0]")
        self.seed = minhash.seed
        self.b, self.r = b, r
        self.hashvalues = (np.asarray(minhash.hashvalues) & ((1 << b) - 1)).astype(np.uint32)

    # -- comparison ---------------------------------------------------------------------------------
    def __eq__(self, other) -> bool:
        if type(self) is not type(other):
            return False
        return ((self.seed, self.b, self.r) == (other.seed, other.b, other.r)
                and np.array_equal(self.hashvalues, other.hashvalues))

    __hash__ = None

    @staticmethod
    def _accidental(r: float, b: int) -> float:
        """A(r, b) of Li & Koenig: probability that two b-bit values agree by chance; r -> 0 gives 2^-b."""
        if r == 0.0:
            return 1.0 / (1 << b)
        return r * (1 - r) ** (2 ** b - 1) / (1 - (1 - r) ** (2 * b))

    def jaccard(self, other: "bBitMinHash") -> float:
        """Collision-corrected resemblance estimate (b_bit_minhash.py:53-72, :118-141)."""
        if self.b != other.b:
            raise ValueError("Cannot compare two b-bit MinHashes with different b values")
        if self.seed != other.seed:
            raise ValueError("Cannot compare two b-bit MinHashes with different set of permutations")
        agree = float(np.count_nonzero(self.hashvalues == other.hashvalues)) / float(self.hashvalues.size)
        a_self, a_other = self._accidental(self.r, self.b), self._accidental(other.r, other.b)
        if self.r == 0.0 and other.r == 0.0:
            c1, c2 = a_self, a_other
        else:
            w = 1 / (self.r + other.r)
            c1 = (a_self * other.r + a_other * self.r) * w
            c2 = (a_self * self.r + a_other * other.r) * w
        return (agree - c1) / (1 - c2)

    # -- serialisation ----------------------------------------------------------------------------------
    def _layout(self):
        """(slot bits, values per block, number of blocks, total bytes) for this object's b and K."""
        slot = _slot_bits(self.b)
        per_block = 64 // slot
        n_blocks = -(-int(self.hashvalues.size) // per_block)
        return slot, per_block, n_blocks, _HEADER.size + 8 * n_blocks

    def bytesize(self) -> int:
        return self._layout()[3]

    def __getstate__(self) -> bytearray:
        slot, per_block, n_blocks, total = self._layout()
        vals = np.zeros(n_blocks * per_block, dtype=np.uint64)
        vals[: self.hashvalues.size] = self.hashvalues
        shifts = (np.arange(per_block - 1, -1, -1, dtype=np.uint64) * np.uint64(slot))
        blocks = np.bitwise_or.reduce(vals.reshape(n_blocks, per_block) << shifts[None, :], axis=1) \
            if n_blocks else np.zeros(0, dtype=np.uint64)
        out = bytearray(total)
        _HEADER.pack_into(out, 0, self.seed, self.b, self.r, int(self.hashvalues.size))
        out[_HEADER.size:] = blocks.astype("<u8").tobytes()
        return out

    def __setstate__(self, buf) -> None:
        view = memoryview(buf)
        self.seed, self.b, self.r, num_perm = _HEADER.unpack_from(view, 0)
        self.hashvalues = np.zeros(num_perm, dtype=np.uint32)      # _layout() reads its size
        slot, per_block, n_blocks, _ = self._layout()
        blocks = np.frombuffer(view, dtype="<u8", count=n_blocks, offset=_HEADER.size).astype(np.uint64)
        shifts = (np.arange(per_block - 1, -1, -1, dtype=np.uint64) * np.uint64(slot))
        vals = (blocks[:, None] >> shifts[None, :]) & np.uint64((1 << slot) - 1)
        self.hashvalues = vals.reshape(-1)[:num_perm].astype(np.uint32)
