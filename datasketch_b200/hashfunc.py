"""Host-side token hash functions (datasketch/hashfunc.py:5-28) plus the two non-cryptographic hashes the
reference documents as ``hashfunc`` alternatives (docs/minhash.rst:79-112).

Hashing of raw tokens is host-side in the reference even in its own GPU mode (datasketch/minhash.py:85-87,
:262-263); the device path starts at integer token hashes.  ``MinHash.bulk`` recognises the functions of this
module and hashes whole batches on the device instead (``dsk_sha1_tokens`` / ``dsk_hash_tokens``) -- same values,
no per-token Python call.  Each function here is the single-token definition those kernels are tested against.
"""
import hashlib
import struct

_M32 = 0xFFFFFFFF


def sha1_hash32(data) -> int:
    """First four bytes of SHA1(data) read little-endian -> 32-bit int."""
    return struct.unpack("<I", hashlib.sha1(data).digest()[:4])[0]


def sha1_hash64(data) -> int:
    """First eight bytes of SHA1(data) read little-endian -> 64-bit int."""
    return struct.unpack("<Q", hashlib.sha1(data).digest()[:8])[0]


def _rotl(x: int, n: int) -> int:
    return ((x << n) | (x >> (32 - n))) & _M32


def xxh32_hash32(data, seed: int = 0) -> int:
    """XXH32 (xxHash specification, 32-bit variant) == ``xxhash.xxh32_intdigest(data, seed)``."""
    d = bytes(data)
    n = len(d)
    p1, p2, p3, p4, p5 = 2654435761, 2246822519, 3266489917, 668265263, 374761393
    i = 0
    if n >= 16:
        v = [(seed + p1 + p2) & _M32, (seed + p2) & _M32, seed & _M32, (seed - p1) & _M32]
        while i + 16 <= n:
            for j, w in enumerate(struct.unpack_from("<4I", d, i)):
                v[j] = (_rotl((v[j] + w * p2) & _M32, 13) * p1) & _M32
            i += 16
        h = (_rotl(v[0], 1) + _rotl(v[1], 7) + _rotl(v[2], 12) + _rotl(v[3], 18)) & _M32
    else:
        h = (seed + p5) & _M32
    h = (h + n) & _M32
    while i + 4 <= n:
        h = (_rotl((h + struct.unpack_from("<I", d, i)[0] * p3) & _M32, 17) * p4) & _M32
        i += 4
    while i < n:
        h = (_rotl((h + d[i] * p5) & _M32, 11) * p1) & _M32
        i += 1
    h ^= h >> 15
    h = (h * p2) & _M32
    h ^= h >> 13
    h = (h * p3) & _M32
    return h ^ (h >> 16)


def murmur3_hash32(data, seed: int = 0) -> int:
    """MurmurHash3 x86_32 (A. Appleby, public domain) == ``mmh3.hash(data, seed, signed=False)``."""
    d = bytes(data)
    n = len(d)
    c1, c2 = 0xCC9E2D51, 0x1B873593
    h = seed & _M32
    i = 0
    while i + 4 <= n:
        k = (struct.unpack_from("<I", d, i)[0] * c1) & _M32
        k = (_rotl(k, 15) * c2) & _M32
        h = (_rotl(h ^ k, 13) * 5 + 0xE6546B64) & _M32
        i += 4
    tail = n - i
    if tail:
        k = 0
        if tail == 3:
            k ^= d[i + 2] << 16
        if tail >= 2:
            k ^= d[i + 1] << 8
        k ^= d[i]
        k = (_rotl((k * c1) & _M32, 15) * c2) & _M32
        h ^= k
    h ^= n
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & _M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & _M32
    return h ^ (h >> 16)


# hash functions MinHash.bulk can run on the device: function -> (C-ABI route, kind)
DEVICE_HASHES = {sha1_hash32: ("sha1", 0), xxh32_hash32: ("hash", 1), murmur3_hash32: ("hash", 2)}
