"""Host-side token hash functions (datasketch/hashfunc.py:5-28).

Hashing of raw tokens is host-side in the reference even in its own GPU mode
(datasketch/minhash.py:85-87, :262-263); the device path starts at integer
token hashes.
"""
import hashlib
import struct


def sha1_hash32(data) -> int:
    """First four bytes of SHA1(data) read little-endian -> 32-bit int."""
    return struct.unpack("<I", hashlib.sha1(data).digest()[:4])[0]


def sha1_hash64(data) -> int:
    """First eight bytes of SHA1(data) read little-endian -> 64-bit int."""
    return struct.unpack("<Q", hashlib.sha1(data).digest()[:8])[0]
