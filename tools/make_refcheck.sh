#!/bin/bash
# Build container only.  Stages the reference's OWN unittest files plus an import shim
# (`datasketch` -> datasketch_b200) in the git-ignored directory _refcheck/, so that
# `gpurun -- bash tools/gpu_refcheck.sh` can run them against this library on a B200.
# Nothing under _refcheck/ is ever committed (it is git-ignored); it travels to the GPU box with the snapshot, where
# tests/test_reference_own_tests_gpu.py (-m gpu) runs the staged files against the real library.
set -e
REF=${DATASKETCH_REF:-/root/reference}
cd "$(dirname "$0")/.."
rm -rf _refcheck && mkdir -p _refcheck/datasketch _refcheck/test
for f in __init__.py utils.py test_minhash.py test_lean_minhash.py test_weighted_minhash.py test_lsh.py test_lshforest.py test_lshensemble.py test_lshbloom.py; do
  cp "$REF/test/$f" _refcheck/test/
done
cat > _refcheck/datasketch/__init__.py <<'PY'
from datasketch_b200 import *  # noqa
from datasketch_b200 import (MinHash, LeanMinHash, MinHashLSH, MinHashLSHForest, MinHashLSHEnsemble, WeightedMinHash,
                             WeightedMinHashGenerator, bBitMinHash, MinHashLSHBloom)
PY
for m in minhash lean_minhash weighted_minhash lsh lshforest lshensemble lsh_bloom b_bit_minhash hashfunc; do
cat > _refcheck/datasketch/$m.py <<PY
import datasketch_b200.$m as _m
from datasketch_b200.$m import *  # noqa
globals().update({k: getattr(_m, k) for k in dir(_m) if not k.startswith("__")})
PY
done
cat > _refcheck/mockredis.py <<'PY'
import unittest
def mock_redis_client(**kwargs):
    raise unittest.SkipTest("redis storage is out of scope")
mock_strict_redis_client = mock_redis_client
PY
echo "staged: _refcheck/ (git-ignored)"
