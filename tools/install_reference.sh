#!/bin/bash
# Installs the UNMODIFIED reference (ekzhu/datasketch, /root/reference) into baseline/_ref (git-ignored; it
# travels to the GPU box with the snapshot) so that `bench.py --impl reference` and bench.py's cpu_baseline
# leg can time the reference's own MinHash.bulk.  The reference's build backend (hatchling) is not in the
# offline wheelhouse, so `pip install /root/reference` fails; the package is pure Python, so this script
# zips a wheel by hand from a scratch copy (nothing is written to /root/reference) and lets pip install
# THAT wheel with --no-index --no-deps --target baseline/_ref.  Build container only.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF=${1:-/root/reference}
[ -d "$REF/datasketch" ] || { echo "no reference checkout at $REF"; exit 1; }
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
VER=$(sed -n 's/^version = "\(.*\)"/\1/p' "$REF/pyproject.toml" | head -1)
mkdir -p "$TMP/w" && cp -r "$REF/datasketch" "$TMP/w/datasketch"
find "$TMP/w" -name __pycache__ -prune -exec rm -rf {} +
DI="$TMP/w/datasketch-$VER.dist-info"; mkdir -p "$DI"
printf 'Metadata-Version: 2.1\nName: datasketch\nVersion: %s\nRequires-Dist: numpy>=1.11\nRequires-Dist: scipy>=1.0.0\n' "$VER" > "$DI/METADATA"
printf 'Wheel-Version: 1.0\nGenerator: tools/install_reference.sh\nRoot-Is-Purelib: true\nTag: py3-none-any\n' > "$DI/WHEEL"
( cd "$TMP/w" && find . -type f ! -name RECORD | sed 's#^\./##' | awk '{print $0",,"}' > "$DI/RECORD" && echo "datasketch-$VER.dist-info/RECORD,," >> "$DI/RECORD" )
( cd "$TMP/w" && python -m zipfile -c "$TMP/datasketch-$VER-py3-none-any.whl" datasketch "datasketch-$VER.dist-info" )
rm -rf "$REPO/baseline/_ref"
python -m pip install --no-index --no-deps --target "$REPO/baseline/_ref" "$TMP/datasketch-$VER-py3-none-any.whl"
python - <<PY
import sys; sys.path.insert(0, "$REPO/baseline/_ref")
import datasketch; m = datasketch.MinHash(4, 1); m.update(b"Hello")
assert m.hashvalues.tolist() == [734825475, 960773806, 359816889, 342714745]
print("baseline/_ref: datasketch", datasketch.__version__, "golden vector ok")
PY
