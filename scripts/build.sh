#!/bin/bash
# Build the native libraries in-tree (sm_100a) and a wheel/sdist of the package.
set -e
cd "$(dirname "$0")/.."
python -m blades_b200.ops.build "$@"
python setup.py sdist bdist_wheel 2>/dev/null || echo "(wheel build skipped)"
