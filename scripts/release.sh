#!/bin/bash
# Build the distributable artefacts: compile the sm_100a kernels + host library in-tree, then sdist + wheel.
# (Reference scripts/release.sh uploads with twine; publishing is left to the caller: `twine upload dist/*`.)
set -e
cd "$(dirname "$0")/.."
python -m blades_b200.ops.build
rm -rf build dist blades_b200.egg-info
python setup.py sdist bdist_wheel
ls -l dist
