#!/bin/bash
# One-time offline install of the UNMODIFIED reference into baseline/_ref (git-ignored, travels with gpurun).
#  1. pip (the documented command) -- resolves no packages: the sdist's find_packages() finds nothing because
#     src/blades/ has no __init__.py, so only blades-0.0.14.dist-info is produced;
#  2. therefore the package tree itself is placed next to the dist-info, byte for byte (python imports it as a
#     namespace package, which is also how the reference's own scripts use it: sys.path.insert(0, '../src')).
set -e
cd "$(dirname "$0")/.."
rm -rf /tmp/_blades_ref_src && cp -r /root/reference /tmp/_blades_ref_src
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target baseline/_ref --upgrade /tmp/_blades_ref_src/src
rm -rf baseline/_ref/blades && cp -r /root/reference/src/blades baseline/_ref/blades
find baseline/_ref -name __pycache__ -prune -exec rm -rf {} +
diff -r /root/reference/src/blades baseline/_ref/blades && echo "reference tree installed unmodified"
