#!/bin/bash
# Run the kernel numerics tests under compute-sanitizer (needs a GPU; slow).
TOOL=${1:-memcheck}
compute-sanitizer --tool "$TOOL" --error-exitcode 1 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x \
    -k "trimmed or row_combine or fill_normal or attack_rows or im2col or client_bn"
