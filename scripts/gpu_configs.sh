#!/bin/bash
# Functional + timing check of the other BASELINE configs on one GPU.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "whole_round" 2>&1 | tail -3 | cut -c1-200
timeout 600 python bench.py --config multikrum --steps 5 --warmup 3 --no-e2e > gpurun_out/cfg_multikrum.json 2> gpurun_out/cfg_multikrum.err; tail -1 gpurun_out/cfg_multikrum.json | cut -c1-200; tail -3 gpurun_out/cfg_multikrum.err | cut -c1-300
timeout 900 python bench.py --config fedavg_median --steps 2 --warmup 3 --no-e2e > gpurun_out/cfg_fedavg.json 2> gpurun_out/cfg_fedavg.err; tail -1 gpurun_out/cfg_fedavg.json | cut -c1-200; tail -3 gpurun_out/cfg_fedavg.err | cut -c1-300
timeout 900 python bench.py --config geomed_r50 --steps 3 --warmup 3 --no-e2e > gpurun_out/cfg_geomed_r50.json 2> gpurun_out/cfg_geomed_r50.err; tail -1 gpurun_out/cfg_geomed_r50.json | cut -c1-200; tail -3 gpurun_out/cfg_geomed_r50.err | cut -c1-300
