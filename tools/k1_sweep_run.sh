#!/bin/bash
# On the GPU box: time every variant built by tools/k1_sweep.sh (K1 alone, interleaved rounds, bit-exactness checked).
cd "$(dirname "$0")/.."
python tools/k1_bench.py --libs tools/_variants/libvti_*.so "$@" > gpurun_out/k1_sweep.jsonl 2> gpurun_out/k1_sweep.err
cat gpurun_out/k1_sweep.jsonl; tail -3 gpurun_out/k1_sweep.err
