#!/bin/bash
# one GPU call: overlay/JPEG tests, the three nvJPEG backends through bench.py --jpeg, cfg4 bench + ncu capture of K3
python -m pytest tests/test_overlay.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25
for be in hardware gpu hybrid; do
  VTI_JPEG_BACKEND=$be python bench.py --steps 5 --warmup 3 --jpeg --no-cpu 2>gpurun_out/jpeg_$be.err > gpurun_out/jpeg_$be.json
  python -c "import sys,json; d=json.loads(open('gpurun_out/jpeg_$be.json').read().strip().splitlines()[-1]); print(json.dumps(d.get('ingest_jpeg')))"
done
python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu > gpurun_out/cfg4_pre_ncu.json 2>gpurun_out/cfg4_pre_ncu.err && \
ncu --set full --clock-control none --import-source on -k regex:k3_nms -c 1 -o gpurun_out/k3_r2_cfg4_rounds -f \
  python bench.py --config cfg4 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_k3cfg4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
