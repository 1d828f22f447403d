#!/bin/bash
# Builds side-by-side libvti variants (one -D switch each) for the K1 A/B runs; run tools/k1_sweep_run.sh on the GPU box.
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/variants
build() {  # name flags...
  name=$1; shift
  VTI_LIB=$PWD/tools/_variants/libvti_$name.so VTI_NVCC_FLAGS="$*" python vision_textile_inspection_b200/build.py --force > /dev/null
  echo built $name
}
rm -rf tools/_variants; mkdir -p tools/_variants
build A_tmaout
build B_stg -DVTI_K1_TMAOUT=0
python vision_textile_inspection_b200/build.py --force > /dev/null   # leave the default objects in build/
