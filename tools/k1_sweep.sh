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
build A_early3
build B_early1 -DVTI_K1_EARLY=1
build C_early2 -DVTI_K1_EARLY=2
build D_early0 -DVTI_K1_EARLY=0
build E_early3_reg56 -DVTI_K1_MAXREG=56
build F_early3_reg64 -DVTI_K1_MAXREG=64
build G_early0_reg56 -DVTI_K1_EARLY=0 -DVTI_K1_MAXREG=56
python vision_textile_inspection_b200/build.py --force > /dev/null   # leave the default objects in build/
