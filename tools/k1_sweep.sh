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
build A_lwt_nostage2_nodivfma -DVTI_K1_LWT=1 -DVTI_K1_STAGE2=0 -DVTI_K1_DIVFMA=0
build B_lwt_nostage2 -DVTI_K1_LWT=1 -DVTI_K1_STAGE2=0
build C_lwt -DVTI_K1_LWT=1
build D_A_t128 -DVTI_K1_LWT=1 -DVTI_K1_STAGE2=0 -DVTI_K1_DIVFMA=0 -DVTI_FT_THREADS=128 -DVTI_FT_MINB=8
build E_A_t128_ty16 -DVTI_K1_LWT=1 -DVTI_K1_STAGE2=0 -DVTI_K1_DIVFMA=0 -DVTI_FT_THREADS=128 -DVTI_FTY=16 -DVTI_FT_MINB=8
build F_B_minb5_lut8 -DVTI_K1_LWT=1 -DVTI_K1_STAGE2=0 -DVTI_K1_LUT8=1
build r1nf -DVTI_K1_STAGE2=0 -DVTI_K1_DIVFMA=0 -DVTI_K1_PADROWS=0
python vision_textile_inspection_b200/build.py --force > /dev/null   # leave the default objects in build/
