// Dev TU: only the HDRx32 / u32 / Full LAv2 kernel, for quick SASS inspection:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false --expt-relaxed-constexpr -Xptxas -v \
//        -cubin -o /tmp/lav2.cubin tools/dev_lav2_only.cu && cuobjdump -sass /tmp/lav2.cubin
#include "../fractalshark_b200/csrc/fs_lav2.cuh"
using namespace fs;
template __global__ void fs::lav2_kernel<NumHdr<float>, uint32_t, Lav2Mode::Full, false, AtPhase::Fused>(const Lav2Args<NumHdr<float>, uint32_t>);
