"""Dev tool: parity + device time of the BLA kernels and the non-HDRx32 LAv2 variants vs the reference kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_probe import run
from fractalshark_b200 import RenderAlgorithm as A
W, H = 1920, 1080
run(5, W, H, A.GpuHDRx32PerturbedBLA)
run(1, W, H, A.GpuHDRx32PerturbedBLA)
run(5, W, H, A.GpuHDRx64PerturbedBLA)
run(100, W, H, A.Gpu1x64PerturbedBLA)
run(1, W, H, A.Gpu1x64PerturbedBLA)
run(5, W, H, A.GpuHDRx32PerturbedBLA, iter_bytes=8)
run(5, W, H, A.GpuHDRx64PerturbedLAv2)
run(5, W, H, A.GpuHDRx64PerturbedLAv2PO, 20000)
run(100, W, H, A.Gpu1x64PerturbedLAv2)
run(100, W, H, A.Gpu1x64PerturbedLAv2PO)
run(1, W, H, A.Gpu1x64PerturbedLAv2)
run(101, W, H, A.Gpu1x32PerturbedLAv2)
run(101, W, H, A.Gpu1x32PerturbedLAv2PO)
run(100, W, H, A.GpuHDRx32PerturbedLAv2)
run(5, 3840, 2160, A.GpuHDRx32PerturbedBLA)
