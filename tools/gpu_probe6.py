"""Dev tool: parity + device time of the RC (compressed-orbit) LAv2 variants vs the reference kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_probe import run
from fractalshark_b200 import RenderAlgorithm as A
W, H = 960, 540
run(5, W, H, A.GpuHDRx32PerturbedRCLAv2)
run(5, W, H, A.GpuHDRx32PerturbedRCLAv2PO, 20000)
run(5, W, H, A.GpuHDRx32PerturbedRCLAv2LAO)
run(1, W, H, A.GpuHDRx32PerturbedRCLAv2, iter_bytes=8)
run(19, 480, 270, A.GpuHDRx32PerturbedRCLAv2, 3000000)
run(100, W, H, A.Gpu1x64PerturbedRCLAv2)
run(100, W, H, A.Gpu1x64PerturbedRCLAv2PO)
run(101, W, H, A.Gpu1x32PerturbedRCLAv2)
run(101, W, H, A.Gpu1x32PerturbedRCLAv2PO)
run(5, W, H, A.GpuHDRx64PerturbedRCLAv2)
run(100, W, H, A.Gpu2x32PerturbedRCLAv2)
run(5, W, H, A.GpuHDRx2x32PerturbedRCLAv2)
run(5, 480, 270, A.GpuHDRx2x32PerturbedRCLAv2PO, 20000)
