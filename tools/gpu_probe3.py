"""Dev tool: parity + device time of the 2x32 / HDRx2x32 LAv2 variants vs the reference kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_probe import run
from fractalshark_b200 import RenderAlgorithm as A
W, H = 1920, 1080
run(100, W, H, A.Gpu2x32PerturbedLAv2)
run(100, W, H, A.Gpu2x32PerturbedLAv2PO)
run(100, W, H, A.Gpu2x32PerturbedLAv2LAO)
run(101, W, H, A.Gpu2x32PerturbedLAv2)
run(5, W, H, A.GpuHDRx2x32PerturbedLAv2)
run(5, 960, 540, A.GpuHDRx2x32PerturbedLAv2PO, 20000)
run(5, W, H, A.GpuHDRx2x32PerturbedLAv2LAO)
run(1, W, H, A.GpuHDRx2x32PerturbedLAv2)
run(100, W, H, A.GpuHDRx2x32PerturbedLAv2)
run(5, 960, 540, A.GpuHDRx2x32PerturbedLAv2, iter_bytes=8)
run(19, 960, 540, A.GpuHDRx2x32PerturbedLAv2, 3000000)
