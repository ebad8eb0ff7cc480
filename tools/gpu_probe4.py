"""Dev tool: parity + device time of the scaled kernels vs the reference kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_probe import run
from fractalshark_b200 import RenderAlgorithm as A
W, H = 1920, 1080
run(100, W, H, A.Gpu1x32PerturbedScaled)
run(101, W, H, A.Gpu1x32PerturbedScaled)
run(1, W, H, A.Gpu1x32PerturbedScaled)
run(100, 960, 540, A.Gpu1x32PerturbedScaled, iter_bytes=8)
run(100, W, H, A.GpuHDRx32PerturbedScaled)
run(1, W, H, A.GpuHDRx32PerturbedScaled)
run(5, W, H, A.GpuHDRx32PerturbedScaled, 100000)
run(1, 960, 540, A.GpuHDRx32PerturbedScaled, iter_bytes=8)
run(19, 960, 540, A.GpuHDRx32PerturbedScaled, 3000000)
run(19, 960, 540, A.Gpu1x32PerturbedScaled, 3000000)
