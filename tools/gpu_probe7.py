"""Dev tool: View 14 (north_star target) parity + device time vs the reference kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_probe import run
from fractalshark_b200 import RenderAlgorithm as A
run(14, 3840, 2160, A.GpuHDRx32PerturbedLAv2)
run(14, 1920, 1080, A.GpuHDRx2x32PerturbedLAv2)
run(14, 1920, 1080, A.GpuHDRx32PerturbedRCLAv2)
run(14, 960, 540, A.GpuHDRx32PerturbedLAv2PO, 100000)
run(14, 960, 540, A.GpuHDRx32PerturbedBLA)
run(19, 3840, 2160, A.GpuHDRx32PerturbedLAv2)
