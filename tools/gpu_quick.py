"""Quick GPU check of the HDRx32 kernels: parity vs reference kernels + device time (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_probe import run
from fractalshark_b200 import RenderAlgorithm as A
W, H = 3840, 2160
run(5, W, H, A.GpuHDRx32PerturbedLAv2PO, 20000)
run(5, W, H, A.GpuHDRx32PerturbedLAv2)
run(1, W, H, A.GpuHDRx32PerturbedLAv2)
run(19, 1920, 1080, A.GpuHDRx32PerturbedLAv2, 2000000)
run(5, 1920, 1080, A.GpuHDRx32PerturbedLAv2, iter_bytes=8)
