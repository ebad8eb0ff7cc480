"""Dev tool: parity + device time of the extended-precision direct kernels vs the reference kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_probe import run
from fractalshark_b200 import RenderAlgorithm as A
run(0, 1920, 1080, A.Gpu2x32, 4096)
run(100, 1920, 1080, A.Gpu2x32, 20000)
run(0, 960, 540, A.Gpu2x32, 1000, iter_bytes=8)
run(0, 1920, 1080, A.Gpu2x64, 2048)
run(102, 960, 540, A.Gpu2x64, 20000)
run(0, 960, 540, A.GpuHDRx32, 1024)
run(100, 960, 540, A.GpuHDRx32, 5000)
run(0, 960, 540, A.Gpu4x32, 1024)
run(104, 480, 270, A.Gpu4x32, 20000)
run(0, 960, 540, A.Gpu4x64, 1024)
run(103, 480, 270, A.Gpu4x64, 20000)
run(102, 480, 270, A.Gpu4x64, 20000, iter_bytes=8)
