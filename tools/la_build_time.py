"""Dev probe: time of the LA table construction (in-tree builder) for a view, per phase (FS_LA_TIMING=1 prints them)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cases
from fractalshark_b200 import RenderAlgorithm as A
from fractalshark_b200.host_inputs import LaTable
view = int(sys.argv[1]) if len(sys.argv) > 1 else 14
_, coords, orbit, la, n = cases.make_inputs(view, 384, 216, A.GpuHDRx32PerturbedLAv2, None, 4)
ts = []
for rep in range(7):
    t0 = time.perf_counter(); l = LaTable(orbit, 4); ts.append((time.perf_counter() - t0) * 1e3)
print(f"view {view}: LA build min {min(ts):.3f} ms median {sorted(ts)[len(ts)//2]:.3f} ms  ({l.num_las} records, {os.cpu_count()} cpus, FS_HOST_THREADS={os.environ.get('FS_HOST_THREADS','default')})", flush=True)
from fractalshark_b200.host_inputs import BlaTable
_, _, borbit, btab, _ = cases.make_inputs(view, 384, 216, A.GpuHDRx32PerturbedBLA, None, 4)
ts = []
for rep in range(5):
    t0 = time.perf_counter(); b = BlaTable(borbit); ts.append((time.perf_counter() - t0) * 1e3)
print(f"view {view}: BLA build min {min(ts):.3f} ms  ({b.num_levels} levels)", flush=True)
