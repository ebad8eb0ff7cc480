"""fractalshark_b200 -- B200-native per-pixel Mandelbrot render path (drop-in for FractalShark's GPURenderer).

The product is ``libfsgpu.so`` (hand-written sm_100a CUDA behind the C-ABI in ``include/fs_gpu.h``).
This package is the thin host-side mirror of the reference interface used by tests and ``bench.py``:

* :mod:`fractalshark_b200.gpu_renderer` -- ``GPURenderer`` with the reference's method names
  (FractalSharkLib/GPU_Render.h:20-227) over the C-ABI.
* :mod:`fractalshark_b200.algorithms`   -- the ``RenderAlgorithm`` enum and its traits table
  (FractalSharkLib/RenderAlgorithm.h:81-159).
* :mod:`fractalshark_b200.host_inputs`  -- view coordinates, reference orbit and LA/BLA tables
  (``libfshost.so``; the inputs the reference's host code would hand over).
* :mod:`fractalshark_b200.views`        -- view presets used by BASELINE.json's configs.

There is no CPU fallback: importing the renderer without the built CUDA library raises.
"""
from .algorithms import RenderAlgorithm, Numeric, LAv2Mode, PerturbExtras, traits  # noqa: F401

__all__ = ["RenderAlgorithm", "Numeric", "LAv2Mode", "PerturbExtras", "traits"]
