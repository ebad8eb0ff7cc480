#!/bin/bash
# Dev: time every library build under fractalshark_b200/variants/ on the same frames (tools/quick_time.py)
for f in fractalshark_b200/variants/libfsgpu_*.so; do
  FS_GPU_LIB=$PWD/$f timeout 300 python tools/quick_time.py $(basename $f .so) 2>&1 | tail -1
done
