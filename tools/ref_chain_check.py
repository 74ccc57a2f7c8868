"""time the reference's unfused OpenCL launch sequence on the GPU and compare its frame with ours"""
import os, sys, asyncio
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
print(bench.reference_kernels_on_gpu("noise"))
