"""phaneron_b200 -- B200-native (sm_100a CUDA) replacement for the GPU hot path of
Streampunk/phaneron: v210 unpack -> linear RGB -> transform -> transition -> N-layer
combine -> v210 pack, fused into one launch per output frame, behind phaneron's own
src/process operator surface and clJobQueue API (see DESIGN.md / INTEGRATION.md)."""
from ._lib import PhaneronError, LIB_PATH  # noqa: F401
from .nodencl import clContext, OpenCLBuffer, OpenCLProgram, KernelSpec, RunTimings  # noqa: F401
from .cl_job_queue import ClJobs, ClProcessJobs, JobID  # noqa: F401

__version__ = "0.1.0"
