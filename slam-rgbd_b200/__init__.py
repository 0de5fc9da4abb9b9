"""slam-rgbd_b200 -- B200 (sm_100a) frame-to-frame depth tracker behind SLAM-RGBD's
AlgorithmModule.  The product is native: ``lib/libyouth_cuda.so`` (CUDA kernels + C ABI,
``include/youth_cuda.h``) and ``lib/libAlgorithmModule.so`` (C facade + host helpers).
This Python package is only the ctypes view of those libraries used by tests, bench.py
and the multi-GPU launcher; it contains no arithmetic of its own and no CPU fallback.
"""
from .binding import (  # noqa: F401
    Codec,
    CudaLibraryMissing,
    SynthConfig,
    Tracker,
    TsdfConfig,
    YouthConfig,
    default_config,
    host_lib,
    cuda_lib,
    lib_paths,
    synth_gt,
    synth_lib,
    synth_sequence,
    tsdf_config,
)
from .build import build_all  # noqa: F401
