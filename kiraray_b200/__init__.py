"""kiraray_b200 -- B200-native WavefrontPathTracer pass (drop-in for cuteday/KiRaRay's pass).

The product is native: `lib/libkrr_wfpt.so` (sm_100a CUDA kernels behind the C ABI of
include/krr_wfpt.h) and `lib/libkrr_host.so` (C++17 host layer mirroring the reference's
RenderPass / RenderApp / SceneImporter surface).  This Python package is only a ctypes binding used
by the tests and bench.py; it contains no rendering logic and NO CPU fallback -- if the native
libraries are missing it raises.
"""
from .binding import (  # noqa: F401
    HostApp, Wfpt, KrrSceneDesc, KrrLeafBsdfQuery, KrrLeafLightQuery, KrrCameraData, KrrStats, KrrColorSpaceData,
    load_host, load_wfpt, lib_dir, data_dir, color_space, NativeLibraryMissing,
    load_image, save_image, save_exr, error_metric, tonemap, accumulate_f64,
)
