"""pathtrace_rs_b200 — B200-native implementation of pathtrace-rs's per-pixel path-tracing loop.

The product is two shared libraries built from this directory:

* ``lib/libptgpu.so``  — hand-written CUDA for sm_100a behind the C ABI of ``include/ptgpu.h``
  (drop-in for ``Scene::update``, src/scene.rs:73-121).
* ``lib/libpthost.so`` — C++ mirror of the reference's host API (Params / Camera / Material / Texture /
  Storage / presets / offline), which flattens a scene and calls the C ABI.

This Python package is only the harness around them: ctypes bindings for tests and bench.py, and the
torch plumbing (device buffers, streams, torch.distributed) for multi-GPU runs.  There is no CPU
rendering path here and nothing in this package imports ``oracle/``.
"""
from .ffi import (  # noqa: F401
    PtCamera, PtParams, PtPartition, PtRenderStats, PtDeviceInfo, PtError, PtOptions,
    libptgpu, libpthost, abi_symbols,
)
from .scene import Params, Preset, device_info, image_open, probe_fp32_peak, render_offline, write_ppm  # noqa: F401

__all__ = [
    "Params", "Preset", "device_info", "image_open", "probe_fp32_peak", "render_offline", "write_ppm",
    "PtCamera", "PtParams", "PtPartition", "PtRenderStats", "PtDeviceInfo", "PtError", "PtOptions",
    "libptgpu", "libpthost", "abi_symbols",
]
