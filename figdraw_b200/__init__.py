"""figdraw_b200 -- B200-native (sm_100a) backend for figdraw's render hot path.

The product is `csrc/libfigdraw_cuda.so` (C ABI in include/figdraw_cuda.h); this package is the host-side mirror of
the reference's interface that the tests and the benchmark drive:

  abi            ctypes signatures, record/POD dtypes; `load_library()` raises when the extension is not built
  cuda_context   `CudaContext` -- the `BackendContext` subclass backed by the CUDA library (no CPU fallback)
  figbackend     `BackendContext` method table, `TraceBackend` (records calls as `fdc_call`)
  fignodes       scene model (`Fig`, `RenderList`, `Renders`), figrender: the front-end restatement
  native_scene   POD marshalling + `fdc_flatten_renders` / `fdc_render_frame` (front-end run natively)
  scenes*        golden scenes, BASELINE configs, fuzzed call streams
  bands          tile-row band layout and the NCCL/gloo all-gather used for N > 1
"""
__version__ = "0.1.0"
