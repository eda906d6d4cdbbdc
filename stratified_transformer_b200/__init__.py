"""B200-native window-attention hot path of Stratified Transformer.

    pointops2_cuda   extension-level mirror of the reference's pybind module (ctypes over libstb200.so)
    pointops         autograd operators with the reference's names / argument orders (+ fused and bf16 entry points)
    index            pair-index construction on the device (PairIndex, LayerIndex, GeometryPrefetcher)
    window_attention WindowAttention / SwinWindowAttention module mirrors
    parallel         scene sharding + gradient all-reduce helpers
    synthetic        S3DIS / ScanNet-shape scene generator

The CUDA library is loaded lazily on first use and there is no CPU fallback (`_cabi.load()` raises if it is missing).
"""
__version__ = "0.1.0"
__all__ = ["pointops2_cuda", "pointops", "index", "window_attention", "parallel", "synthetic"]
