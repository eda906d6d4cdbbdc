"""ctypes binding of libstb200.so (the C ABI declared in include/stb200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, this raises.
Nothing in this package imports the CPU checker that lives outside it.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# STB200_LIB: development knob to load an alternative build of the same library (A/B measurements)
LIB_PATH = os.environ.get("STB200_LIB") or os.path.join(_HERE, "lib", "libstb200.so")

_c_int, _c_uint, _c_void_p, _c_size_t = ctypes.c_int, ctypes.c_uint, ctypes.c_void_p, ctypes.c_size_t
P = _c_void_p

class IndexStruct(ctypes.Structure):
    """mirror of `stb200_index` (include/stb200.h)"""
    _fields_ = [("N", ctypes.c_int), ("M", ctypes.c_int), ("index0_offsets", ctypes.c_void_p), ("index1", ctypes.c_void_p),
                ("rel_idx", ctypes.c_void_p), ("t_offsets", ctypes.c_void_p), ("t_pair", ctypes.c_void_p),
                ("t_index0", ctypes.c_void_p), ("rel_packed", ctypes.c_void_p), ("t_rel_packed", ctypes.c_void_p),
                ("row_order", ctypes.c_void_p), ("len_order", ctypes.c_void_p), ("t_len_order", ctypes.c_void_p)]


_IX = ctypes.POINTER(IndexStruct)


class FusedPass(ctypes.Structure):
    """mirror of `stb200_fused_pass` (include/stb200.h)"""
    _fields_ = [("items", ctypes.c_void_p), ("n_items", ctypes.c_int), ("q_order", ctypes.c_void_p), ("k_order", ctypes.c_void_p),
                ("rel", ctypes.c_void_p), ("pos_win", ctypes.c_void_p), ("wstart", ctypes.c_void_p), ("tile_base", ctypes.c_void_p),
                ("bin_lo", ctypes.c_int), ("RB", ctypes.c_int), ("BQ", ctypes.c_int), ("BK", ctypes.c_int)]


_FP = ctypes.POINTER(FusedPass)

# name -> argtypes (every function returns int unless listed in _RESTYPES)
_SIGNATURES = {
    "stb200_transpose_csr": [_c_int, _c_int, P, P, P, P, P, P, _c_size_t, P],
    "stb200_length_order": [_c_int, P, P, P, P, _c_size_t, P],
    "stb200_attention_step1_forward_v2": [_c_int] * 4 + [_c_uint] + [P] * 6,
    "stb200_attention_step1_backward_v2": [_c_int] * 4 + [_c_uint] + [P] * 11,
    "stb200_dot_prod_with_idx_forward_v3": [_c_int] * 6 + [P] * 9,
    "stb200_dot_prod_with_idx_backward_v3": [_c_int] * 6 + [P] * 16,
    "stb200_attention_step2_with_rel_pos_value_forward_v2": [_c_int] * 6 + [P] * 8,
    "stb200_attention_step2_with_rel_pos_value_backward_v2": [_c_int] * 6 + [P] * 14,
    "stb200_segment_softmax_forward": [_c_int] * 3 + [P] * 5,
    "stb200_segment_softmax_backward": [_c_int] * 3 + [P] * 5,
    "stb200_attention_step1_forward": [_c_int] * 4 + [P] * 6,
    "stb200_attention_step1_backward": [_c_int] * 4 + [P] * 8,
    "stb200_attention_step2_forward": [_c_int] * 4 + [P] * 6,
    "stb200_attention_step2_backward": [_c_int] * 4 + [P] * 8,
    "stb200_dot_prod_with_idx_forward": [_c_int] * 5 + [P] * 6,
    "stb200_dot_prod_with_idx_backward": [_c_int] * 5 + [P] * 8,
    "stb200_attention_step2_with_rel_pos_value_forward": [_c_int] * 5 + [P] * 8,
    "stb200_attention_step2_with_rel_pos_value_backward": [_c_int] * 5 + [P] * 11,
    "stb200_furthestsampling": [_c_int, _c_int] + [P] * 6,
    "stb200_furthestsampling_ws": [_c_int, _c_int, _c_int, P, P, P, P, P, P, _c_size_t, P],
    "stb200_stratified_pairs_count": [_c_int, _c_int, P, P, ctypes.c_float, _c_int, P, _c_int, P, _c_size_t, P, P, P],
    "stb200_stratified_pairs_fill": [_c_int, P, ctypes.c_float, ctypes.c_float, _c_int, P, _c_size_t, P, P, P, P, P, P, _c_int, _c_int, P],
    "stb200_rel_pos_index_stratified": [_c_int, P, P, P, ctypes.c_float, ctypes.c_float, P, P],
    "stb200_pack_rel": [_c_int, _c_int, P, P, P, P],
    "stb200_window_logits_forward": [_IX, _c_int, _c_int, _c_int] + [P] * 6,
    "stb200_window_logits_backward": [_IX, _c_int, _c_int, _c_int] + [P] * 10,
    "stb200_window_logits_backward_ws": [_IX, _c_int, _c_int, _c_int] + [P] * 10 + [_c_size_t, P],
    "stb200_window_aggregate_forward": [_IX, _c_int, _c_int, _c_int] + [P] * 5,
    "stb200_window_aggregate_backward": [_IX, _c_int, _c_int, _c_int] + [P] * 8,
    "stb200_window_logits_forward_bf16": [_IX, _c_int, _c_int, _c_int] + [P] * 6,
    "stb200_window_aggregate_forward_bf16": [_IX, _c_int, _c_int, _c_int] + [P] * 5,
    "stb200_knnquery": [_c_int, _c_int, _c_int, P, P, P, P, P, P, P],
    "stb200_knnquery_ws": [_c_int, _c_int, _c_int, _c_int, P, P, P, P, P, P, P, _c_size_t, P],
    "stb200_classify_windows": [_c_int, P, P, P, P, P, P, P, P],
    "stb200_window_attention_forward_fused": [_IX, _c_int, P, P, _c_int, _c_int, _c_int] + [P] * 9,
    "stb200_segment_softmax_forward_rows": [_c_int, P, _c_int, P, P, P, P, P],
    "stb200_fused_plan_count": [_c_int, P, _c_size_t, _c_int, _c_int, _c_int, _c_int, _c_int, P, _c_size_t, P, P],
    "stb200_fused_plan_fill": [_c_int, P, ctypes.c_float, ctypes.c_float] + [_c_int] * 6 + [ctypes.c_float, ctypes.c_float, P, _c_size_t,
                               P, _c_size_t] + [P] * 10 + [_c_int, P, P],
    "stb200_fused_attention_forward": [_FP, _c_int, _c_int, _c_int, _c_int] + [P] * 10,
    "stb200_fused_attention_backward": [_FP, _c_int, _c_int, _c_int, _c_int] + [P] * 16,
    "stb200_tc_selftest": [_c_int] * 4 + [P] * 5,
    "stb200_qkv_split": [_c_int] * 3 + [P] * 6,
    "stb200_qkv_merge": [_c_int] * 3 + [P] * 6,
    "stb200_layer_norm_forward": [ctypes.c_longlong, _c_int, ctypes.c_float] + [P] * 7,
    "stb200_layer_norm_backward": [ctypes.c_longlong, _c_int] + [P] * 8,
    "stb200_kpconv_weighted": [_c_int] * 5 + [ctypes.c_float] + [P] * 7,
    "stb200_kpconv_weighted_backward": [_c_int] * 5 + [ctypes.c_float] + [P] * 7,
    "stb200_batch_from_offset": [_c_int, _c_int, P, P, P],
    "stb200_ball_query": [_c_int, _c_int, ctypes.c_float, _c_int, P, P, P, P, P, _c_size_t, P, P, P],
    "stb200_set_torch_semantics": [_c_int],
    "stb200_rel_pos_index_swin": [_c_int, P, P, P, ctypes.c_float, ctypes.c_float, ctypes.c_float, _c_int, P, P, P, P],
}
_RESTYPES = {
    "stb200_last_error": (ctypes.c_char_p, []),
    "stb200_launch_count": (ctypes.c_longlong, []),
    "stb200_version": (_c_int, []),
    "stb200_transpose_csr_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "stb200_length_order_workspace_bytes": (_c_size_t, [_c_int]),
    "stb200_window_logits_backward_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "stb200_pair_builder_workspace_bytes": (_c_size_t, [_c_int]),
    "stb200_fused_max_keys": (_c_int, []),
    "stb200_fused_plan_scratch_bytes": (_c_size_t, [_c_int]),
    "stb200_qkv_partial_rows": (_c_int, [_c_int, _c_int]),
    "stb200_layer_norm_partial_rows": (_c_int, [ctypes.c_longlong, _c_int]),
    "stb200_ball_query_workspace_bytes": (_c_size_t, [_c_int]),
    "stb200_knnquery_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "stb200_fps_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "stb200_profile_enable": (None, [_c_int]),
    "stb200_profile_dump": (_c_size_t, [ctypes.c_char_p, _c_size_t]),
}

EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + list(_RESTYPES))

_lib = None


class Stb200Error(RuntimeError):
    pass


def load():
    """Load libstb200.so; raise loudly when it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Stb200Error(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C stratified_transformer_b200/csrc`). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _c_int
    for name, (restype, argtypes) in _RESTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    if lib.stb200_version() < 103:   # the IndexStruct mirror below needs the 101 layout (len_order / t_len_order)
        raise Stb200Error(f"{LIB_PATH} is stale (ABI {lib.stb200_version()} < 103): rebuild it with `make -C stratified_transformer_b200/csrc`")
    _lib = lib
    return lib


def call(name: str, *args):
    """Call an int-returning entry point and turn a non-zero status into an exception."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.stb200_last_error().decode(errors="replace")
        raise Stb200Error(f"{name} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().stb200_launch_count())


def profile_enable(on: bool = True) -> None:
    load().stb200_profile_enable(1 if on else 0)


def profile_dump() -> dict:
    """Per-kernel totals since the last dump: {name: {"launches", "ms", "bytes"}} (synchronises)."""
    import json
    lib = load()
    n = lib.stb200_profile_dump(None, 0)
    buf = ctypes.create_string_buffer(n + 16)
    lib.stb200_profile_dump(buf, n + 16)
    return json.loads(buf.value.decode())
