"""ctypes binding of ``include/taco2dec.h`` (the C-ABI drop-in boundary).

There is NO fallback: if the shared library is missing or the device is not sm_100 the
decoder raises.  Build with ``python -m tacotron2_subword_b200.build`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TACO2DEC_LIB") or os.path.join(_HERE, "csrc", "libtaco2dec.so")  # env override: A/B builds

ATTN_SMA, ATTN_LSA = 0, 1
PATH_AUTO, PATH_GENERIC, PATH_LATENCY, PATH_TENSOR, PATH_TENSOR_GRAPH = 0, 1, 2, 3, 4
W_FP32, W_FP16 = 0, 1
ABI_VERSION = 8

EXPORTED_SYMBOLS = (
    "taco2dec_abi_version", "taco2dec_last_error", "taco2dec_create", "taco2dec_destroy",
    "taco2dec_set_weights", "taco2dec_workspace_bytes", "taco2dec_forward_teacher_forced",
    "taco2dec_infer", "taco2dec_check", "taco2dec_launch_count", "taco2dec_philox_keep_mask",
    "taco2dec_launch_geometry", "taco2dec_set_profiling", "taco2dec_last_kernel_ms",
    "taco2dec_read_phase_clocks", "taco2dec_set_mode", "taco2dec_last_path",
    "taco2dec_test_gemm", "taco2dec_saved_layout_query", "taco2dec_grad_layout_query", "taco2dec_backward",
    "taco2dec_postnet_create", "taco2dec_postnet_destroy", "taco2dec_postnet_set_weights",
    "taco2dec_postnet_workspace_bytes", "taco2dec_postnet_forward",
    "taco2dec_set_batched_precision", "taco2dec_poll_abort", "taco2dec_read_debug_stamps", "taco2dec_measure_machine",
    "taco2dec_memprep_create", "taco2dec_memprep_destroy", "taco2dec_memprep_set_weights", "taco2dec_memprep_workspace_bytes",
    "taco2dec_memprep_forward", "taco2dec_memprep_project", "taco2dec_loss_workspace_bytes", "taco2dec_loss_forward",
    "taco2dec_postnet_set_precision", "taco2dec_wgrad_workspace_bytes", "taco2dec_wgrad_gemm",
    "taco2dec_postnet_rows_gemm_workspace_bytes", "taco2dec_postnet_rows_gemm", "taco2dec_postnet_bn_act_forward",
    "taco2dec_postnet_bn_act_backward", "taco2dec_postnet_bn_backward_input", "taco2dec_postnet_wgrad_workspace_bytes",
    "taco2dec_postnet_wgrad",
    "taco2dec_sgemm_nn", "taco2dec_bmm_tn",
)

_fp = C.c_void_p  # device pointers travel as integers


class Config(C.Structure):
    _fields_ = [("n_mel", C.c_int), ("enc_dim", C.c_int), ("attn_rnn_dim", C.c_int), ("dec_rnn_dim", C.c_int),
                ("prenet_dim", C.c_int), ("attn_dim", C.c_int), ("loc_filters", C.c_int), ("loc_kernel", C.c_int),
                ("attention", C.c_int), ("n_streams", C.c_int),
                ("p_attention_dropout", C.c_float), ("p_decoder_dropout", C.c_float)]


class StreamWeights(C.Structure):
    _fields_ = [(n, _fp) for n in ("prenet_w0", "prenet_w1", "arnn_w_ih", "arnn_w_hh", "arnn_b_ih", "arnn_b_hh",
                                   "query_w", "memory_w", "v", "loc_conv_w", "loc_dense_w")]


class Weights(C.Structure):
    _fields_ = [("stream", StreamWeights * 2)] + [(n, _fp) for n in (
        "drnn_w_ih", "drnn_w_hh", "drnn_b_ih", "drnn_b_hh", "proj_w", "proj_b", "gate_w", "gate_b")]


class Rng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("prenet_keep", (_fp * 2) * 2), ("lstm_keep", _fp), ("sma_noise", _fp * 2)]


class TFArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("T", C.c_int), ("T_in", C.c_int), ("T_sub", C.c_int),
                ("memory", _fp), ("embeddings", _fp), ("decoder_inputs", _fp),
                ("memory_lengths", _fp), ("bert_lengths", _fp), ("training", C.c_int), ("rng", Rng),
                ("mel", _fp), ("gate", _fp), ("align", _fp), ("align_bert", _fp),
                ("workspace", _fp), ("workspace_bytes", C.c_size_t),
                ("independent", C.c_int), ("saved", _fp), ("saved_bytes", C.c_size_t), ("processed_memory", _fp * 2)]


class SavedLayout(C.Structure):
    _fields_ = [("gates1", C.c_size_t), ("gates2", C.c_size_t), ("h1", C.c_size_t), ("ctx", C.c_size_t),
                ("h2", C.c_size_t), ("q", C.c_size_t), ("p", C.c_size_t * 2), ("pm", C.c_size_t * 2),
                ("pre", C.c_size_t * 2), ("pre0", C.c_size_t * 2), ("total", C.c_size_t)]


class GradLayout(C.Structure):
    _fields_ = [("dg1", C.c_size_t), ("dg2", C.c_size_t), ("dq", C.c_size_t), ("dctx", C.c_size_t),
                ("dpre", C.c_size_t), ("dv", C.c_size_t), ("dpm", C.c_size_t * 2), ("dloc_dense", C.c_size_t),
                ("dloc_conv", C.c_size_t), ("scratch", C.c_size_t),
                ("total", C.c_size_t)]


class BwdArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("T", C.c_int), ("T_in", C.c_int), ("T_sub", C.c_int),
                ("memory", _fp), ("embeddings", _fp), ("memory_lengths", _fp), ("bert_lengths", _fp),
                ("training", C.c_int), ("rng", Rng),
                ("align", _fp), ("align_bert", _fp), ("d_mel", _fp), ("d_gate", _fp),
                ("d_align", _fp), ("d_align_bert", _fp), ("independent", C.c_int),
                ("saved", _fp), ("saved_bytes", C.c_size_t), ("grads", _fp), ("grads_bytes", C.c_size_t)]


class PostnetLayer(C.Structure):
    _fields_ = [(n, _fp) for n in ("conv_w", "conv_b", "bn_weight", "bn_bias", "bn_mean", "bn_var")]


class PostnetWeights(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("bn_eps", C.c_float), ("layer", PostnetLayer * 8)]


class InferArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("T_in", C.c_int), ("T_sub", C.c_int), ("max_decoder_steps", C.c_int),
                ("gate_threshold", C.c_float), ("memory", _fp), ("embeddings", _fp),
                ("memory_lengths", _fp), ("bert_lengths", _fp), ("rng", Rng),
                ("mel", _fp), ("gate", _fp), ("align", _fp), ("align_bert", _fp),
                ("n_frames", _fp), ("reached_max", _fp), ("workspace", _fp), ("workspace_bytes", C.c_size_t),
                ("processed_memory", _fp * 2)]


class LossArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("n_mel", C.c_int), ("T", C.c_int), ("mel", _fp), ("mel_stride_b", C.c_int64),
                ("mel_stride_c", C.c_int64), ("mel_stride_t", C.c_int64), ("mel_postnet", _fp), ("gate", _fp),
                ("mel_target", _fp), ("gate_target", _fp), ("align", _fp * 2), ("align_target", _fp * 2), ("T_align", C.c_int * 2),
                ("d_mel", _fp), ("d_mel_postnet", _fp), ("d_gate", _fp), ("d_align", _fp * 2), ("losses", _fp),
                ("workspace", _fp), ("workspace_bytes", C.c_size_t)]


class Taco2DecError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """dlopen the in-tree library and declare prototypes.  Loud failure if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise Taco2DecError(
            f"{LIB_PATH} not found: the CUDA decoder has not been built and there is no CPU fallback. "
            "Run `python -m tacotron2_subword_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.taco2dec_abi_version.restype = C.c_int
    lib.taco2dec_abi_version.argtypes = []
    lib.taco2dec_last_error.restype = C.c_char_p
    lib.taco2dec_last_error.argtypes = []
    lib.taco2dec_create.restype = C.c_int
    lib.taco2dec_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(H)]
    lib.taco2dec_destroy.restype = C.c_int
    lib.taco2dec_destroy.argtypes = [H]
    lib.taco2dec_set_weights.restype = C.c_int
    lib.taco2dec_set_weights.argtypes = [H, C.POINTER(Weights), C.c_void_p]
    lib.taco2dec_workspace_bytes.restype = C.c_size_t
    lib.taco2dec_workspace_bytes.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.taco2dec_forward_teacher_forced.restype = C.c_int
    lib.taco2dec_forward_teacher_forced.argtypes = [H, C.POINTER(TFArgs), C.c_void_p]
    lib.taco2dec_infer.restype = C.c_int
    lib.taco2dec_infer.argtypes = [H, C.POINTER(InferArgs), C.c_void_p]
    lib.taco2dec_saved_layout_query.restype = C.c_int
    lib.taco2dec_saved_layout_query.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(SavedLayout)]
    lib.taco2dec_grad_layout_query.restype = C.c_int
    lib.taco2dec_grad_layout_query.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(GradLayout)]
    lib.taco2dec_backward.restype = C.c_int
    lib.taco2dec_backward.argtypes = [H, C.POINTER(BwdArgs), C.c_void_p]
    lib.taco2dec_postnet_create.restype = C.c_int
    lib.taco2dec_postnet_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(H)]
    lib.taco2dec_postnet_set_precision.restype = C.c_int
    lib.taco2dec_postnet_set_precision.argtypes = [H, C.c_int]
    lib.taco2dec_postnet_destroy.restype = C.c_int
    lib.taco2dec_postnet_destroy.argtypes = [H]
    lib.taco2dec_postnet_set_weights.restype = C.c_int
    lib.taco2dec_postnet_set_weights.argtypes = [H, C.POINTER(PostnetWeights), C.c_void_p]
    lib.taco2dec_postnet_workspace_bytes.restype = C.c_size_t
    lib.taco2dec_postnet_workspace_bytes.argtypes = [H, C.c_int, C.c_int]
    lib.taco2dec_postnet_forward.restype = C.c_int
    lib.taco2dec_postnet_forward.argtypes = [H, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.taco2dec_check.restype = C.c_int
    lib.taco2dec_check.argtypes = [H, C.c_void_p]
    lib.taco2dec_launch_count.restype = C.c_int64
    lib.taco2dec_launch_count.argtypes = [H]
    lib.taco2dec_philox_keep_mask.restype = C.c_int
    lib.taco2dec_philox_keep_mask.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    lib.taco2dec_launch_geometry.restype = C.c_int
    lib.taco2dec_launch_geometry.argtypes = [H, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.taco2dec_set_profiling.restype = C.c_int
    lib.taco2dec_set_profiling.argtypes = [H, C.c_int]
    lib.taco2dec_last_kernel_ms.restype = C.c_int
    lib.taco2dec_last_kernel_ms.argtypes = [H, C.POINTER(C.c_float)]
    lib.taco2dec_set_mode.restype = C.c_int
    lib.taco2dec_set_mode.argtypes = [H, C.c_int, C.c_int]
    lib.taco2dec_set_batched_precision.restype = C.c_int
    lib.taco2dec_set_batched_precision.argtypes = [H, C.c_int]
    lib.taco2dec_read_debug_stamps.restype = C.c_int
    lib.taco2dec_read_debug_stamps.argtypes = [H, C.c_void_p, C.POINTER(C.c_longlong)]
    lib.taco2dec_measure_machine.restype = C.c_int
    lib.taco2dec_measure_machine.argtypes = [H, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.taco2dec_memprep_create.restype = C.c_int
    lib.taco2dec_memprep_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(H)]
    lib.taco2dec_memprep_destroy.restype = C.c_int
    lib.taco2dec_memprep_destroy.argtypes = [H]
    lib.taco2dec_memprep_set_weights.restype = C.c_int
    lib.taco2dec_memprep_set_weights.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.taco2dec_memprep_workspace_bytes.restype = C.c_size_t
    lib.taco2dec_memprep_workspace_bytes.argtypes = [H, C.c_int]
    lib.taco2dec_memprep_forward.restype = C.c_int
    lib.taco2dec_memprep_forward.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.taco2dec_memprep_project.restype = C.c_int
    lib.taco2dec_memprep_project.argtypes = [H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.taco2dec_loss_workspace_bytes.restype = C.c_size_t
    lib.taco2dec_loss_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.taco2dec_loss_forward.restype = C.c_int
    lib.taco2dec_loss_forward.argtypes = [C.POINTER(LossArgs), C.c_void_p]
    lib.taco2dec_wgrad_workspace_bytes.restype = C.c_size_t
    lib.taco2dec_wgrad_workspace_bytes.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.taco2dec_wgrad_gemm.restype = C.c_int
    lib.taco2dec_wgrad_gemm.argtypes = [H, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.taco2dec_postnet_wgrad_workspace_bytes.restype = C.c_size_t
    lib.taco2dec_postnet_wgrad_workspace_bytes.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.taco2dec_postnet_wgrad.restype = C.c_int
    lib.taco2dec_postnet_wgrad.argtypes = lib.taco2dec_wgrad_gemm.argtypes
    lib.taco2dec_postnet_rows_gemm_workspace_bytes.restype = C.c_size_t
    lib.taco2dec_postnet_rows_gemm_workspace_bytes.argtypes = [H, C.c_int, C.c_int, C.c_int]
    lib.taco2dec_postnet_rows_gemm.restype = C.c_int
    lib.taco2dec_postnet_rows_gemm.argtypes = [H, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                               C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.taco2dec_postnet_bn_act_forward.restype = C.c_int
    lib.taco2dec_postnet_bn_act_forward.argtypes = [H, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_int, C.c_uint64, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                                    C.c_int64, C.c_void_p]
    lib.taco2dec_postnet_bn_act_backward.restype = C.c_int
    lib.taco2dec_postnet_bn_act_backward.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                     C.c_int, C.c_uint64, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.taco2dec_postnet_bn_backward_input.restype = C.c_int
    lib.taco2dec_postnet_bn_backward_input.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.taco2dec_sgemm_nn.restype = C.c_int
    lib.taco2dec_sgemm_nn.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    lib.taco2dec_bmm_tn.restype = C.c_int
    lib.taco2dec_bmm_tn.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.taco2dec_poll_abort.restype = C.c_int
    lib.taco2dec_poll_abort.argtypes = [H]
    lib.taco2dec_last_path.restype = C.c_int
    lib.taco2dec_last_path.argtypes = [H]
    lib.taco2dec_test_gemm.restype = C.c_int
    lib.taco2dec_test_gemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.taco2dec_read_phase_clocks.restype = C.c_int
    lib.taco2dec_read_phase_clocks.argtypes = [H, C.c_void_p, C.POINTER(C.c_longlong)]
    if lib.taco2dec_abi_version() != ABI_VERSION:
        raise Taco2DecError("libtaco2dec.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load_library().taco2dec_last_error()
        raise Taco2DecError(f"taco2dec error {rc}: {msg.decode() if msg else '?'}")
