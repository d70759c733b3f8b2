"""ctypes binding of libkaldi_fp16.so -- the same C ABI a cgo caller links (include/*.h).

There is no fallback: if the shared library is missing or a symbol cannot be resolved the import
fails loudly (build it with `python -m kaldi_fp16_b200.build` or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libkaldi_fp16.so"

c_void_p, c_int, c_float, c_size_t = C.c_void_p, C.c_int, C.c_float, C.c_size_t
c_u32, c_i64, c_u64 = C.c_uint32, C.c_int64, C.c_uint64
FP = C.POINTER(C.c_float)


class KMat(C.Structure):
    _fields_ = [("ptr", c_void_p), ("rows", c_int), ("cols", c_int), ("ld", c_int), ("halo", c_int)]


class ConvAddr(C.Structure):
    """struct kfp16_conv_addr (include/kaldi_fp16_fused.h)."""

    _fields_ = [("mode", c_int), ("x", c_void_p), ("T", c_int), ("H", c_int), ("P", c_int), ("C", c_int),
                ("rows_h", c_int), ("ntaps", c_int), ("dt", c_int * 16), ("hq", c_int * 16), ("par", c_int * 16),
                ("brow", c_int * 16)]


class GemmDesc(C.Structure):
    """struct kfp16_gemm_desc (include/kaldi_fp16_fused.h)."""

    _fields_ = [
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("a_major", c_int), ("b_major", c_int),
        ("A", KMat), ("B", KMat),
        ("groups", c_int), ("kslabs", c_int), ("kslab_len", c_int),
        ("a_row_off", (c_int * 2) * 2), ("a_col_off", (c_int * 2) * 2),
        ("b_row_off", (c_int * 2) * 2), ("b_col_off", (c_int * 2) * 2),
        ("D", c_void_p * 2), ("ldd", c_int),
        ("R", c_void_p * 2), ("ldr", c_int),
        ("flags", c_u32),
        ("alpha", c_float), ("beta", c_float), ("res_scale", c_float),
        ("bias", c_void_p), ("bn_scale", c_void_p), ("bn_shift", c_void_p),
        ("vec_gstride", c_int),
        ("mask_out", c_void_p), ("mask_in", c_void_p), ("mask_ld", c_int),
        ("split_k", c_int), ("ws", c_void_p * 2), ("ws_ld", c_int), ("ws_transposed", c_int),
        ("drop_p", c_float), ("drop_seed", c_u32),
        ("force_bn", c_int), ("force_generic", c_int), ("force_cg", c_int), ("no_share", c_int), ("debug_clock_buf", c_void_p),
        ("A2", KMat), ("B2", KMat), ("a2_row_off", c_int * 2), ("b2_row_off", c_int * 2),
        ("ws2", c_void_p * 2), ("ws2_ld", c_int), ("ws2_transposed", c_int),
        ("conv", ConvAddr),
        ("drop_seed_dev", c_void_p),
        ("zero_row_period", c_int), ("zero_row_lo", c_int), ("zero_row_hi", c_int),
    ]


class WgradProb(C.Structure):
    """struct kfp16_wgrad_prob (include/kaldi_fp16_fused.h)."""

    _fields_ = [("A", KMat), ("B", KMat), ("a_row_off", c_int * 2), ("b_row_off", c_int * 2),
                ("ws", c_void_p * 2), ("ws_ld", c_int), ("ws_transposed", c_int)]


class ChainFst(C.Structure):
    """struct kfp16_chain_fst (include/kaldi_fp16_chain.h): host CSR"""

    _fields_ = [("row_ptr", c_void_p), ("col_idx", c_void_p), ("labels", c_void_p), ("weights", c_void_p),
                ("final_states", c_void_p), ("final_weights", c_void_p),
                ("num_states", c_int), ("num_arcs", c_int), ("num_final", c_int), ("start_state", c_int)]


class CmDesc(C.Structure):
    """struct kfp16_cm_desc (include/kaldi_fp16_fused.h)"""

    _fields_ = [("format", c_int), ("rows", c_int), ("cols", c_int), ("global_min", c_float), ("global_range", c_float),
                ("payload_offset", c_size_t), ("dst_row", c_int)]


class NetOpts(C.Structure):
    """struct kfp16_net_opts (include/kaldi_fp16_nnet.h)."""

    _fields_ = [("n_seq", c_int), ("seq_len", c_int), ("ref_round", c_int), ("train", c_int),
                ("lr", c_float), ("momentum", c_float), ("conv_cartesian", c_int),
                ("grad_scale", c_float), ("round_grad", c_int)]


class GPUBatchPtrs(C.Structure):
    """struct GPUBatchPtrs (include/kaldi_fp16_bridge.h; reference cpp/include/bridge.h:33-50)."""

    _fields_ = [
        ("d_features", c_void_p), ("d_ivectors", c_void_p), ("d_csr_row_ptr", c_void_p),
        ("d_csr_col_idx", c_void_p), ("d_csr_labels", c_void_p), ("d_csr_weights", c_void_p),
        ("d_buffer", c_void_p),
        ("total_bytes", c_size_t), ("features_bytes", c_size_t), ("ivectors_bytes", c_size_t),
        ("csr_rowptr_bytes", c_size_t), ("csr_colidx_bytes", c_size_t),
        ("csr_labels_bytes", c_size_t), ("csr_weights_bytes", c_size_t),
    ]


EPI_BIAS, EPI_RELU, EPI_BN, EPI_RESID, EPI_BETA = 1, 2, 4, 8, 16
EPI_REF_ROUND, EPI_MASK, EPI_SPLITK, EPI_DROPOUT, EPI_GRADMASK = 32, 64, 128, 256, 512
K_MAJOR, MN_MAJOR = 0, 1

# name -> (restype, argtypes).  Every symbol declared in include/*.h is listed here; the
# "not gpu" test-suite checks the list against the headers and against the built library.
SIGNATURES: dict[str, tuple] = {
    # ---- kaldi_fp16_fused.h
    "kfp16_ctx_create": (c_void_p, [c_int]),
    "kfp16_ctx_destroy": (None, [c_void_p]),
    "kfp16_ctx_set_stream": (c_int, [c_void_p, c_void_p]),
    "kfp16_ctx_get_stream": (c_void_p, [c_void_p]),
    "kfp16_ctx_num_sms": (c_int, [c_void_p]),
    "kfp16_ctx_set_max_ctas": (c_int, [c_void_p, c_int]),
    "kfp16_ctx_set_profile": (c_int, [c_void_p, c_int]),
    "kfp16_ctx_profile_read": (c_int, [c_void_p, C.POINTER(c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "kfp16_set_default_stream": (None, [c_void_p]),
    "kfp16_launch_count": (c_u64, []),
    "kfp16_gemm_kind_launches": (c_u64, [c_int]),
    "kfp16_dropout_uniform": (c_float, [c_u32, c_u32, c_u32]),
    "kfp16_scale_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float]),
    "kfp16_bump_counter": (c_int, [c_void_p, c_void_p]),
    "kfp16_last_error": (C.c_char_p, []),
    "kfp16_gemm_ex": (c_int, [c_void_p, C.POINTER(GemmDesc)]),
    "kfp16_gemm": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_int, c_void_p, c_int, c_float, c_void_p]),
    "kfp16_bn_fold": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_void_p, c_void_p]),
    "kfp16_bn_relu_backward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int]),
    "kfp16_add_bias": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int]),
    "kfp16_colsum": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "kfp16_f32_to_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "kfp16_pad_edges": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "kfp16_fold_edges": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "kfp16_sgd_update_flat": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_float, c_float, c_size_t]),
    "kfp16_pack_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int]),
    "kfp16_pack_rows_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int]),
    "kfp16_sgd_update_flat_hp": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t]),
    "kfp16_scale_f32_to_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kfp16_unpack_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int]),
    "kfp16_bcast_rows": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int]),
    "kfp16_seq_sum": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int]),
    "kfp16_zero_halo": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "kfp16_scale_shift": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "kfp16_bn_batch_stats": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int]),
    "kfp16_bn_finalize": (c_int, [c_void_p, c_void_p, C.c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "kfp16_bn_apply": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_int, c_int, c_int, c_int, c_int]),
    "kfp16_attention_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p] + [c_int] * 9 + [c_float]),
    "kfp16_attention_backward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p] + [c_int] * 9 + [c_float]),
    "kfp16_spec_augment": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 9 + [c_u32, c_void_p]),
    "kfp16_scale_shift_ld": (c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_void_p, c_void_p]),
    "kfp16_zero_rows_except": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "kfp16_half_sq_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "kfp16_bn_relu_backward_bias": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "kfp16_bn_relu_backward_bias_fold": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "kfp16_colsum_accum": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "kfp16_cm_payload_bytes": (c_size_t, [C.POINTER(CmDesc)]),
    "kfp16_decode_matrices": (c_int, [c_void_p, c_void_p, c_size_t, C.POINTER(CmDesc), c_int, c_void_p, c_int, c_int]),
    "kfp16_net_set_input_compressed": (c_int, [c_void_p, C.c_char_p, c_void_p, c_size_t, C.POINTER(CmDesc), c_int]),
    "kfp16_im2col": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 9 + [C.POINTER(c_int), C.POINTER(c_int)]),
    "kfp16_col2im": (c_int, [c_void_p, c_void_p, c_int, c_void_p] + [c_int] * 8 + [C.POINTER(c_int), C.POINTER(c_int)]),
    # ---- kaldi_fp16_nnet.h
    "kfp16_net_create": (c_void_p, [c_void_p, C.c_char_p, C.POINTER(NetOpts)]),
    "kfp16_net_destroy": (None, [c_void_p]),
    "kfp16_net_num_layers": (c_int, [c_void_p]),
    "kfp16_net_layer_name": (C.c_char_p, [c_void_p, c_int]),
    "kfp16_net_layer_type": (C.c_char_p, [c_void_p, c_int]),
    "kfp16_net_layer_dim": (c_int, [c_void_p, c_int]),
    "kfp16_net_padded_rows": (c_int, [c_void_p]),
    "kfp16_net_halo": (c_int, [c_void_p]),
    "kfp16_net_flops_forward": (C.c_double, [c_void_p]),
    "kfp16_net_flops_backward": (C.c_double, [c_void_p]),
    "kfp16_net_flops_skipped": (C.c_double, [c_void_p]),
    "kfp16_net_num_params": (c_int, [c_void_p]),
    "kfp16_net_param_name": (C.c_char_p, [c_void_p, c_int]),
    "kfp16_net_param_shape": (c_int, [c_void_p, c_int, C.POINTER(c_int), C.POINTER(c_int)]),
    "kfp16_net_param_offset": (c_size_t, [c_void_p, c_int]),
    "kfp16_net_bucket_size": (c_size_t, [c_void_p]),
    "kfp16_net_params_f16": (c_void_p, [c_void_p]),
    "kfp16_net_params_f32": (c_void_p, [c_void_p]),
    "kfp16_net_velocity": (c_void_p, [c_void_p]),
    "kfp16_net_grads_f32": (c_void_p, [c_void_p]),
    "kfp16_net_set_param": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, c_int]),
    "kfp16_net_get_param": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, c_int]),
    "kfp16_net_set_bn": (c_int, [c_void_p, C.c_char_p, C.c_char_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int]),
    "kfp16_net_init_random": (c_int, [c_void_p, c_u64]),
    "kfp16_net_set_input": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, c_int]),
    "kfp16_net_set_input_device": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, c_int]),
    "kfp16_net_prefetch_input": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, c_int]),
    "kfp16_net_commit_input": (c_int, [c_void_p, C.c_char_p]),
    "kfp16_net_set_input_f32": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, c_int]),
    "kfp16_net_prefetch_input_f32": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, c_int]),
    "kfp16_net_forward": (c_int, [c_void_p]),
    "kfp16_net_get_output": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, c_int]),
    "kfp16_net_get_mask": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, c_int]),
    "kfp16_net_zero_grads": (c_int, [c_void_p]),
    "kfp16_net_loss_half_sq": (c_int, [c_void_p, C.c_char_p]),
    "kfp16_net_set_output_grad": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, c_int]),
    "kfp16_net_backward": (c_int, [c_void_p]),
    "kfp16_net_get_grad": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, c_int]),
    "kfp16_net_sgd_step": (c_int, [c_void_p, c_float, c_int]),
    "kfp16_net_set_lr": (c_int, [c_void_p, c_float]),
    "kfp16_net_set_momentum": (c_int, [c_void_p, c_float]),
    "kfp16_net_get_lr": (c_float, [c_void_p]),
    "kfp16_net_set_dropout_seed": (c_int, [c_void_p, c_u32]),
    "kfp16_net_get_dropout_seed": (c_int, [c_void_p, C.POINTER(c_u32)]),
    "kfp16_net_grads_to_f16": (c_int, [c_void_p]),
    "kfp16_net_grads_f16": (c_void_p, [c_void_p]),
    "kfp16_net_sgd_step_f16": (c_int, [c_void_p]),
    "kfp16_net_read_loss": (c_int, [c_void_p, C.POINTER(c_float)]),
    "kfp16_wgrad_group_create": (c_void_p, [c_void_p, c_int, c_int, c_int, C.POINTER(WgradProb), c_int]),
    "kfp16_wgrad_group_launch": (c_int, [c_void_p, c_void_p]),
    "kfp16_wgrad_group_destroy": (None, [c_void_p]),
    "kfp16_net_capture_segments": (c_int, [c_void_p, c_int]),
    "kfp16_net_capture_segments_ex": (c_int, [c_void_p, c_int, C.c_char_p, c_int, c_int]),
    "kfp16_net_launch_segment": (c_int, [c_void_p, c_int]),
    "kfp16_net_segment_grads": (c_int, [c_void_p, c_int, C.POINTER(c_size_t), C.POINTER(c_size_t)]),
    "kfp16_net_read_loss_async": (c_int, [c_void_p, c_int]),
    "kfp16_net_wait_loss": (c_int, [c_void_p, c_int, C.POINTER(c_float)]),
    "kfp16_net_capture": (c_int, [c_void_p, c_int]),
    "kfp16_net_launch": (c_int, [c_void_p, c_int]),
    "kfp16_net_launches_per_step": (c_int, [c_void_p, c_int]),
    "kfp16_peer_comm_create": (c_void_p, [c_void_p, c_int, c_int, c_void_p, c_size_t]),
    "kfp16_peer_comm_handle": (c_int, [c_void_p, c_void_p]),
    "kfp16_peer_comm_connect": (c_int, [c_void_p, c_void_p]),
    "kfp16_peer_comm_set_timeout": (c_int, [c_void_p, C.c_double]),
    "kfp16_peer_allreduce_f16": (c_int, [c_void_p]),
    "kfp16_peer_allreduce_f16_range": (c_int, [c_void_p, c_size_t, c_size_t, c_int, c_int, c_int, c_void_p]),
    "kfp16_peer_comm_status": (c_int, [c_void_p]),
    "kfp16_peer_comm_destroy": (None, [c_void_p]),
    # ---- kaldi_fp16_chain.h
    "kfp16_chain_create": (c_void_p, [c_void_p, c_int, c_int, c_int, C.POINTER(ChainFst)]),
    "kfp16_chain_destroy": (None, [c_void_p]),
    "kfp16_chain_set_numerators": (c_int, [c_void_p, C.POINTER(ChainFst), c_int]),
    "kfp16_chain_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "kfp16_chain_read_results": (c_int, [c_void_p, c_void_p, c_int]),
    "kfp16_chain_force_general": (c_int, [c_void_p, c_int]),
    "kfp16_chain_set_debug": (c_int, [c_void_p, c_void_p]),
    # the reference's own chain interface (cpp/include/chain.h), exported on top of the kernels above (csrc/chain_compat.cu)
    "chain_forward_backward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, FP]),
    "chain_compute_posteriors": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p]),
    "chain_compute_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "chain_workspace_bytes": (c_size_t, [c_int, c_int]),
    "chain_last_error": (C.c_char_p, []),
    "chain_clear_error": (None, []),
    "kfp16_chain_num_sequences": (c_int, [c_void_p]),
    "kfp16_chain_frames": (c_int, [c_void_p]),
    "kfp16_net_loss_chain": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, c_int, c_float]),
    "kfp16_net_set_chain": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float]),
    "kfp16_net_set_sparse_output_grad": (c_int, [c_void_p, c_int]),
    "kfp16_net_set_fuse_conv_backward": (c_int, [c_void_p, c_int]),
    "kfp16_net_set_overlap_loss": (c_int, [c_void_p, c_int]),
    "kfp16_net_set_spec_augment": (c_int, [c_void_p, c_int]),
    "kfp16_net_set_train_batchnorm": (c_int, [c_void_p, c_int, c_float]),
    "kfp16_net_set_bn_stats_hook": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "kfp16_net_set_idct": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int]),
    "kfp16_net_get_bn": (c_int, [c_void_p, C.c_char_p, C.c_char_p, c_void_p, c_void_p, c_int]),
    # ---- kaldi_fp16_ops.h
    "ops_cublas_create": (c_void_p, []),
    "ops_cublas_destroy": (None, [c_void_p]),
    "ops_gemm": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_int, c_void_p, c_int, c_float, c_void_p, c_int]),
    "ops_gemm_strided": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_int, c_i64, c_void_p, c_int, c_i64, c_float, c_void_p, c_int, c_i64, c_int]),
    "ops_relu": (c_int, [c_void_p, c_int]),
    "ops_sigmoid": (c_int, [c_void_p, c_int]),
    "ops_tanh_act": (c_int, [c_void_p, c_int]),
    "ops_clipped_relu": (c_int, [c_void_p, c_int, c_float]),
    "ops_softmax": (c_int, [c_void_p, c_int, c_int]),
    "ops_log_softmax": (c_int, [c_void_p, c_int, c_int]),
    "ops_batchnorm_forward": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float]),
    "ops_batchnorm_forward_rms": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_float]),
    "ops_add_scaled": (c_int, [c_void_p, c_void_p, c_int, c_float, c_float]),
    "ops_add": (c_int, [c_void_p, c_void_p, c_int]),
    "ops_copy": (c_int, [c_void_p, c_void_p, c_int]),
    "ops_fill": (c_int, [c_void_p, c_int, c_float]),
    "ops_concat_cols": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int]),
    "ops_slice_cols": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int]),
    "ops_combine_feature_maps": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "ops_subsample_rows": (None, [c_void_p, c_void_p, c_int, c_int, c_int, c_int]),
    "ops_relu_backward": (c_int, [c_void_p, c_void_p, c_int]),
    "ops_sigmoid_backward": (c_int, [c_void_p, c_void_p, c_int]),
    "ops_tanh_backward": (c_int, [c_void_p, c_void_p, c_int]),
    "ops_transpose": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "ops_batchnorm_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int]),
    "ops_fp16_to_fp32": (c_int, [c_void_p, c_void_p, c_int]),
    "ops_sgd_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int]),
    "ops_last_error": (C.c_char_p, []),
    "ops_clear_error": (None, []),
    # ---- kaldi_fp16_cnn.h (go/kaldibridge surface)
    "kaldi_get_last_error": (C.c_char_p, []), "kaldi_clear_error": (None, []),
    "kaldi_cublas_create": (c_void_p, []), "kaldi_cublas_destroy": (None, [c_void_p]),
    "kaldi_cublas_enable_tensor_cores": (None, [c_void_p]),
    "kaldi_tensor_create": (c_void_p, [c_int, c_int]), "kaldi_tensor_zeros": (c_void_p, [c_int, c_int]),
    "kaldi_tensor_ones": (c_void_p, [c_int, c_int]), "kaldi_tensor_free": (None, [c_void_p]),
    "kaldi_tensor_rows": (c_int, [c_void_p]), "kaldi_tensor_cols": (c_int, [c_void_p]), "kaldi_tensor_size": (c_size_t, [c_void_p]),
    "kaldi_tensor_copy_from_host_fp32": (None, [c_void_p, c_void_p, c_size_t]),
    "kaldi_tensor_copy_to_host_fp32": (None, [c_void_p, c_void_p, c_size_t]),
    "kaldi_gemm": (None, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_int]),
    "kaldi_relu": (None, [c_void_p]), "kaldi_sigmoid": (None, [c_void_p]), "kaldi_tanh": (None, [c_void_p]),
    "kaldi_softmax": (None, [c_void_p]), "kaldi_add": (None, [c_void_p, c_void_p]), "kaldi_scale": (None, [c_void_p, c_float]),
    "kaldi_loss_scaler_create": (c_void_p, [c_float]), "kaldi_loss_scaler_free": (None, [c_void_p]),
    "kaldi_loss_scaler_get_scale": (c_float, [c_void_p]), "kaldi_loss_scaler_update": (None, [c_void_p, c_int]),
    "launch_conv1d_forward_fp16": (None, [c_void_p] * 4 + [c_int] * 8 + [c_void_p]),
    "launch_conv1d_backward_fp16": (None, [c_void_p] * 6 + [c_int] * 8 + [c_void_p]),
    "launch_maxpool1d_forward_fp16": (None, [c_void_p] * 3 + [c_int] * 5 + [c_void_p]),
    "launch_maxpool1d_backward_fp16": (None, [c_void_p] * 3 + [c_int] * 4 + [c_void_p]),
    "launch_batchnorm1d_forward_fp16": (None, [c_void_p] * 8 + [c_int] * 3 + [c_float, c_float, C.c_bool, c_void_p]),
    "launch_pointwise_conv1d_fp16": (None, [c_void_p] * 4 + [c_int] * 4 + [c_void_p]),
    # ---- kaldi_fp16_bridge.h
    "bridge_last_error": (C.c_char_p, []),
    "bridge_clear_error": (None, []),
    "bridge_gpu_init": (c_int, [c_int]),
    "bridge_gpu_get_free_memory": (c_int, [C.POINTER(c_size_t), C.POINTER(c_size_t)]),
    "bridge_gpu_sync": (c_int, []),
    "bridge_gpu_malloc": (c_void_p, [c_size_t]),
    "bridge_gpu_free": (None, [c_void_p]),
    "bridge_host_alloc": (c_void_p, [c_size_t]),
    "bridge_host_free": (None, [c_void_p]),
    "bridge_transfer_fp16": (c_int, [c_void_p, c_void_p, c_size_t]),
    "bridge_read_fp16": (c_int, [c_void_p, c_void_p, c_size_t]),
    "bridge_transfer_int32": (c_int, [c_void_p, c_void_p, c_size_t]),
    "bridge_transfer_float32": (c_int, [c_void_p, c_void_p, c_size_t]),
    "bridge_read_float32": (c_int, [c_void_p, c_void_p, c_size_t]),
    "bridge_batch_alloc": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, C.POINTER(GPUBatchPtrs)]),
    "bridge_batch_transfer": (c_int, [C.POINTER(GPUBatchPtrs), c_void_p, c_size_t]),
    "bridge_batch_free": (None, [C.POINTER(GPUBatchPtrs)]),
    "bridge_gpu_memset": (None, [c_void_p, c_int, c_size_t]),
    "bridge_fp16_to_fp32_gpu": (c_int, [c_void_p, c_void_p, c_size_t]),
    "bridge_fp32_to_fp16_gpu": (c_int, [c_void_p, c_void_p, c_size_t]),
}

_lib = None


def load() -> C.CDLL:
    """Load the native library and bind every declared symbol (raises if anything is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. "
            "Run `python -m kaldi_fp16_b200.build`; there is no CPU / PyTorch fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    missing = []
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing:
        raise ImportError(f"{LIB_PATH} does not export: {', '.join(missing)}")
    _lib = lib
    return lib


def last_error() -> str:
    lib = load()
    for getter in (lib.kfp16_last_error, lib.bridge_last_error):
        e = getter()
        if e:
            return e.decode(errors="replace")
    return "unknown error"


class NativeError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise NativeError(f"{what}: {last_error()}")
