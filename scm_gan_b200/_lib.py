"""ctypes binding of libscmgan.so (include/scmgan.h).

This is the only place the shared library is touched.  There is deliberately no fallback: if the library is
missing or a call fails, a RuntimeError is raised (the product path never routes through torch/CPU code).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libscmgan.so")

ACT_NONE, ACT_LRELU, ACT_SIGMOID = 0, 1, 2
FMT_BF16, FMT_F16 = 0, 1

c_f32p = C.c_void_p  # device pointers travel as integers


class PackJob(C.Structure):
    _fields_ = [("w", C.c_void_p), ("out", C.c_void_p), ("sigma", C.c_void_p),
                ("n_pad", C.c_int), ("k_pad", C.c_int), ("n_valid", C.c_int), ("k_valid", C.c_int),
                ("s_n", C.c_longlong), ("s_k", C.c_longlong), ("k_src_off", C.c_int), ("flip", C.c_int),
                ("out_ld", C.c_int), ("fmt", C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("x", C.c_void_p), ("x_cs", C.c_int), ("x_c_off", C.c_int), ("cin", C.c_int),
                ("w", C.c_void_p), ("n", C.c_int),
                ("scale", C.c_float), ("bias", C.c_void_p), ("sample_bias", C.c_void_p),
                ("act", C.c_int), ("slope", C.c_float),
                ("out", C.c_void_p), ("out_cs", C.c_int), ("out_c_off", C.c_int), ("wrap", C.c_int),
                ("add", C.c_void_p), ("add_cs", C.c_int), ("add_c_off", C.c_int),
                ("gate", C.c_void_p), ("gate_cs", C.c_int), ("gate_c_off", C.c_int),
                ("out_f32", C.c_void_p), ("n_valid", C.c_int),
                ("sample_out", C.c_void_p), ("uniforms", C.c_void_p), ("rng_state", C.c_void_p),
                ("bias_n", C.c_int), ("x_fmt", C.c_int), ("w_fmt", C.c_int), ("out_fmt", C.c_int),
                ("sample_scale", C.c_void_p), ("coord_c1", C.c_int), ("weights_stable", C.c_int)]


class DecoderBceDesc(C.Structure):
    _fields_ = [("conv", ConvDesc), ("target", C.c_void_p), ("target_bstride", C.c_longlong),
                ("target_tstride", C.c_longlong), ("mask", C.c_void_p), ("mask_bstride", C.c_longlong),
                ("mask_tstride", C.c_longlong), ("T", C.c_int), ("B", C.c_int), ("loss_t", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_longlong)]


class WgradReduceJob(C.Structure):
    _fields_ = [("ws", C.c_void_p), ("splits", C.c_int), ("n", C.c_int), ("g", C.c_void_p),
                ("g_sm", C.c_longlong), ("g_sn", C.c_longlong), ("g_st", C.c_longlong),
                ("flip", C.c_int), ("m_valid", C.c_int), ("n_valid", C.c_int), ("scale", C.c_float),
                ("ws_bias", C.c_void_p), ("db", C.c_void_p), ("lanes", C.c_int)]


class WgradDesc(C.Structure):
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("dy", C.c_void_p), ("dy_cs", C.c_int), ("dy_c_off", C.c_int), ("cout", C.c_int),
                ("x", C.c_void_p), ("x_cs", C.c_int), ("x_c_off", C.c_int), ("cin", C.c_int),
                ("g", C.c_void_p), ("g_s_co", C.c_longlong), ("g_s_ci", C.c_longlong), ("g_s_tap", C.c_longlong),
                ("flip", C.c_int), ("co_valid", C.c_int), ("ci_valid", C.c_int), ("scale", C.c_float),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_longlong), ("db", C.c_void_p),
                ("defer_jobs", C.POINTER(WgradReduceJob)), ("defer_cap", C.c_int), ("defer_count", C.POINTER(C.c_int)),
                ("workspace_cursor", C.POINTER(C.c_longlong)), ("dy_fmt", C.c_int), ("x_fmt", C.c_int)]


class SnLayer(C.Structure):
    _fields_ = [("w", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("sigma", C.c_void_p),
                ("u_save", C.c_void_p), ("v_save", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int)]


class SnBwdLayer(C.Structure):
    _fields_ = [("g", C.c_void_p), ("wbar", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
                ("sigma", C.c_void_p), ("dot", C.c_void_p), ("out", C.c_void_p),
                ("rows", C.c_int), ("cols", C.c_int), ("accumulate", C.c_int), ("sigma2", C.c_void_p)]


class CsrnSweepDesc(C.Structure):
    _fields_ = [("x", C.c_void_p), ("xs_b", C.c_longlong), ("xs_c", C.c_longlong), ("xs_line", C.c_longlong),
                ("xs_pix", C.c_longlong), ("B", C.c_int), ("C", C.c_int), ("L", C.c_int), ("n", C.c_int),
                ("reverse", C.c_int), ("w_ih", C.c_void_p), ("w_hh", C.c_void_p), ("conv_w", C.c_void_p),
                ("conv_b", C.c_void_p), ("ctx", C.c_void_p), ("states", C.c_void_p), ("dctx", C.c_void_p),
                ("dx", C.c_void_p), ("dparams", C.c_void_p)]


class ReplayDesc(C.Structure):
    _fields_ = [("frames", C.c_void_p), ("rewards", C.c_void_p), ("actions", C.c_void_p), ("ep_len", C.c_void_p),
                ("n_filled", C.c_void_p), ("slots", C.c_int), ("max_len", C.c_int), ("R", C.c_int),
                ("per_frame", C.c_longlong), ("B", C.c_int), ("Hn", C.c_int), ("random_start", C.c_int),
                ("rng_state", C.c_void_p), ("states", C.c_void_p), ("rewards_out", C.c_void_p), ("dones", C.c_void_p),
                ("actions_out", C.c_void_p), ("plan", C.c_void_p)]


class AdamChunk(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
                ("n", C.c_int), ("clip", C.c_float), ("step", C.c_void_p)]


# name -> (restype, argtypes); must list every symbol include/scmgan.h declares (tests check this)
SIGNATURES = {
    "scmgan_version": (C.c_int, []),
    "scmgan_last_error": (C.c_char_p, []),
    "scmgan_num_sms": (C.c_int, []),
    "scmgan_launch_count": (C.c_longlong, []),
    "scmgan_pack_nchw": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "scmgan_pack_weights": (C.c_int, [C.c_int, C.POINTER(PackJob), C.c_void_p]),
    "scmgan_conv3x3_fwd": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "scmgan_conv3x3_dgrad": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "scmgan_conv3x3_wgrad": (C.c_int, [C.POINTER(WgradDesc), C.c_void_p]),
    "scmgan_wgrad_workspace_bytes": (C.c_longlong, []),
    "scmgan_plane_colsum": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "scmgan_spectral_norm_fwd": (C.c_int, [C.c_int, C.POINTER(SnLayer), C.c_void_p]),
    "scmgan_spectral_norm_fwd_n": (C.c_int, [C.c_int, C.POINTER(SnLayer), C.c_int, C.c_int, C.c_void_p]),
    "scmgan_spectral_norm_bwd": (C.c_int, [C.c_int, C.POINTER(SnBwdLayer), C.c_void_p]),
    "scmgan_action_bias": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p]),
    "scmgan_action_wgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "scmgan_bce_logits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_longlong,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "scmgan_coord_wgrad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong,
                                     C.c_longlong, C.c_int, C.c_void_p]),
    "scmgan_replay_sample": (C.c_int, [C.POINTER(ReplayDesc), C.c_void_p]),
    "scmgan_eval_sqerr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_longlong,
                                    C.c_void_p, C.c_void_p]),
    "scmgan_eval_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p,
                                    C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "scmgan_reward_head_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "scmgan_reward_head_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p]),
    "scmgan_decoder_bce_fwd": (C.c_int, [C.POINTER(DecoderBceDesc), C.c_void_p]),
    "scmgan_decoder_bce_workspace_rows": (C.c_int, []),
    "scmgan_decoder_bce_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p]),
    "scmgan_cf_loss_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scmgan_cf_loss_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scmgan_transition_tail": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scmgan_masked_mse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int,
                                    C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scmgan_bce_logits_seq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong,
                                        C.c_longlong, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "scmgan_masked_mse_seq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong,
                                        C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "scmgan_wgrad_reduce": (C.c_int, [C.c_int, C.POINTER(WgradReduceJob), C.c_void_p]),
    "scmgan_pack_coords": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "scmgan_gru_conv_sweep_fwd": (C.c_int, [C.POINTER(CsrnSweepDesc), C.c_void_p]),
    "scmgan_gru_conv_sweep_bwd": (C.c_int, [C.POINTER(CsrnSweepDesc), C.c_void_p]),
    "scmgan_philox_uniform": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "scmgan_clip_adam": (C.c_int, [C.c_int, C.POINTER(AdamChunk), C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.c_int, C.c_void_p, C.c_float, C.c_void_p]),
}

_lib = None


def lib():
    """Load libscmgan.so (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no non-CUDA fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().scmgan_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()
