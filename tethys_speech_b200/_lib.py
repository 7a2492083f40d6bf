"""ctypes binding of libtethys.so (include/tethys.h). There is no CPU fallback: if the library or a B200 is
missing, creating a context raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtethys.so")

TS_F32, TS_BF16, TS_I32, TS_I64 = 0, 1, 2, 3

_STATUS = {0: "TS_OK", -1: "TS_EINVAL", -2: "TS_ESHAPE", -3: "TS_EDTYPE", -4: "TS_ECUDA", -5: "TS_ENCCL",
           -6: "TS_EUNSUPPORTED", -7: "TS_EWATCHDOG"}


class TethysError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_STATUS.get(code, code)}: {msg}")
        self.code = code


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p),
        ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
        ("a_major", C.c_int32), ("b_major", C.c_int32),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64),
        ("batch1", C.c_int32), ("batch2", C.c_int32),
        ("a_bs1", C.c_int64), ("a_bs2", C.c_int64), ("b_bs1", C.c_int64), ("b_bs2", C.c_int64),
        ("c_bs1", C.c_int64), ("c_bs2", C.c_int64),
        ("in_dtype", C.c_int32), ("out_dtype", C.c_int32),
        ("alpha", C.c_float),
        ("bias", C.c_void_p),
        ("act", C.c_int32),
        ("residual", C.c_void_p),
        ("ldr", C.c_int64), ("r_bs1", C.c_int64), ("r_bs2", C.c_int64),
        ("accumulate", C.c_int32),
        ("c_preact", C.c_void_p),
        ("force_engine", C.c_int32),
        ("drop", C.c_float),
        ("seed", C.c_uint64),
        ("bias_bs1", C.c_int64),
        ("act_aux", C.c_void_p), ("ld_aux", C.c_int64),
        ("gn_accum", C.c_void_p), ("gn_rows_per_batch", C.c_int32), ("gn_valid_rows", C.c_int32), ("gn_groups", C.c_int32),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
        ("q_ld", C.c_int64), ("q_bs", C.c_int64), ("kv_ld", C.c_int64), ("kv_bs", C.c_int64),
        ("o_ld", C.c_int64), ("o_bs", C.c_int64),
        ("stats", C.c_void_p),
        ("batch", C.c_int32), ("heads", C.c_int32), ("tq", C.c_int32), ("tk", C.c_int32), ("head_dim", C.c_int32),
        ("scale", C.c_float), ("mask_mode", C.c_int32), ("drop", C.c_float),
        ("seed", C.c_uint64),
        ("d_o", C.c_void_p), ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("dq_ld", C.c_int64), ("dq_bs", C.c_int64), ("dkv_ld", C.c_int64), ("dkv_bs", C.c_int64),
        ("dsum", C.c_void_p),
        ("o_lo", C.c_void_p), ("dq_accum", C.c_void_p),
    ]


class W2VConfig(C.Structure):
    _fields_ = [
        ("hidden", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32), ("ffn", C.c_int32),
        ("n_conv", C.c_int32),
        ("conv_dim", C.c_int32 * 8), ("conv_kernel", C.c_int32 * 8), ("conv_stride", C.c_int32 * 8),
        ("pos_kernel", C.c_int32), ("pos_groups", C.c_int32),
        ("cv_groups", C.c_int32), ("cv_per_group", C.c_int32), ("cv_dim", C.c_int32), ("proj_dim", C.c_int32),
        ("num_negatives", C.c_int32),
        ("ln_eps", C.c_float), ("temperature", C.c_float), ("diversity_weight", C.c_float),
        ("hidden_dropout", C.c_float), ("activation_dropout", C.c_float), ("attention_dropout", C.c_float),
        ("head", C.c_int32), ("vocab_size", C.c_int32), ("classifier_proj", C.c_int32), ("num_labels", C.c_int32),
    ]


class WhisperCfg(C.Structure):
    _fields_ = [
        ("d_model", C.c_int32), ("enc_layers", C.c_int32), ("dec_layers", C.c_int32), ("heads", C.c_int32),
        ("d_ff", C.c_int32), ("n_mels", C.c_int32), ("n_ctx", C.c_int32), ("vocab", C.c_int32),
        ("max_target", C.c_int32), ("start_token", C.c_int32),
        ("ln_eps", C.c_float), ("dropout", C.c_float), ("attention_dropout", C.c_float),
        ("activation_dropout", C.c_float),
    ]


class StepArgs(C.Structure):
    """ts_step_args: optimizer / replica arguments of the composite entries ts_w2v_step / ts_whisper_step."""
    _fields_ = [
        ("optim", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("step", C.c_int32), ("global_clip", C.c_float), ("clipnorm", C.c_float),
        ("seed", C.c_uint64),
        ("dropout", C.c_int32), ("reserved", C.c_int32),
        ("comm", C.c_void_p), ("grads_bf16", C.c_void_p), ("scratch_dev", C.c_void_p), ("loss_out_dev", C.c_void_p),
    ]


# every symbol include/tethys.h declares: name -> (restype, argtypes)
_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    "ts_version": (_I, []),
    "ts_create": (_I, [_I, C.POINTER(_P)]),
    "ts_destroy": (None, [_P]),
    "ts_last_error": (C.c_char_p, [_P]),
    "ts_watchdog_check": (_I, [_P]),
    "ts_launch_count": (_L, [_P]),
    "ts_simt_downgrades": (_L, [_P]),
    "ts_debug_gemm_trace": (_I, [_P, _P]),
    "ts_step_state_set": (_I, [_P, C.c_uint64, _L, _P]),
    "ts_step_state_advance": (_I, [_P, _P]),
    "ts_step_state_get": (_I, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]),
    "ts_comm_unique_id": (_I, [_P, _P]),
    "ts_comm_init": (_I, [_P, _P, _I, _I, C.POINTER(_P)]),
    "ts_comm_info": (_I, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "ts_comm_alloc": (_I, [_P, _L, C.POINTER(_P)]),
    "ts_comm_free": (_I, [_P, _P]),
    "ts_comm_broadcast": (_I, [_P, _P, _L, _I, _I, _P]),
    "ts_comm_allreduce_bucket": (_I, [_P, _P, _L, _I, _P, _P]),
    "ts_comm_check": (_I, [_P]),
    "ts_comm_finalize": (_I, [_P]),
    "ts_gemm": (_I, [_P, C.POINTER(GemmDesc), _P]),
    "ts_attn_fwd": (_I, [_P, C.POINTER(AttnDesc), _P]),
    "ts_attn_bwd": (_I, [_P, C.POINTER(AttnDesc), _P]),
    "ts_logmel_num_frames": (_I, [_I]),
    "ts_logmel": (_I, [_P, _P, _L, _I, _I, _P, _I, _I, _P]),
    "ts_layernorm_fwd": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P]),
    "ts_layernorm_bwd": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "ts_groupnorm_gelu_fwd": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "ts_groupnorm_fwd": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "ts_w2v_sample_negatives": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "ts_ctc_workspace_floats": (_L, [_I, _I, _I]),
    "ts_ctc_loss": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _F, _I, _P]),
    "ts_span_mask_apply": (_I, [_P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ts_optim_create": (_I, [_P, _I, C.POINTER(_L), C.POINTER(_I), C.POINTER(_I), C.POINTER(_L), _L, C.POINTER(_P)]),
    "ts_optim_destroy": (None, [_P]),
    "ts_optim_clip_global": (_I, [_P, _P, _F, _P, _P]),
    "ts_optim_global_clip_scale": (_I, [_P, _P, _F, _P, _P]),
    "ts_optim_step": (_I, [_P, _P, _P, _P, _P, _P, _F, _F, _F, _F, _I, _F, _F, _I, _P]),
    "ts_optim_step_lp": (_I, [_P, _P, _P, _P, _P, _P, _F, _F, _F, _F, _I, _F, _F, _I, _P]),
    "ts_cast_f32_to_bf16": (_I, [_P, _P, _P, _L, _P]),
    "ts_grad_pack_bf16": (_I, [_P, _P, _P, _L, _P, _P]),
    "ts_grad_unpack_bf16": (_I, [_P, _P, _P, _L, _P]),
    "ts_dropout": (_I, [_P, _I, _P, _P, _L, _F, C.c_uint64, _P]),
    "ts_w2v_create": (_I, [_P, C.POINTER(W2VConfig), _I, C.POINTER(_P)]),
    "ts_w2v_destroy": (None, [_P]),
    "ts_w2v_arena_elems": (_L, [_P]),
    "ts_w2v_num_params": (_I, [_P]),
    "ts_w2v_param_info": (_I, [_P, _I, C.c_char_p, _I, C.POINTER(_L), C.POINTER(_I), C.POINTER(_L), C.POINTER(_L)]),
    "ts_w2v_num_stages": (_I, [_P]),
    "ts_w2v_stage_end": (_L, [_P, _I]),
    "ts_w2v_workspace_bytes": (_L, [_P, _I, _I]),
    "ts_w2v_bind": (_I, [_P, _P, _P, _P, _P, _L]),
    "ts_w2v_sync_compute_weights": (_I, [_P, _P]),
    "ts_w2v_forward": (_I, [_P, _P, _I, _I, _P, _L, _L, _F, C.c_uint64, _I, _P]),
    "ts_w2v_forward_head": (_I, [_P, _P, _I, _I, _P, _F, C.c_uint64, _I, _I, _P]),
    "ts_w2v_forward_features": (_I, [_P, _P, _I, _I, _P]),
    "ts_w2v_backward": (_I, [_P, _I, _I, _P]),
    "ts_w2v_step": (_I, [_P, _P, _I, _I, _P, _L, _L, _P, C.POINTER(StepArgs), _P]),
    "ts_w2v_get_buffer": (_I, [_P, C.c_char_p, C.POINTER(_P), C.POINTER(_I), C.POINTER(_I), C.POINTER(_L)]),
    "ts_whisper_create": (_I, [_P, C.POINTER(WhisperCfg), _I, C.POINTER(_P)]),
    "ts_whisper_destroy": (None, [_P]),
    "ts_whisper_arena_elems": (_L, [_P]),
    "ts_whisper_num_params": (_I, [_P]),
    "ts_whisper_param_info": (_I, [_P, _I, C.c_char_p, _I, C.POINTER(_L), C.POINTER(_I), C.POINTER(_L), C.POINTER(_L)]),
    "ts_whisper_num_stages": (_I, [_P]),
    "ts_whisper_stage_end": (_L, [_P, _I]),
    "ts_whisper_workspace_bytes": (_L, [_P, _I, _I, _I]),
    "ts_whisper_bind": (_I, [_P, _P, _P, _P, _P, _L]),
    "ts_whisper_sync_compute_weights": (_I, [_P, _P]),
    "ts_whisper_forward": (_I, [_P, _P, _I, _I, _P, _I, C.c_uint64, _I, _I, _P]),
    "ts_whisper_encode": (_I, [_P, _P, _I, _I, _I, _P]),
    "ts_whisper_decode_step": (_I, [_P, _P, _L, _I, _P]),
    "ts_whisper_backward": (_I, [_P, _I, _I, _P]),
    "ts_whisper_step": (_I, [_P, _P, _I, _I, _P, _I, C.POINTER(StepArgs), _P]),
    "ts_whisper_get_buffer": (_I, [_P, C.c_char_p, C.POINTER(_P), C.POINTER(_I), C.POINTER(_I), C.POINTER(_L)]),
}

_lib = None


def load():
    """dlopen libtethys.so and declare every prototype. Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found — run `make` (or __graft_entry__.build()); there is no fallback path")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class Context:
    """One per (process, GPU) — wraps ts_ctx."""

    def __init__(self, device=0):
        self.lib = load()
        h = _P()
        rc = self.lib.ts_create(int(device), C.byref(h))
        if rc != 0:
            raise TethysError(rc, f"ts_create(device={device}) failed — a B200 (sm_100) GPU is required; no CPU fallback exists")
        self.h = h
        self.device = int(device)

    def check(self, rc):
        if rc != 0:
            raise TethysError(rc, self.lib.ts_last_error(self.h).decode("utf-8", "replace"))

    def watchdog(self):
        self.check(self.lib.ts_watchdog_check(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.ts_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts = {}


def context(device=0):
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]
