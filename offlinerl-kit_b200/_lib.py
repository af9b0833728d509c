"""ctypes binding of liborlk_b200.so (the C ABI declared in include/orlk_b200.h).

Loading fails loudly: there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liborlk_b200.so")
ABI_VERSION = 30

c_f32p = C.c_void_p     # all device pointers travel as integers
c_stream = C.c_void_p


class GemmDesc(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("C2", C.c_void_p),
                ("bias", C.c_void_p), ("aux", C.c_void_p), ("rowsum", C.c_void_p), ("colsum", C.c_void_p),
                ("CT", C.c_void_p),
                ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64), ("ldaux", C.c_int64), ("ldct", C.c_int64),
                ("c_split_stride", C.c_int64), ("sum_split_stride", C.c_int64),
                ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("a_layout", C.c_int32), ("b_layout", C.c_int32), ("epi", C.c_int32),
                ("k_splits", C.c_int32), ("k_chunk", C.c_int32), ("split_base", C.c_int32),
                ("tile_start", C.c_int32), ("tiles_m", C.c_int32), ("tiles_n", C.c_int32)]


class ConcatSeg(C.Structure):
    _fields_ = [("dst", C.c_void_p), ("src1", C.c_void_p), ("src2", C.c_void_p),
                ("ld_dst", C.c_int64), ("ld1", C.c_int64), ("ld2", C.c_int64),
                ("M", C.c_int32), ("w1", C.c_int32), ("w2", C.c_int32), ("rep1", C.c_int32),
                ("row_start", C.c_int32), ("pad_", C.c_int32)]


class AdamGroup(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("tau", C.c_float), ("step", C.c_int32), ("bc1", C.c_double), ("bc2_sqrt", C.c_float), ("pad_", C.c_int32)]

    def refresh(self) -> "AdamGroup":
        """Bias corrections of the next step (t = step + 1) from the (float) betas, as the device keeps them up to date."""
        t = self.step + 1
        self.bc1 = 1.0 - float(self.beta1) ** t
        self.bc2_sqrt = (1.0 - float(self.beta2) ** t) ** 0.5
        return self


class AdamDesc(C.Structure):
    _fields_ = [("p", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("tgt", C.c_void_p), ("grad", C.c_void_p),
                ("n", C.c_int64), ("g_split_stride", C.c_int64), ("g_splits", C.c_int32), ("group", C.c_int32),
                ("wd", C.c_float), ("block_start", C.c_int32), ("flags", C.c_int32), ("cols", C.c_int32),
                ("pT", C.c_void_p)]


class SampleUse(C.Structure):
    _fields_ = [("r0", C.c_int32), ("r1", C.c_int32), ("rep", C.c_int32), ("obs_dim", C.c_int32),
                ("eps", C.c_void_p), ("act", C.c_void_p), ("ld_act", C.c_int64), ("logp", C.c_void_p),
                ("obs", C.c_void_p), ("ld_obs", C.c_int64), ("xout", C.c_void_p), ("ld_x", C.c_int64)]


class TcGemm(C.Structure):
    _fields_ = [("A", C.c_void_p), ("lda", C.c_int64), ("a_gs", C.c_int64),
                ("B", C.c_void_p), ("ldb", C.c_int64), ("b_gs", C.c_int64),
                ("C", C.c_void_p), ("ldc", C.c_int64), ("c_gs", C.c_int64), ("c_split_stride", C.c_int64),
                ("CT", C.c_void_p), ("ldct", C.c_int64), ("ct_gs", C.c_int64),
                ("bias", C.c_void_p), ("bias_gs", C.c_int64),
                ("aux", C.c_void_p), ("ldaux", C.c_int64), ("aux_gs", C.c_int64),
                ("rowsum", C.c_void_p), ("rowsum_gs", C.c_int64), ("rowsum_split_stride", C.c_int64),
                ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("G", C.c_int32),
                ("epi", C.c_int32), ("k_splits", C.c_int32), ("passes", C.c_int32), ("n_tile", C.c_int32),
                ("gen_row", C.c_void_p), ("gen_row_gs", C.c_int64), ("gen_col", C.c_void_p), ("gen_col_gs", C.c_int64),
                ("a_mn", C.c_int32), ("b_mn", C.c_int32)]


FUSED_MAX_LAYERS = 4


class FusedFwd(C.Structure):
    _fields_ = [("X", C.c_void_p), ("ldx", C.c_int64), ("W0pad", C.c_void_p), ("W0pad_lo", C.c_void_p),
                ("W", C.c_void_p * FUSED_MAX_LAYERS), ("Wlo", C.c_void_p * FUSED_MAX_LAYERS),
                ("bias", C.c_void_p * FUSED_MAX_LAYERS), ("H", C.c_void_p * FUSED_MAX_LAYERS),
                ("gs", C.c_int64), ("h_gs", C.c_int64),
                ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("out", C.c_void_p), ("out_gs", C.c_int64), ("relu_bits", C.c_void_p),
                ("M", C.c_int32), ("N", C.c_int32), ("K0", C.c_int32), ("G", C.c_int32), ("n_hidden", C.c_int32),
                ("flags", C.c_int32)]


FUSED_PAIRS = 1


class FusedPrep(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst_lo", C.c_void_p), ("n", C.c_int64), ("W0", C.c_void_p), ("gs", C.c_int64),
                ("w0pad", C.c_void_p), ("N", C.c_int32), ("K0", C.c_int32), ("G", C.c_int32), ("pad_", C.c_int32)]


class FusedBwd(C.Structure):
    _fields_ = [("dq", C.c_void_p), ("dq_gs", C.c_int64), ("head_w", C.c_void_p), ("relu_bits", C.c_void_p),
                ("WT", C.c_void_p * FUSED_MAX_LAYERS), ("WTlo", C.c_void_p * FUSED_MAX_LAYERS),
                ("dZ", C.c_void_p * FUSED_MAX_LAYERS), ("gs", C.c_int64), ("dz_gs", C.c_int64),
                ("M", C.c_int32), ("N", C.c_int32), ("G", C.c_int32), ("n_hidden", C.c_int32),
                ("flags", C.c_int32), ("pad_", C.c_int32)]


EPI_NONE, EPI_RELU, EPI_RELU_MASK, EPI_SWISH, EPI_DSWISH = range(5)
CFG_BIG, CFG_MID, CFG_SMALL, CFG_KPAR, CFG_TINY = range(5)
CFG_TILES = {CFG_BIG: (128, 128, 16), CFG_MID: (64, 64, 16), CFG_SMALL: (32, 32, 32), CFG_KPAR: (32, 32, 256), CFG_TINY: (32, 16, 256)}
OPT_ADAM, OPT_POLYAK = 1, 2
SC_LOG_ALPHA, SC_ALPHA, SC_CQL_LOG_ALPHA, SC_COUNT = 0, 1, 2, 8

_I, _L, _F, _P = C.c_int, C.c_int64, C.c_float, C.c_void_p

# name -> argtypes (restype is always int unless noted)
_PROTOS = {
    "orlk_abi_version": [],
    "orlk_sizeof_gemm_desc": [], "orlk_sizeof_adam_desc": [], "orlk_sizeof_adam_group": [], "orlk_sizeof_concat_seg": [],
    "orlk_device_info": [_I, C.POINTER(C.c_int)],
    "orlk_graph_begin": [_P], "orlk_graph_end": [_P, C.POINTER(C.c_void_p)], "orlk_graph_launch": [_P, _P],
    "orlk_graph_destroy": [_P], "orlk_stream_sync": [_P], "orlk_graph_launch_sync": [_P, _P], "orlk_capture_status": [_P],
    "orlk_graph_launch_wait_event": [_P, _P, _P], "orlk_event_record_external": [_P, _P],
    "orlk_stream_create": [C.POINTER(C.c_void_p)], "orlk_stream_destroy": [_P], "orlk_stream_wait_event": [_P, _P],
    "orlk_event_create_notiming": [C.POINTER(C.c_void_p)],
    "orlk_memcpy_h2d_async": [_P, _P, C.c_size_t, _P], "orlk_memcpy_d2h_async": [_P, _P, C.c_size_t, _P],
    "orlk_memcpy_d2d_async": [_P, _P, C.c_size_t, _P], "orlk_memset_async": [_P, _I, C.c_size_t, _P],
    "orlk_event_create": [C.POINTER(C.c_void_p)], "orlk_event_record": [_P, _P], "orlk_event_sync": [_P],
    "orlk_event_elapsed_ms": [_P, _P, C.POINTER(C.c_float)], "orlk_event_destroy": [_P],
    "orlk_replay_pack": [_P, _P, _P, _P, _P, _L, _I, _I, _P, _I, _L, _P],
    "orlk_replay_gather": [_P, _L, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P],
    "orlk_replay_gather_into": [_P, _L, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P],
    "orlk_replay_sample": [_P, _L, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P],
    "orlk_gemm_grouped": [_P, _I, _I, _I, _I, _I, _P], "orlk_gemm_init": [],
    "orlk_gemm_tiny": [_P, _I, _I, _I, _I, _I, _P], "orlk_gemm_tiny_init": [],
    "orlk_head_sample": [_P, _L, _P, _L, _P, _P, _I, _I, _I, _P, _I, _P], "orlk_sizeof_sample_use": [],
    "orlk_actor_bwd_entry": [_P, _L, _I, _I, _P, _L, _I, _I, _P, _P, _P, _L, _P, _I, _I, _P, _P, _I, _P, _P, _P],
    "orlk_gemm_chain": [_P, _I, _I, _I, _I, _P], "orlk_gemm_chain_init": [],
    "orlk_tc_init": [], "orlk_tc_gemm": [C.POINTER(TcGemm), _P], "orlk_tc_effective_splits": [_I, _I], "orlk_tc_set_trace": [_P],
    "orlk_sizeof_tc_gemm": [],
    "orlk_fused_init": [], "orlk_critic_fwd_fused": [C.POINTER(FusedFwd), _I, _P], "orlk_sizeof_fused_fwd": [],
    "orlk_critic_bwd_fused": [C.POINTER(FusedBwd), _P], "orlk_sizeof_fused_bwd": [],
    "orlk_fused_prep_multi": [C.POINTER(FusedPrep), _I, _P], "orlk_sizeof_fused_prep": [],
    "orlk_fused_prep": [_P, _P, _L, _P, _L, _I, _I, _I, _P, _P],
    "orlk_skinny_fwd": [_P, _L, _L, _P, _L, _L, _L, _P, _L, _P, _L, _L, _I, _I, _I, _I, _P],
    "orlk_skinny_dgrad": [_P, _L, _L, _P, _L, _L, _P, _L, _L, _P, _L, _L, _P, _L, _L, _I, _I, _I, _I, _P],
    "orlk_concat_rows": [_P, _I, _I, _P],
    "orlk_compact_blocks": [_P, _L, _P, _I, _I, _P, _P],
    "orlk_narrow_fwd": [_P, _L, _L, _P, _L, _L, _P, _L, _P, _L, _L, _P, _L, _L, _I, _I, _I, _I, _I, _P],
    "orlk_narrow_wgrad_chunks": [_I], "orlk_narrow_init": [],
    "orlk_narrow_wgrad": [_P, _L, _L, _P, _L, _L, _P, _L, _L, _L, _L, _P, _L, _L, _P, _L, _L, _I, _I, _I, _I, _P],
    "orlk_philox_fill": [_P, _L, _L, _F, _F, C.c_uint64, _P, _P, _P],
    "orlk_tanh_gauss_sample": [_P, _L, _I, _I, _P, _I, _I, _P, _L, _P, _P, _L, _I, _P, _L, _P],
    "orlk_tanh_gauss_bwd": [_P, _L, _P, _P, _L, _P, _I, _L, _L, _P, _I, _I, _P, _L, _P],
    "orlk_edac_div": [_P, _I, _I, _I, _F, _P, _P, _P, _P],
    "orlk_sac_actor_loss": [_P, _L, _I, _P, _I, _P, _I, _I, _F, _P, _I, _P, _P, _L, _P, _P, _P],
    "orlk_cql_critic_loss": [_P, _L, _P, _L, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _F, _I, _I, _F, _P, _P, _I, _P,
                             _P, _L, _P, _P, _P],
    "orlk_cql_critic_loss_scratch_floats": [_I, _I],
    "orlk_twin_head_actor": [_P, _L, _P, _L, _P, _L, _P, _I, _I, _P, _L, _P, _L, _P, _P, _L, _P],
    "orlk_segment_max": [_P, _L, _I, _I, _I, _P, _L, _P],
    "orlk_td_loss": [_P, _L, _I, _P, _L, _I, _P, _P, _I, _P, _P, _I, _F, _P, _L, _P, _P, _P, _P],
    "orlk_iql_v_loss": [_P, _L, _P, _I, _F, _P, _P, _P, _P],
    "orlk_iql_actor_loss": [_P, _L, _P, _P, _L, _P, _P, _I, _I, _F, _F, _P, _L, _P, _P, _P],
    "orlk_det_actor_fwd": [_P, _L, _P, _I, _I, _F, _F, _F, _P, _L, _P, _L, _I, _P, _L, _P],
    "orlk_td3bc_actor_loss": [_P, _P, _L, _P, _L, _I, _I, _F, _P, _P, _L, _P, _P],
    "orlk_det_actor_bwd": [_P, _L, _P, _L, _P, _L, _I, _I, _F, _P, _L, _P],
    "orlk_dyn_input": [_P, _L, _P, _L, _P, _P, _I, _I, _I, _P, _L, _P],
    "orlk_gather_rows": [_P, _L, _I, _P, _L, _L, _I, _I, _P, _L, _L, _P],
    "orlk_sumsq_chunks": [_L], "orlk_sumsq": [_P, _L, _F, _P, _P],
    "orlk_dyn_nll": [_P, _P, _I, _I, _I, _P, _P, _F, _P, _I, _P, _P, _P, _P, _P, _P],
    "orlk_dyn_nll_scratch_floats": [_I, _I, _I],
    "orlk_dyn_val_mse": [_P, _P, _I, _I, _I, _P, _P],
    "orlk_dyn_step": [_P, _I, _I, _I, _P, _P, _P, _L, _P, _P, _P, _P, _P, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P],
    "orlk_compact_rows": [_P, _I, _P, _L, _I, _P, _L, _P, _P],
    "orlk_adam_step": [_P, _I, _I, _P, _P],
    "orlk_step_end": [_P, C.c_uint, _P, _P],
}

EXPORTS = tuple(_PROTOS) + ("orlk_last_error",)


class OrlkError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once) and check that the struct mirrors match the C side."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OrlkError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). offlinerlkit_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, args in _PROTOS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.orlk_last_error.argtypes = []
    lib.orlk_last_error.restype = C.c_char_p
    if lib.orlk_abi_version() != ABI_VERSION:
        raise OrlkError(f"ABI mismatch: library {lib.orlk_abi_version()} vs binding {ABI_VERSION}; rebuild")
    for fn, st in (("orlk_sizeof_gemm_desc", GemmDesc), ("orlk_sizeof_adam_desc", AdamDesc),
                   ("orlk_sizeof_adam_group", AdamGroup), ("orlk_sizeof_concat_seg", ConcatSeg),
                   ("orlk_sizeof_tc_gemm", TcGemm), ("orlk_sizeof_sample_use", SampleUse),
                   ("orlk_sizeof_fused_fwd", FusedFwd), ("orlk_sizeof_fused_bwd", FusedBwd),
                   ("orlk_sizeof_fused_prep", FusedPrep)):
        if getattr(lib, fn)() != C.sizeof(st):
            raise OrlkError(f"struct size mismatch for {st.__name__}: C {getattr(lib, fn)()} vs ctypes {C.sizeof(st)}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().orlk_last_error().decode("utf-8", "replace")
        raise OrlkError(f"orlk call failed ({what}) rc={rc}: {msg}")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
