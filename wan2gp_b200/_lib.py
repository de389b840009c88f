"""ctypes binding of libwan2gp_b200.so (C ABI: include/wan2gp_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwan2gp_b200.so")

c_void_p, c_int, c_ll, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float

# name -> argtypes  (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "b200_last_error": [],
    "b200_version": [],
    "b200_launch_count": [],
    "b200_gemm_bf16": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_void_p, c_void_p,
                       c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "b200_ln_modulate": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_float, c_void_p],
    "b200_rmsnorm_rope": [c_void_p, c_ll, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p],
    "b200_qk_rmsnorm_rope": [c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p],
    "b200_attention_d128": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_ll,
                            c_float, c_void_p],
    "b200_attention_d128_batched": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_float, c_void_p],
    "b200_cast_f32_bf16": [c_void_p, c_void_p, c_ll, c_void_p],
    "b200_patch_embed": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                         c_int, c_void_p],
    "b200_unpatchify": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "b200_gemv_f32": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "b200_sinusoid": [c_float, c_void_p, c_int, c_void_p],
    "b200_sinusoid_dev": [c_void_p, c_void_p, c_int, c_void_p],
    "b200_col_mean_f32": [c_void_p, c_void_p, c_int, c_int, c_void_p],
    "b200_add_vec": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "b200_embed_rows": [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p],
    "b200_t5_rmsnorm": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p],
    "b200_mul_bf16": [c_void_p, c_void_p, c_void_p, c_ll, c_void_p],
    "b200_t5_attention": [c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p],
    "b200_rope_half": [c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "b200_causal_gqa_attention": [c_void_p, c_void_p, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_float, c_void_p],
    "b200_cfg_euler_step": [c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_ll, c_void_p],
    "b200_cfg_euler_step_dev": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_void_p],
    "b200_cfg_unipc_step": [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_ll, c_void_p],
    "b200_cfg_unipc_step_dev": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_void_p],
    "b200_conv3d_cl": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                       c_int, c_int, c_int, c_int, c_void_p],
    "b200_conv_norm_fusable": [c_int, c_int, c_int, c_int, c_int],
    "b200_conv3d_cl_norm": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                            c_int, c_void_p],
    "b200_upconv2x_cl_norm": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "b200_conv3d_head_cl": [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "b200_conv3d_cl_stream": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_int, c_int, c_int, c_void_p],
    "b200_upconv2x_cl": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "b200_rms_silu_cl": [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p],
    "b200_upsample2x_cl": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "b200_vae_prologue": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "b200_attention_1head": [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_float, c_int, c_void_p],
    "b200_pad_replicate_cl": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "b200_rms_silu_pad_cl": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "b200_conv3d_cl_prepadded": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "b200_planar_to_cl": [c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p],
    "b200_hy_upsample_cl": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "b200_group_stats_cl": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_float, c_void_p],
    "b200_group_norm_apply_cl": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "b200_conv3d_cl_view": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                            c_int, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_void_p],
    "b200_hy_downsample_cl": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "b200_group_mean_cl": [c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p],
    "b200_blend_edge_f32": [c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "b200_planar_to_cl_pad": [c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p],
    "b200_space_to_depth_cl": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "b200_frames_to_u8": [c_void_p, c_void_p, c_ll, c_void_p],
    "b200_frames_to_u8_allgather": [c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p],
}
_RESTYPES = {"b200_last_error": ctypes.c_char_p, "b200_launch_count": c_ll}

_lib = None


class B200Error(RuntimeError):
    pass


def load():
    """Load the library (building is __graft_entry__.build()'s job).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(f"{LIB_PATH} not found: run `python -m wan2gp_b200.build` (there is no CPU / eager fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise B200Error(f"{name} failed ({rc}): {lib.b200_last_error().decode()}")


def query(name, *args):
    """Host-side query functions of the ABI (no launch, plain int result; no error convention)."""
    return int(getattr(load(), name)(*args))


def launch_count():
    return int(load().b200_launch_count())
