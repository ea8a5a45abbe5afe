"""ctypes binding of libvfd_b200.so (C-ABI declared in include/vfd_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C vfd_gan_b200/csrc``.
There is no CPU fallback: if the shared object is missing, every op raises.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvfd_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

_p, _ll, _i, _f, _ull = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float, ctypes.c_ulonglong

# name -> argument ctypes, in the order of include/vfd_b200.h
SIGNATURES = {
    "vfd_conv3d_fwd": [_p, _ll, _i, _p, _i, _i, _p, _p, _ll, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "vfd_conv3d_wgrad": [_p, _ll, _i, _p, _ll, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "vfd_conv3d_wgrad_layout": [_i, _i, _i, _i, _i, _i, _i],
    "vfd_conv3d_wgrad_thin": [_p, _ll, _i, _p, _ll, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "vfd_conv3d_fwd_narrow": [_p, _ll, _i, _p, _i, _p, _p, _ll, _i, _i, _i, _i, _i, _p],
    "vfd_conv3d_dgrad_narrow": [_p, _ll, _p, _i, _i, _p, _ll, _i, _i, _i, _i, _p],
    "vfd_conv3d_wgrad_narrow": [_p, _ll, _p, _ll, _p, _i, _i, _i, _i, _i, _p],
    "vfd_conv3d_wgrad_first": [_p, _ll, _i, _p, _ll, _p, _i, _i, _i, _i, _i, _p],
    "vfd_convlstm_step_fwd": [_p, _ll, _i, _p, _i, _p, _p, _i, _p, _p, _ll, _p, _i, _i, _i, _i, _i, _i, _p],
    "vfd_conv3d_wgrad_det_workspace": [_i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i],
    "vfd_conv3d_wgrad_det": [_p, _ll, _i, _p, _ll, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _ll, _p],
    "vfd_conv3d_wgrad_thin_det_workspace": [_i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i],
    "vfd_conv3d_wgrad_thin_det": [_p, _ll, _i, _p, _ll, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _ll, _p],
    "vfd_pack_ncdhw": [_p, _p, _i, _i, _ll, _i, _ll, _i, _i, _p],
    "vfd_unpack_ncdhw": [_p, _i, _p, _i, _i, _ll, _ll, _p],
    "vfd_pack_weight": [_p, _p, _i, _i, _i, _i, _i, _i, _p],
    "vfd_pack_weights_batched": [_p, _i, _ll, _p],
    "vfd_unpack_wgrad": [_p, _p, _i, _i, _i, _i, _i, _i, _p],
    "vfd_bn_stats": [_p, _ll, _i, _ll, _p, _p],
    "vfd_bn_finalize": [_p, _i, _i, _ll, _p, _p, _p, _p, _p, _f, _f, _i, _p, _p, _p, _p, _p],
    "vfd_bn_act_fwd": [_p, _ll, _i, _i, _i, _i, _i, _p, _p, _f, _p, _ll, _p, _ll, _i, _i, _i, _f, _ull, _p, _p],
    "vfd_bn_act_bwd": [_p, _ll, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _f, _p, _ll, _p, _ll, _i, _i, _i,
                       _f, _ull, _p, _i, _p, _p, _p, _p, _p, _p, _ll, _p, _p],
    "vfd_tap_gather": [_p, _ll, _i, _p, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "vfd_channel_sum": [_p, _ll, _i, _ll, _p, _p],
    "vfd_upsample2x_fwd": [_p, _ll, _i, _i, _i, _i, _i, _p, _ll, _p],
    "vfd_upsample2x_bwd": [_p, _ll, _i, _i, _i, _i, _i, _p, _ll, _p, _ll, _p],
    "vfd_sigmoid_head_fwd": [_p, _ll, _ll, _p, _p],
    "vfd_sigmoid_head_bwd": [_p, _p, _ll, _p, _p],
    "vfd_weighted_bce": [_p, _p, _ll, _f, _f, _p, _p, _p],
    "vfd_sqdiff": [_p, _ll, _p, _ll, _i, _ll, _p, _p],
    "vfd_convlstm_cell_fwd": [_p, _ll, _p, _i, _ll, _p, _p, _p, _p],
    "vfd_convlstm_cell_bwd": [_p, _p, _p, _p, _p, _i, _ll, _p, _ll, _p, _p],
    "vfd_latent_score": [_p, _ll, _p, _ll, _i, _ll, _i, _p, _p],
    "vfd_sqdiff_bwd": [_p, _ll, _p, _ll, _i, _ll, _p, _f, _p, _ll, _p, _ll, _p],
    "vfd_l1_loss": [_p, _p, _ll, _f, _p, _p, _p],
    "vfd_bce_loss": [_p, _p, _ll, _f, _p, _p, _p],
    "vfd_score_finalize": [_p, _i, ctypes.c_double, _p, _p, _p],
    "vfd_score_scale": [_p, _ll, _p, _p, _p],
    "vfd_threshold_open": [_p, _i, _i, _i, _i, _f, _p, _p, _p],
    "vfd_confusion_counts": [_p, _p, _ll, _f, _p, _p],
    "vfd_roc_auc": [_p, _p, _i, _p, _p],
    "vfd_roc_auc_large": [_p, _p, _ll, _p, _p, _ll, _p],
    "vfd_roc_auc_large_workspace": [_ll],
    "vfd_resize_frames_u8_workspace": [_ll, _i, _i, _i, _i, _i],
    "vfd_resize_frames_u8": [_p, _ll, _i, _i, _i, _p, _i, _i, _p, _ll, _p],
    "vfd_frames_to_clip": [_p, _ll, _i, _i, _i, _i, _i, _i, _p, _p],
    "vfd_video_to_flow": [_p, _i, _i, _i, _i, _p, _p, _p, _ll, _p],
    "vfd_video_to_flow_workspace": [_i, _i, _i, _i],
}
# test / tool-only entry points of libvfd_b200_debug.so (declared in csrc/conv_direct.cu and csrc/conv_tc.cu)
DEBUG_SIGNATURES = {
    "vfd_conv3d_fwd_direct": [_p, _ll, _i, _p, _i, _i, _p, _p, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "vfd_conv3d_wgrad_direct": [_p, _ll, _i, _p, _ll, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "vfd_set_debug": [_i],
}
DEBUG_LIB_PATH = os.path.join(_HERE, "libvfd_b200_debug.so")
_debug_lib = None

_lib = None
LAUNCHES = 0         # C-ABI compute calls issued by this process
KERNEL_LAUNCHES = 0  # CUDA kernels those calls launched (bench.py reports it as gpu_launches)
_KERNELS_PER_CALL = {"vfd_bn_act_bwd": 2, "vfd_upsample2x_bwd": 3, "vfd_roc_auc_large": 18, "vfd_conv3d_wgrad_det": 3, "vfd_conv3d_wgrad_thin_det": 2, "vfd_resize_frames_u8": 4}


def build(force=False):
    """Compile libvfd_b200.so for sm_100a with the in-tree Makefile (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", CSRC_DIR, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libvfd_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(vfd_gan_b200 has no CPU fallback)")
        # tools/gpu_stage_probe.py sets VFD_DEBUG_LIB=1: every entry point then comes from libvfd_b200_debug.so (a
        # superset of the product library), whose stage-isolation switches act on the kernels it launches
        L = ctypes.CDLL(os.environ.get("VFD_LIB_OVERRIDE") or (DEBUG_LIB_PATH if os.environ.get("VFD_DEBUG_LIB") == "1" else LIB_PATH))
        L.vfd_last_error.restype = ctypes.c_char_p
        L.vfd_last_error.argtypes = []
        L.vfd_abi_version.restype = ctypes.c_int
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
        L.vfd_video_to_flow_workspace.restype = ctypes.c_longlong
        L.vfd_roc_auc_large_workspace.restype = ctypes.c_longlong
        L.vfd_conv3d_wgrad_det_workspace.restype = ctypes.c_longlong
        L.vfd_conv3d_wgrad_thin_det_workspace.restype = ctypes.c_longlong
        L.vfd_resize_frames_u8_workspace.restype = ctypes.c_longlong
        _lib = L
    return _lib


def debug_lib():
    """libvfd_b200_debug.so: the product sources built with -DVFD_DEBUG plus the CUDA-core cross-check convs. Only
    tests/ and tools/ reach it (``ops.CONV_IMPL_DIRECT``, tools/gpu_stage_probe.py)."""
    global _debug_lib
    if _debug_lib is None and os.environ.get("VFD_DEBUG_LIB") == "1":
        _debug_lib = lib()     # one image, one set of switches
    if _debug_lib is None:
        if not os.path.exists(DEBUG_LIB_PATH):
            raise RuntimeError(f"{DEBUG_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = ctypes.CDLL(DEBUG_LIB_PATH)
        L.vfd_last_error.restype = ctypes.c_char_p
        L.vfd_last_error.argtypes = []
        for table in (SIGNATURES, DEBUG_SIGNATURES):
            for name, args in table.items():
                fn = getattr(L, name)
                fn.argtypes = args
                fn.restype = ctypes.c_int
        _debug_lib = L
    return _debug_lib


def call_debug(name, *args):
    """Invoke an entry point of the debug library (cross-check kernels)."""
    L = debug_lib()
    rc = getattr(L, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {L.vfd_last_error().decode()}")


def call(name, *args):
    """Invoke a C-ABI entry point; a non-zero status becomes RuntimeError (the reference's only
    error convention on this path is torch raising RuntimeError)."""
    global LAUNCHES, KERNEL_LAUNCHES
    L = lib()
    rc = getattr(L, name)(*args)
    LAUNCHES += 1
    KERNEL_LAUNCHES += _KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {L.vfd_last_error().decode()}")
