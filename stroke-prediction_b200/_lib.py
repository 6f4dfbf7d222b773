"""Build + ctypes binding of libstroke_b200.so (the C-ABI declared in include/stroke_b200.h).

The library is compiled in-tree with nvcc for sm_100a only and loaded with ctypes; there is no CPU fallback: if the
shared object is missing and cannot be built, importing any compute op raises.
"""
import ctypes
import os
import shutil
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libstroke_b200.so")
SOURCES = ["sp_api.cu", "sp_conv.cu", "sp_norm.cu", "sp_resample.cu", "sp_loss.cu", "sp_optim.cu", "sp_metrics.cu", "sp_augment.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]

c_f32p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_float = ctypes.c_float
c_double = ctypes.c_double
c_vp = ctypes.c_void_p
c_size = ctypes.c_size_t


class SpConvDesc(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int32),
                ("Di", ctypes.c_int32), ("Hi", ctypes.c_int32), ("Wi", ctypes.c_int32), ("Ci", ctypes.c_int32),
                ("ldi", ctypes.c_int32),
                ("Do", ctypes.c_int32), ("Ho", ctypes.c_int32), ("Wo", ctypes.c_int32), ("Co", ctypes.c_int32),
                ("ldo", ctypes.c_int32),
                ("k", ctypes.c_int32), ("s", ctypes.c_int32),
                ("pd", ctypes.c_int32), ("ph", ctypes.c_int32), ("pw", ctypes.c_int32),
                ("act", ctypes.c_int32), ("alpha", ctypes.c_float)]


class SpAdamTensor(ctypes.Structure):
    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("n", ctypes.c_int64), ("block_start", ctypes.c_int64)]


SP_ADAM_CHUNK = 4096
ACT_NONE, ACT_ELU, ACT_LEAKY, ACT_SIGMOID = 0, 1, 2, 3

_D = ctypes.POINTER(SpConvDesc)

# name -> (restype, argtypes); must list every symbol include/stroke_b200.h declares (tests check this)
SIGNATURES = {
    "sp_version": (c_int, []),
    "sp_last_error": (ctypes.c_char_p, []),
    "sp_get_tc_terms": (c_int, []),
    "sp_set_tc_terms": (c_int, [c_int]),
    "sp_set_wgrad_tc_options": (c_int, [c_int, c_int]),
    "sp_packed_weight_floats": (c_size, [_D, c_int]),
    "sp_pack_weights": (c_int, [_D, c_int, c_vp, c_vp, c_vp]),
    "sp_conv_workspace_bytes": (c_size, [_D, c_int]),
    "sp_corr": (c_int, [_D, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_size, c_vp]),
    "sp_corrT": (c_int, [_D, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_size, c_vp]),
    "sp_wgrad_workspace_bytes": (c_size, [_D]),
    "sp_wgrad": (c_int, [_D, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_float, c_vp, c_size, c_vp]),
    "sp_bias_grad": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_float, c_vp, c_vp]),
    "sp_bn_stats": (c_int, [c_vp, c_int, c_i64, c_int, c_int, c_int, c_vp, c_vp]),
    "sp_bn_finalize": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_float, c_float, c_int,
                               c_vp, c_vp, c_vp, c_vp, c_vp]),
    "sp_bn_bwd_reduce": (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_i64, c_int, c_int, c_vp, c_vp]),
    "sp_bn_bwd_finalize": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_float, c_vp, c_vp]),
    "sp_bn_act_bwd_apply": (c_int, [c_vp, c_int, c_vp, c_int, c_vp, c_int, c_i64, c_int, c_int, c_int, c_float,
                                    c_vp, c_int, c_int, c_vp, c_vp]),
    "sp_bias_from_colsum": (c_int, [c_vp, c_int, c_vp, c_float, c_vp]),
    "sp_bn_grads_from_wgrad": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_float, c_vp, c_vp, c_float,
                                       c_vp]),
    "sp_border_tap_sums": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "sp_maxpool2_fwd": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "sp_maxpool2_bwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "sp_upsample2_fwd": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_vp]),
    "sp_upsample2_bwd": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_vp]),
    "sp_crop_copy": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_int, c_int, c_int, c_vp]),
    "sp_crop_add": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int,
                            c_int, c_int, c_int, c_vp]),
    "sp_ncdhw_to_ndhwc": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_vp]),
    "sp_ndhwc_to_ncdhw": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_vp]),
    "sp_dice_sums": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "sp_dice_loss": (c_int, [c_vp, c_float, c_float, c_vp, c_vp]),
    "sp_dice_bwd": (c_int, [c_vp, c_vp, c_i64, c_vp, c_float, c_float, c_vp, c_float, c_vp, c_int, c_vp]),
    "sp_absdiff_mean": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    "sp_absdiff_bwd": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_float, c_vp, c_int, c_vp, c_int, c_vp]),
    "sp_binary_counts": (c_int, [c_vp, c_vp, c_i64, c_float, c_vp, c_vp]),
    "sp_surface_distances_workspace_bytes": (c_size, [c_i64]),
    "sp_surface_distances": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_float, c_vp, c_vp, c_size, c_vp]),
    "sp_signed_distance_workspace_bytes": (c_size, [c_i64]),
    "sp_signed_distance": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_float, c_int, c_float, c_vp, c_vp, c_size, c_vp]),
    "sp_gauss3d": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_double, c_double, c_vp, c_vp, c_vp]),
    "sp_elastic_warp": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_double, c_double, c_vp, c_vp]),
    "sp_zoom_plane_xy": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "sp_flip_w": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    "sp_pad_volume": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_vp, c_vp]),
    "sp_latent_interp_fwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_i64, c_vp, c_vp]),
    "sp_latent_interp_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp]),
    "sp_adam_multi": (c_int, [c_vp, c_int, c_i64, c_double, c_double, c_double, c_double, c_double, c_i64, c_double,
                              c_int, c_vp]),
}

_lib = None
_lock = threading.Lock()


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_ROOT, "include", "stroke_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into libstroke_b200.so (in-tree, so the .so travels with the repo)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libstroke_b200.so")
    objdir = os.path.join(_HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose and out.strip():
            print(out)
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc, "-shared", "-o", tmp] + objs + ["-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


def load():
    """Return the ctypes handle (building the library first if it is missing and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            build()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        if lib.sp_version() != 100:
            raise RuntimeError("libstroke_b200.so version mismatch")
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().sp_last_error()
        raise RuntimeError("libstroke_b200 %s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))
