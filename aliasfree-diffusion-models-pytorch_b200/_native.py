"""Builds and loads ``libafr_b200.so`` (hand-written CUDA for sm_100a behind the C ABI of
``include/afr.h``).  There is deliberately no fallback: if the library is missing and
cannot be built, or a tensor is not on a CUDA device, the ops raise."""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libafr_b200.so")
SOURCES = ["afr_api.cu", "afr_generic.cu", "afr_n3.cu", "afr_stripn.cu", "afr_rotate.cu", "afr_small.cu", "afr_actdown.cu", "afr_norm.cu", "afr_nhwc_resample.cu"]
HEADERS = ["afr_common.cuh", "afr_kernels.h", os.path.join("..", "..", "include", "afr.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared", "--threads", "0"]

_lock = threading.Lock()
_lib = None


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> libafr_b200.so (in-tree)."""
    if not force and not _stale():
        return LIB_PATH
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:          # several ranks may import at once
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():                # another process built it while we waited
                return LIB_PATH
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            tmp = LIB_PATH + ".tmp.%d" % os.getpid()
            cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + [os.path.join(CSRC, f) for f in SOURCES]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB_PATH)                     # atomic: readers never see a partial file
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def _declare(L):
    vp, ci, cd, cf, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_float, ctypes.c_int64
    L.afr_version.restype = ci
    L.afr_last_error.restype = ctypes.c_char_p
    L.afr_last_kernel.restype = ctypes.c_char_p
    L.afr_status_string.restype = ctypes.c_char_p
    L.afr_status_string.argtypes = [ci]
    L.afr_set_path.restype = ci
    L.afr_set_path.argtypes = [ci]
    L.afr_launch_count.restype = ctypes.c_uint64
    L.afr_up2x_fwd.argtypes = [vp, vp, ci, ci, ci, ci, vp, ci, ci, ci, vp]
    L.afr_up2x_bwd.argtypes = [vp, vp, ci, ci, ci, ci, vp, ci, ci, ci, vp]
    L.afr_up2x_fwd_strided.argtypes = [vp, vp, ci, ci, ci, ci, i64, vp, ci, ci, ci, vp]
    L.afr_up2x_bwd_strided.argtypes = [vp, vp, ci, ci, ci, ci, i64, vp, ci, ci, vp]
    L.afr_down2x_fwd.argtypes = [vp, vp, ci, ci, ci, ci, vp, ci, ci, vp]
    L.afr_down2x_bwd.argtypes = [vp, vp, ci, ci, ci, ci, vp, ci, ci, vp]
    L.afr_filtered_gelu_fwd.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp, ci, vp, ci, ci, vp]
    L.afr_filtered_gelu_affine_fwd.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, vp, ci, vp, ci, ci, vp]
    L.afr_groupnorm1_affine.argtypes = [vp, vp, vp, cf, vp, vp, ci, ci, ci, ci, ci, vp]
    L.afr_filtered_gelu_affine_bwd.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp, ci, vp, ci, ci, vp]
    L.afr_groupnorm1_stats.argtypes = [vp, vp, vp, cf, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]
    L.afr_affine_apply.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]
    L.afr_filtered_gelu_nhwc_fwd.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, vp, ci, vp, ci, ci, vp]
    L.afr_filtered_gelu_nhwc_bwd.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp, ci, vp, ci, ci, vp]
    L.afr_affine_apply_nhwc.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]
    L.afr_up2x_nhwc.argtypes = [vp, vp, ci, ci, ci, ci, i64, i64, vp, ci, ci, ci, ci, vp]
    L.afr_down2x_nhwc.argtypes = [vp, vp, ci, ci, ci, ci, i64, i64, vp, ci, ci, ci, vp]
    L.afr_groupnorm1_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp]
    L.afr_filtered_gelu_bwd.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, vp, ci, vp, ci, ci, vp]
    L.afr_gelu_down2x_fwd.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, vp, ci, ci, vp]
    L.afr_gelu_down2x_bwd.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp, ci, ci, vp]
    L.afr_rotate_periodic_cubic.argtypes = [vp, vp, ci, ci, ci, ci, cd, ci, vp]
    L.afr_ddpm_update.argtypes = [vp, vp, vp, i64, cf, cf, cf, vp]
    L.afr_ddpm_update_table.argtypes = [vp, vp, vp, i64, vp, vp, vp]
    for n in ("afr_up2x_fwd", "afr_up2x_bwd", "afr_down2x_fwd", "afr_down2x_bwd", "afr_up2x_fwd_strided", "afr_up2x_bwd_strided",
              "afr_filtered_gelu_fwd", "afr_filtered_gelu_bwd", "afr_filtered_gelu_affine_fwd", "afr_groupnorm1_affine",
              "afr_rotate_periodic_cubic", "afr_ddpm_update", "afr_ddpm_update_table", "afr_gelu_down2x_fwd",
              "afr_gelu_down2x_bwd", "afr_filtered_gelu_affine_bwd", "afr_groupnorm1_stats", "afr_affine_apply",
              "afr_filtered_gelu_nhwc_fwd", "afr_filtered_gelu_nhwc_bwd", "afr_affine_apply_nhwc", "afr_groupnorm1_bwd",
              "afr_up2x_nhwc", "afr_down2x_nhwc"):
        getattr(L, n).restype = ci
    return L


EXPORTS = ("afr_version", "afr_last_error", "afr_status_string", "afr_set_path", "afr_last_kernel",
           "afr_launch_count", "afr_up2x_fwd", "afr_up2x_bwd", "afr_up2x_fwd_strided", "afr_up2x_bwd_strided",
           "afr_down2x_fwd", "afr_down2x_bwd",
           "afr_filtered_gelu_fwd", "afr_filtered_gelu_bwd", "afr_filtered_gelu_affine_fwd", "afr_groupnorm1_affine",
           "afr_rotate_periodic_cubic", "afr_ddpm_update", "afr_ddpm_update_table", "afr_gelu_down2x_fwd",
           "afr_gelu_down2x_bwd", "afr_filtered_gelu_affine_bwd", "afr_groupnorm1_stats", "afr_affine_apply",
           "afr_filtered_gelu_nhwc_fwd", "afr_filtered_gelu_nhwc_bwd", "afr_affine_apply_nhwc", "afr_groupnorm1_bwd",
           "afr_up2x_nhwc", "afr_down2x_nhwc")


def lib():
    """The loaded C-ABI library.  ``import torch`` first so that libcudart.so.12 (shared with
    torch: one runtime, one current device) is already mapped."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                import torch  # noqa: F401  (maps libcudart.so.12)
                alt = os.environ.get("AFR_LIB_PATH")          # tuning builds (tools/), never set in production
                if alt:
                    _lib = _declare(ctypes.CDLL(alt, mode=ctypes.RTLD_GLOBAL))
                    return _lib
                if _stale():
                    build()
                _lib = _declare(ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL))
    return _lib


PATHS = {"auto": 0, "direct": 1, "tma": 2, "generic": 3, "direct_general": 4, "tma_general": 5}


def PATHS_AUTO():
    """True while the kernel-family selection is AUTO (the strided resamplers exist only there).
    An out-of-range argument makes afr_set_path a pure query."""
    return lib().afr_set_path(-1) == PATHS["auto"]


def set_path(name):
    """Select the kernel family for N==3 calls: auto | direct | tma | generic.  Returns the old name."""
    old = lib().afr_set_path(PATHS[name])
    return {v: k for k, v in PATHS.items()}[old]


def last_kernel():
    return lib().afr_last_kernel().decode()


def launch_count():
    return int(lib().afr_launch_count())
