"""ctypes binding of libnplane.so (include/nplane.h) and the in-tree nvcc build recipe.

The product path has NO CPU fallback: if the shared library is missing or a call fails, a RuntimeError
carrying np_last_error() is raised.
"""
import ctypes as C
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_PKG, "_lib")
LIB_PATH = os.environ.get("NPLANE_LIB") or os.path.join(LIB_DIR, "libnplane.so")
CSRC = os.path.join(_PKG, "csrc")
HEADER = os.path.join(_PKG, "..", "include", "nplane.h")

NVCC_FLAGS = ["-O3", "-std=c++17", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-fmad=false", "-shared", "-Xcompiler", "-fPIC"]

NP_OK = 0
TASK_IDS = {"heading": 0, "control": 1, "tracking": 2}
MODEL_IDS = {"F16": 0, "UAV": 1}
NUM_NETS, NUM_OBS, NUM_DRAWS, NUM_COUNTERS = 43, 22, 5, 8
NUM_OBS_COMBAT = 15
COMBAT_RECORD_FLOATS = 28
COUNTER_NAMES = ("overload", "low_altitude", "high_speed", "low_speed", "extreme_state", "unreach", "reached", "resets")


class NetDesc(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("sel", C.c_int32 * 3), ("n_layers", C.c_int32), ("dims", C.c_int32 * 5),
                ("w_off", C.c_int32), ("used", C.c_int32)]


class EnvCfg(C.Structure):
    _fields_ = [("n", C.c_int32), ("ld", C.c_int32), ("task", C.c_int32), ("use_coef_cache", C.c_int32),
                ("seed", C.c_uint64), ("index_base", C.c_uint64),
                ("dt", C.c_float), ("airspeed", C.c_float), ("noise_scale", C.c_float),
                ("altitude_limit", C.c_float), ("acceleration_limit", C.c_float), ("max_velocity", C.c_float),
                ("min_velocity", C.c_float),
                ("min_alpha", C.c_float), ("max_alpha", C.c_float), ("min_beta", C.c_float), ("max_beta", C.c_float),
                ("max_heading_increment", C.c_float), ("max_pitch_increment", C.c_float),
                ("max_velocities_u_increment", C.c_float),
                ("max_distance", C.c_float), ("min_distance", C.c_float),
                ("max_check_interval", C.c_int32), ("min_check_interval", C.c_int32),
                ("init_T", C.c_float), ("max_altitude", C.c_float), ("min_altitude", C.c_float),
                ("max_vt", C.c_float), ("min_vt", C.c_float), ("model", C.c_int32), ("max_steps", C.c_int32),
                ("distance_limit", C.c_float), ("target_dist", C.c_float), ("max_heading", C.c_float),
                ("min_heading", C.c_float), ("max_npos", C.c_float), ("min_npos", C.c_float), ("max_epos", C.c_float),
                ("min_epos", C.c_float), ("index_stride", C.c_int32), ("combat_pairs_per_env", C.c_int32), ("combat_reward_scale", C.c_float)]


class Buffers(C.Structure):
    _fields_ = [("s_dev", C.c_void_p), ("u_dev", C.c_void_p), ("tgt_dev", C.c_void_p), ("step_count_dev", C.c_void_p),
                ("flags_dev", C.c_void_p), ("obs_dev", C.c_void_p), ("reward_dev", C.c_void_p),
                ("workspace_dev", C.c_void_p)]


# every symbol include/nplane.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "np_version": (C.c_int, []),
    "np_last_error": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "np_aero_create": (C.c_int, [_P, C.c_size_t, C.POINTER(NetDesc), _P, C.c_int, C.POINTER(_P)]),
    "np_aero_destroy": (C.c_int, [_P]),
    "np_aero_pack_host": (C.c_int, [_P, C.c_size_t, C.POINTER(NetDesc), _P, C.c_int, _P, C.c_size_t,
                                    C.POINTER(C.c_size_t)]),
    "np_env_workspace_bytes": (C.c_size_t, [C.POINTER(EnvCfg)]),
    "np_env_create": (C.c_int, [C.POINTER(EnvCfg), _P, C.POINTER(_P)]),
    "np_env_bind": (C.c_int, [_P, C.POINTER(Buffers), _P]),
    "np_env_set_cfg": (C.c_int, [_P, C.POINTER(EnvCfg)]),
    "np_env_destroy": (C.c_int, [_P]),
    "np_env_reset": (C.c_int, [_P, _P, _P, _P]),
    "np_env_step": (C.c_int, [_P, _P, _P, _P, _P]),
    "np_env_step_range": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "np_env_step_host": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_int), C.c_int, _P]),
    "np_env_step_mapped": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, _P]),
    "np_f16_update": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_double, _P]),
    "np_f16_table_update": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_double, _P]),
    "np_uav_update": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_double, _P]),
    "np_env_plan_step": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "np_env_combat_step": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "np_env_combat_records": (C.c_int, [_P, _P, _P]),
    "np_env_combat_role_local": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "np_env_combat_role_pair": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "np_env_pair_reset_offset_bytes": (C.c_size_t, [C.POINTER(EnvCfg)]),
    "np_combat_relgeo": (C.c_int, [_P, _P, _P, _P, C.c_int, _P]),
    "np_combat_relgeo_peers": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P]),
    "np_env_blood_offset_bytes": (C.c_size_t, [C.POINTER(EnvCfg)]),
    "np_env_pid_offset_bytes": (C.c_size_t, [C.POINTER(EnvCfg)]),
    "np_env_set_pid_started": (C.c_int, [_P, C.c_int]),
    "np_env_counters": (C.c_int, [_P, C.POINTER(C.c_uint64), _P]),
    "np_env_rng_advance": (C.c_int, [_P, C.c_uint32, _P]),
    "np_env_launch_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "np_f16_nlplant": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "np_tables_create": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_size_t, C.POINTER(_P)]),
    "np_tables_destroy": (C.c_int, [_P]),
    "np_f16_table_coeffs": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "np_env_create_tables": (C.c_int, [_P, _P, _P]),
    "np_env_rebind_outputs": (C.c_int, [_P, _P, _P]),
    "np_rollout_masks": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "np_rollout_returns": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _P]),
    "np_f16_table_nlplant": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "np_uav_nlplant": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P]),
    "np_f16_coeffs": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
}

_lib = None


def build(force=False, verbose=False, extra_flags=()):
    """Compile csrc/nplane.cu for sm_100a into _lib/libnplane.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-o", LIB_PATH, os.path.join(CSRC, "nplane.cu")]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


def lib():
    """The loaded library; raises loudly when it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the F-16 step has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError here == header / library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def last_error():
    buf = C.create_string_buffer(1024)
    lib().np_last_error(buf, 1024)
    return buf.value.decode(errors="replace")


def check(status, what):
    if status != NP_OK:
        raise RuntimeError(f"{what} failed (status {status}): {last_error()}")
