"""ctypes binding of the C ABI in include/dirb200.h (the only way Python reaches the kernels).

There is deliberately no fallback: if libdirb200.so is missing or no sm_100a device is present,
every compute entry point raises. The oracle under oracle/ is never imported from here.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DIRB200_LIB") or os.path.join(_HERE, "libdirb200.so")  # override: A/B of builds

PRECISION = {"fp32": 0, "bf16": 1, "tf32": 2}
BACKBONE = {"resnet50": 0, "hrnet_w32": 32, "hrnet_w48": 48}
DTYPE_F32, DTYPE_I64 = 0, 1
STAGE_FLOATS = 4887
RECORD_FLOATS = 3 * STAGE_FLOATS
# offsets inside one stage of the record (include/dirb200.h)
OFF = {"mesh_l": 0, "mesh_r": 2334, "joint_l": 4668, "joint_r": 4731, "uv_l": 4794, "uv_r": 4836,
       "proj_l": 4878, "proj_r": 4881, "offset": 4884}

EXPORTS = [
    "dirb200_create", "dirb200_destroy", "dirb200_last_error", "dirb200_set_weight", "dirb200_finalize_weights",
    "dirb200_num_required_keys", "dirb200_required_key", "dirb200_workspace_bytes", "dirb200_forward",
    "dirb200_forward_launches", "dirb200_backbone", "dirb200_residual", "dirb200_init_regressor", "dirb200_mano",
    "dirb200_joint2bone", "dirb200_bone_proj", "dirb200_nccl_unique_id", "dirb200_nccl_init",
    "dirb200_allgather_records", "dirb200_profile_layer", "dirb200_profile_read", "dirb200_conv_layer", "dirb200_profile_dump", "dirb200_forward_u8", "dirb200_preprocess_u8", "dirb200_eval_metrics",
    "dirb200_img2joint", "dirb200_gcn", "dirb200_ste", "dirb200_regressor_offset",
    "dirb200_bone_fusion",
]


class Config(C.Structure):
    _fields_ = [("precision", C.c_int), ("max_batch", C.c_int), ("aux_outputs", C.c_int), ("device", C.c_int),
                ("refine_stages", C.c_int), ("backbone", C.c_int)]


class Outputs(C.Structure):
    _fields_ = [("record", C.c_void_p), ("mano_para", C.c_void_p), ("seg", C.c_void_p), ("dense", C.c_void_p),
                ("proj_feat", C.c_void_p)]


class DirB200Error(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen libdirb200.so (built in-tree by `make` / __graft_entry__.build()) and declare signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DirB200Error(f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()); "
                           "there is no CPU/PyTorch fallback")
    lib = C.CDLL(LIB_PATH)
    vp, ip, fp = C.c_void_p, C.c_int, C.c_float
    lib.dirb200_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.dirb200_destroy.argtypes = [vp]
    lib.dirb200_destroy.restype = None
    lib.dirb200_last_error.argtypes = [vp]
    lib.dirb200_last_error.restype = C.c_char_p
    lib.dirb200_set_weight.argtypes = [vp, C.c_char_p, vp, ip, ip, C.POINTER(C.c_int64)]
    lib.dirb200_finalize_weights.argtypes = [vp, vp]
    lib.dirb200_num_required_keys.argtypes = [vp]
    lib.dirb200_required_key.argtypes = [vp, ip]
    lib.dirb200_required_key.restype = C.c_char_p
    lib.dirb200_workspace_bytes.argtypes = [vp, ip, C.POINTER(C.c_size_t)]
    lib.dirb200_forward.argtypes = [vp, vp, ip, vp, C.c_size_t, C.POINTER(Outputs), vp]
    lib.dirb200_forward_u8.argtypes = [vp, vp, ip, vp, C.c_size_t, C.POINTER(Outputs), vp]
    lib.dirb200_preprocess_u8.argtypes = [vp, vp, ip, ip, ip, vp, vp]
    lib.dirb200_eval_metrics.argtypes = [vp, vp, vp, vp, vp, vp, ip, ip, vp, vp, vp, vp, vp, vp]
    lib.dirb200_forward_launches.argtypes = [vp, ip]
    lib.dirb200_backbone.argtypes = [vp, vp, ip, ip, ip, vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.dirb200_residual.argtypes = [vp, C.c_char_p, vp, ip, ip, ip, ip, vp, vp, C.c_size_t, vp]
    lib.dirb200_init_regressor.argtypes = [vp, vp, ip, vp, vp, vp, C.c_size_t, vp]
    lib.dirb200_mano.argtypes = [vp, ip, vp, ip, vp, vp]
    lib.dirb200_joint2bone.argtypes = [vp, ip, vp, vp, vp, ip, vp, vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.dirb200_img2joint.argtypes = [vp, ip, vp, vp, vp, ip, vp, vp, vp, C.c_size_t, vp]
    lib.dirb200_gcn.argtypes = [vp, ip, vp, vp, ip, vp, vp, vp, C.c_size_t, vp]
    lib.dirb200_ste.argtypes = [vp, ip, vp, ip, vp, vp]
    lib.dirb200_regressor_offset.argtypes = [vp, ip, vp, vp, vp, vp, vp, ip, vp, vp, vp, C.c_size_t, vp]
    lib.dirb200_bone_fusion.argtypes = [vp, ip, vp, vp, vp, vp, ip, vp, vp, C.c_size_t, vp]
    lib.dirb200_bone_proj.argtypes = [vp, vp, vp, ip, ip, fp, vp, vp]
    lib.dirb200_nccl_unique_id.argtypes = [vp, C.c_char_p]
    lib.dirb200_nccl_init.argtypes = [vp, C.c_char_p, ip, ip]
    lib.dirb200_allgather_records.argtypes = [vp, vp, vp, ip, vp]
    lib.dirb200_conv_layer.argtypes = [vp, C.c_char_p, vp, vp, ip, ip, ip, vp, C.POINTER(C.c_int), vp, C.c_size_t, vp]
    lib.dirb200_profile_dump.argtypes = [vp, C.c_char_p, C.c_size_t]
    lib.dirb200_profile_layer.argtypes = [vp, C.c_char_p]
    lib.dirb200_profile_read.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    for name in EXPORTS:
        if name not in ("dirb200_destroy", "dirb200_last_error", "dirb200_required_key"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


class Handle:
    """Owns one dirb200_handle*. All methods raise DirB200Error on a non-zero return code."""

    def __init__(self, precision="fp32", max_batch=128, aux_outputs=True, device=0, refine_stages=2, backbone="resnet50"):
        self.lib = load_library()
        cfg = Config(PRECISION[precision], int(max_batch), int(bool(aux_outputs)), int(device), int(refine_stages),
                     BACKBONE[backbone])
        h = C.c_void_p()
        rc = self.lib.dirb200_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise DirB200Error(f"dirb200_create failed ({rc}): {self.lib.dirb200_last_error(None).decode()}")
        self.h = h
        self.device = int(device)
        self.precision = precision
        self.aux_outputs = bool(aux_outputs)
        self.max_batch = int(max_batch)

    def close(self):
        if getattr(self, "h", None):
            self.lib.dirb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, what):
        if rc != 0:
            raise DirB200Error(f"{what} failed ({rc}): {self.lib.dirb200_last_error(self.h).decode()}")

    def required_keys(self):
        n = self.lib.dirb200_num_required_keys(self.h)
        return [self.lib.dirb200_required_key(self.h, i).decode() for i in range(n)]

    def set_weight(self, name, ptr, dtype, shape):
        arr = (C.c_int64 * max(len(shape), 1))(*shape)
        self.check(self.lib.dirb200_set_weight(self.h, name.encode(), C.c_void_p(ptr), dtype, len(shape), arr),
                   f"set_weight({name})")

    def finalize(self, stream):
        self.check(self.lib.dirb200_finalize_weights(self.h, C.c_void_p(stream)), "finalize_weights")

    def profile_layer(self, prefix):
        self.check(self.lib.dirb200_profile_layer(self.h, None if prefix is None else prefix.encode()),
                   "profile_layer")

    def profile_dump(self):
        buf = C.create_string_buffer(1 << 20)
        self.check(self.lib.dirb200_profile_dump(self.h, buf, len(buf)), "profile_dump")
        rows = []
        for line in buf.value.decode().splitlines():
            f = line.split("\t")
            rows.append({"layer": f[0], "tc": int(f[1]), "kernel": f[2], "stride": f[3], "cin": int(f[4]),
                         "cout": int(f[5]), "ms": float(f[6]), "flops": float(f[7]), "bytes": float(f[8])})
        return rows

    def profile_read(self):
        ms, n, fl = C.c_float(), C.c_int(), C.c_double()
        self.check(self.lib.dirb200_profile_read(self.h, C.byref(ms), C.byref(n), C.byref(fl)), "profile_read")
        return ms.value, n.value, fl.value

    def workspace_bytes(self, batch):
        n = C.c_size_t()
        self.check(self.lib.dirb200_workspace_bytes(self.h, batch, C.byref(n)), "workspace_bytes")
        return n.value
