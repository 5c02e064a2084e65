"""Drop-in host module: same constructor, forward signature and state_dict keys as the reference's
`DIR` (models/dir.py:486-596), with the eval forward executed by the sm_100a kernels behind the
C ABI (include/dirb200.h). `apps/eval.py:104-111,168` works unchanged with
`from dir_b200 import DIR`.

    network = DIR(21, './misc/mano').cuda()
    network.load_state_dict(state, strict=False); network.eval()
    result, _ = network({'img': imgTensors}, None, None)

Host code is PyTorch only for plumbing (device memory, streams, state_dict); no torch op touches
the data path. The training branch (models/dir.py:542-594) is out of scope and raises.
"""
import ctypes as C
import json
import math
import os

import torch
import torch.nn as nn

from . import assets, capi

_HERE = os.path.dirname(os.path.abspath(__file__))
_KEYS_JSON = os.path.join(_HERE, "state_dict_keys.json")
_KEYS_JSON_HRNET = {w: os.path.join(_HERE, f"state_dict_keys_hrnet_w{w}.json") for w in (32, 48)}

STAGE_KEYS = [  # (key in outs_list[i], offset name, trailing shape) — models/dir.py:521-535
    ("pd_joint_uv_left", "uv_l", (21, 2)), ("pd_joint_uv_right", "uv_r", (21, 2)),
    ("pd_mesh_xyz_left", "mesh_l", (778, 3)), ("pd_mesh_xyz_right", "mesh_r", (778, 3)),
    ("pd_joint_xyz_left", "joint_l", (21, 3)), ("pd_joint_xyz_right", "joint_r", (21, 3)),
    ("pd_proj_left", "proj_l", (3,)), ("pd_proj_right", "proj_r", (3,)), ("pd_offset", "offset", (3,)),
]


class _Node(nn.Module):
    """Anonymous container; the tree only exists to reproduce the reference's state_dict key names."""


def _is_buffer(name):
    leaf = name.rsplit(".", 1)[-1]
    return leaf in ("running_mean", "running_var", "num_batches_tracked", "img_gird") or "mano_layer_" in name \
        or name == "seg_loss.weight"


def _default_init(name, shape):
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros((), dtype=torch.long)
    if leaf == "th_faces":
        return torch.zeros(shape, dtype=torch.long)
    if leaf == "running_var" or leaf in ("e_0", "e_1"):
        return torch.ones(shape)
    if leaf == "weight" and len(shape) == 1:
        return torch.ones(shape)
    if leaf in ("weight", "W") and len(shape) >= 2:
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        return torch.randn(shape) * math.sqrt(2.0 / fan_in)
    if leaf == "img_gird":
        s = int(round(math.sqrt(shape[0])))
        r = torch.arange(s, dtype=torch.float32) + 0.5
        gx, gy = torch.meshgrid(r, r, indexing="ij")
        return torch.stack((gy, gx), dim=-1).reshape(s * s, 2).contiguous()
    return torch.zeros(shape)


def reference_key_shapes(backbone="resnet50"):
    """Key/shape inventory: the reference's 963 keys, or (backbone='hrnet_w32' | 'hrnet_w48') the HRNet extension's."""
    with open(_KEYS_JSON if backbone == "resnet50" else _KEYS_JSON_HRNET[int(backbone.split("_w")[1])]) as f:
        return json.load(f)


class DIR(nn.Module):
    """B200-native DIR. Extra keyword arguments (all optional, defaults keep the reference call valid):
    precision 'fp32' (default: the reference's numerics, <=1e-4 relative; error-compensated 3xTF32 tcgen05 convs) |
    'bf16' (explicit opt-in: bf16 feature maps, fastest; drifts like the reference under bf16 autocast);
    aux_outputs: also return seg/dense/proj_feat (outs_list[3]);
    max_batch: larger batches are processed in chunks; use_cuda_graph: capture the forward once per batch size and replay
    it (the returned tensors are then the graph's own output buffers, double-buffered: valid until the call after next);
    refine_stages: 2 = the reference forward (stage_num 3); 1 = init regression + projecter_4 only ("1 refine iter" of
    BASELINE.json configs[0]; outs_list then holds two stage dicts; needs aux_outputs=False)."""

    def __init__(self, joint_num, mano_path, root_joint=0, precision="fp32", aux_outputs=True, max_batch=128,
                 use_cuda_graph=False, refine_stages=2, backbone="resnet50"):
        super().__init__()
        if joint_num != 21:
            raise ValueError("DIR is defined for the 21-joint hand skeleton (models/dir.py:25-26)")
        if root_joint != 0:
            raise ValueError("only root_joint=0 (wrist) is built; the reference default (config.py:10)")
        self.joint_num = joint_num
        self.mano_path = mano_path
        self.precision = precision
        self.aux_outputs = bool(aux_outputs)
        self.max_batch = int(max_batch)
        self.use_cuda_graph = bool(use_cuda_graph)
        self.refine_stages = int(refine_stages)
        if backbone not in capi.BACKBONE:
            raise ValueError(f"backbone must be one of {sorted(capi.BACKBONE)} (the reference has ResNet-50 only; "
                             "'hrnet_w32' / 'hrnet_w48' are extensions with a self-authored oracle)")
        self.backbone_name = backbone
        if self.refine_stages not in (1, 2):
            raise ValueError("the reference defines two refinement stages (models/dir.py:437-471): refine_stages is 1 or 2")
        if self.refine_stages == 1 and self.aux_outputs:
            raise ValueError("refine_stages=1 needs aux_outputs=False (seg/dense/proj_feat hang off the second stage)")
        for name, shape in reference_key_shapes(backbone).items():
            parts = name.split(".")
            m = self
            for p in parts[:-1]:
                if p not in m._modules:
                    m.add_module(p, _Node())
                m = m._modules[p]
            t = _default_init(name, tuple(shape))
            if _is_buffer(name):
                m.register_buffer(parts[-1], t)
            else:
                m.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))
        self._handle = None
        self._loaded_keys = None
        # manolayer.py:62-101 reads MANO_{LEFT,RIGHT}.pkl at construction; the released checkpoint also embeds the
        # same buffers, so the pickles are optional here: read them when they exist, else expect them in the state_dict
        self._asset_keys = set()
        if assets.has_mano_pickles(mano_path):
            mano_state = assets.mano_state_from_dir(mano_path)
            super().load_state_dict(mano_state, strict=False)
            self._asset_keys = set(mano_state)
        self._packed = False
        self._workspace = {}
        self._graphs = {}
        self.eval()

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict, strict=True, allow_missing=False, **kw):
        """Same call as the reference (apps/eval.py:107-108 uses strict=False). Unlike nn.Module we do
        not let strict=False hide a key the kernels need: that raises unless allow_missing=True."""
        res = super().load_state_dict(state_dict, strict=strict, **kw)
        self._loaded_keys = None if allow_missing else set(state_dict.keys()) | self._asset_keys
        self._packed = False
        self._graphs.clear()
        return res

    def load_checkpoint(self, path, strict=False):
        """apps/eval.py:107-108 in one call: torch.load(path)['net'] (DataParallel prefixes stripped) -> load_state_dict.
        Returns the missing/unexpected key report."""
        state, report = assets.read_checkpoint(path, expected_keys=reference_key_shapes(self.backbone_name).keys())
        self.load_state_dict(state, strict=strict)
        return report

    def _apply(self, fn, *a, **k):
        self._packed = False
        self._graphs = {}
        self._workspace = {}  # belongs to the old device; the handle is re-created in _ensure_handle if the GPU changed
        self._copy_stream = None  # staging ring and copy stream belong to the old device
        self._staging = {}
        return super()._apply(fn, *a, **k)

    def _device(self):
        return self.init_regressor.offset.weight.device

    def _ensure_handle(self):
        dev = self._device()
        if dev.type != "cuda":
            raise capi.DirB200Error("DIR runs on an sm_100a CUDA device only (call .cuda()); there is no CPU fallback")
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._handle is not None and self._handle.device != index:  # the module moved to another GPU
            self._handle.close()
            self._handle = None
            self._packed = False
            self._workspace = {}
        if self._handle is None:
            self._handle = capi.Handle(self.precision, self.max_batch, self.aux_outputs, index, self.refine_stages,
                                       self.backbone_name)
        return self._handle

    def required_keys(self):
        return self._ensure_handle().required_keys()

    def _pack(self):
        h = self._ensure_handle()
        if self._loaded_keys is not None:  # coverage check that strict=False would have hidden
            missing = sorted(k for k in h.required_keys() if k not in self._loaded_keys)
            if missing:
                raise KeyError(f"the loaded state_dict lacks {len(missing)} keys the DIR forward needs, "
                               f"e.g. {missing[:4]} (pass allow_missing=True to load_state_dict to override)")
        keep = []
        with torch.cuda.device(self._device()):
            for name, t in self.state_dict().items():
                if t.dtype == torch.int64:
                    dtype = capi.DTYPE_I64
                else:
                    t = t.detach().to(torch.float32)
                    dtype = capi.DTYPE_F32
                t = t.contiguous()
                keep.append(t)
                h.set_weight(name, t.data_ptr(), dtype, list(t.shape))
            stream = torch.cuda.current_stream()
            h.finalize(stream.cuda_stream)
            stream.synchronize()
        del keep
        self._packed = True
        self._workspace.clear()

    # ------------------------------------------------------------------ forward
    def _workspace_for(self, B):
        if B not in self._workspace:
            n = self._handle.workspace_bytes(B)
            self._workspace[B] = torch.empty(n, dtype=torch.uint8, device=self._device())
        return self._workspace[B]

    def _alloc_outputs(self, B):
        dev = self._device()
        o = {"record": torch.empty(B, capi.RECORD_FLOATS, device=dev),
             "mano_para": torch.empty(B, 3, 2, 64, device=dev)}
        if self.aux_outputs:
            o["seg"] = torch.empty(B, 3, 32, 32, device=dev)
            o["dense"] = torch.empty(B, 3, 32, 32, device=dev)
            o["proj_feat"] = torch.empty(B, 1280, 32, 32, device=dev)
        return o

    def _enqueue(self, x, o):
        h = self._handle
        B = x.shape[0]
        ws = self._workspace_for(B)
        outs = capi.Outputs(o["record"].data_ptr(), o["mano_para"].data_ptr(),
                            o["seg"].data_ptr() if self.aux_outputs else None,
                            o["dense"].data_ptr() if self.aux_outputs else None,
                            o["proj_feat"].data_ptr() if self.aux_outputs else None)
        fwd = h.lib.dirb200_forward_u8 if x.dtype == torch.uint8 else h.lib.dirb200_forward
        rc = fwd(h.h, C.c_void_p(x.data_ptr()), B, C.c_void_p(ws.data_ptr()), ws.numel(),
                 C.byref(outs), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        h.check(rc, "dirb200_forward")

    def _run_chunk(self, x):
        B = x.shape[0]
        if not self.use_cuda_graph:
            o = self._alloc_outputs(B)
            self._enqueue(x, o)
            return o
        key = (B, x.dtype)
        if key not in self._graphs:
            # two complete (graph, static input, static outputs) sets used alternately: the outputs of call i stay valid
            # until call i+2, so nothing has to be cloned out of the graph's buffers (proj_feat alone is 671 MB at B=128)
            sets = []
            for _ in range(2):
                sx = torch.empty_like(x)
                so = self._alloc_outputs(B)
                sx.copy_(x)
                self._enqueue(sx, so)  # warm-up outside capture (sets function attributes)
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue(sx, so)
                sets.append((g, sx, so))
            self._graphs[key] = {"sets": sets, "next": 0}
        entry = self._graphs[key]
        g, sx, so = entry["sets"][entry["next"]]
        entry["next"] ^= 1
        sx.copy_(x)
        g.replay()
        return so

    def _to_device(self, img):
        """models/dir.py:514 does `input['img'].cuda()` on the compute stream. Host tensors are uploaded on a
        dedicated copy stream into a two-slot staging ring owned by the module (no allocator traffic, pinned
        memory => truly asynchronous), so the H2D copy of call i+1 overlaps the kernels of call i; the compute
        stream waits on an event, never the host. Returns (device tensor, ring slot or None)."""
        dev = self._device()
        dt = torch.uint8 if img.dtype == torch.uint8 else torch.float32
        if img.device.type == "cuda":
            return img.to(device=dev, dtype=dt).contiguous(), None
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staging = {}
        key = (tuple(img.shape), dt)
        ring = self._staging.get(key)
        if ring is None:
            if len(self._staging) >= 4:  # shapes seen long ago: let the allocator have the memory back
                self._staging.clear()
            ring = self._staging[key] = {"buf": [torch.empty(img.shape, dtype=dt, device=dev) for _ in range(2)],
                                         "free": [None, None], "next": 0}
        slot = ring["next"]
        ring["next"] = slot ^ 1
        buf = ring["buf"][slot]
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(self._copy_stream):
            if ring["free"][slot] is not None:  # the forward that last read this slot must be done with it
                self._copy_stream.wait_event(ring["free"][slot])
            buf.copy_(img, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        cur.wait_event(ev)
        return buf, (ring, slot)

    def _release_staging(self, token):
        if token is not None:
            ring, slot = token
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self._device()))
            ring["free"][slot] = ev

    def run_raw(self, img):
        """img (B,3,256,256) on the module's device -> dict of packed output buffers (record, mano_para, ...)."""
        self._ensure_handle()
        if not self._packed:
            self._pack()
        u8 = img.dtype == torch.uint8
        if img.dim() != 4 or tuple(img.shape[1:]) != ((256, 256, 3) if u8 else (3, 256, 256)):
            raise ValueError("expected (B,3,256,256) float images (apps/eval.py:50-61) or raw (B,256,256,3) uint8 BGR "
                             f"frames, got {tuple(img.shape)} {img.dtype}")
        with torch.cuda.device(self._device()):
            x, token = self._to_device(img)
            try:
                if x.shape[0] <= self.max_batch:
                    return self._run_chunk(x)
                parts = [self._run_chunk(x[i:i + self.max_batch]) for i in range(0, x.shape[0], self.max_batch)]
                return {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}
            finally:
                self._release_staging(token)

    @staticmethod
    def unpack_record(record, aux=None):
        """(B, 3*4887) packed record -> the reference's outs_list (views, no copies)."""
        B = record.shape[0]
        outs = []
        for i in range(3):
            d = {}
            for key, off, shp in STAGE_KEYS:
                a = i * capi.STAGE_FLOATS + capi.OFF[off]
                n = 1
                for s in shp:
                    n *= s
                d[key] = record[:, a:a + n].unflatten(1, shp) if len(shp) > 1 else record[:, a:a + n]
            d["pd_rel_joint"] = None
            outs.append(d)
        if aux is not None:
            outs.append(aux)
        return outs

    def forward(self, input, target=None, meta_info=None):
        if self.training:
            raise NotImplementedError("only the eval branch of DIR.forward (models/dir.py:513-540) is built; "
                                      "call .eval()")
        o = self.run_raw(input["img"])
        aux = {"dense": o["dense"], "seg": o["seg"], "proj_feat": o["proj_feat"]} if self.aux_outputs else \
              {"dense": None, "seg": None, "proj_feat": None}
        outs = self.unpack_record(o["record"], aux)
        if self.refine_stages == 1:
            outs = outs[:2] + outs[3:]
        return outs, {}

    # ------------------------------------------------------------------ multi-GPU (batch sharding + one all-gather)
    def init_nccl(self, rank, world, broadcast_bytes):
        """Create the library's NCCL communicator. `broadcast_bytes(b: bytes|None) -> bytes` must return rank 0's
        128-byte id on every rank (e.g. via torch.distributed.broadcast_object_list)."""
        h = self._ensure_handle()
        buf = C.create_string_buffer(128)
        if rank == 0:
            h.check(h.lib.dirb200_nccl_unique_id(h.h, buf), "nccl_unique_id")
        uid = broadcast_bytes(bytes(buf.raw) if rank == 0 else None)
        with torch.cuda.device(self._device()):
            h.check(h.lib.dirb200_nccl_init(h.h, uid, rank, world), "nccl_init")
        self._world = world

    def allgather_records(self, record):
        h = self._handle
        B = record.shape[0]
        out = torch.empty(self._world * B, capi.RECORD_FLOATS, device=record.device)
        rc = h.lib.dirb200_allgather_records(h.h, C.c_void_p(record.data_ptr()), C.c_void_p(out.data_ptr()), B,
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
        h.check(rc, "allgather_records")
        return out
