"""Host-side mirror of the reference model interface for the dense-inference hot path.

`ConvolutionalOccupancyNetwork` here has the reference class's public surface
(/root/reference/src/vgn/ConvONets/conv_onet/models/__init__.py:15-164): the same
parameter names/shapes (so reference checkpoints load and ours load into the
reference), the same methods (`forward/encode_inputs/decode/decode_occ/infer_geo/
query_feature/decode_feature/to`) and attributes (`encoder`, `decoder_qual/rot/width/
tsdf`, `_device`, `detach_tsdf`).  The sub-modules are parameter containers only: all
arithmetic happens in libgiga_b200.so (hand-written sm_100a CUDA) through the C ABI in
include/giga_b200.h.  PyTorch provides device memory, streams and the nn.Module
plumbing -- nothing on the compute path.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch
import torch.nn as nn
from torch import distributions as dist

from . import _lib
import weakref

from ._lib import HEAD_GRASP, HEAD_QUAL, HEAD_RAW, HEAD_ROT, HEAD_TSDF, HEAD_WIDTH, check, lib

_STRUCT_VERSION = [0]   # bumped whenever any weight/bias attribute is (re)bound
PLANES = ("xz", "xy", "yz")
GRID = 40
CDIM = 32


# --------------------------------------------------------------------------------------
# parameter containers (names/shapes/initialisation of the reference modules)
# --------------------------------------------------------------------------------------
class _Affine(nn.Module):
    """weight + bias holder with torch's default (kaiming-uniform a=sqrt(5)) initialisation
    of nn.Linear / nn.ConvNd / nn.ConvTranspose2d."""

    def __init__(self, *shape, fan_in: int, bias_len: Optional[int] = None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*shape))
        self.bias = nn.Parameter(torch.empty(shape[0] if bias_len is None else bias_len))
        bound = 1.0 / math.sqrt(fan_in)
        nn.init.uniform_(self.weight, -bound, bound)  # kaiming_uniform_(a=sqrt(5)) == U(+-1/sqrt(fan_in))
        nn.init.uniform_(self.bias, -bound, bound)

    def __setattr__(self, name, value):   # a re-bound Parameter object invalidates the engine's cached tensor list
        _STRUCT_VERSION[0] += 1
        super().__setattr__(name, value)


def _linear(n_in, n_out):
    return _Affine(n_out, n_in, fan_in=n_in)


def _conv3x3(c_in, c_out):
    m = _Affine(c_out, c_in, 3, 3, fan_in=c_in * 9)
    nn.init.xavier_normal_(m.weight)  # encoder/unet.py:213-217 (UNet.weight_init, Conv2d only)
    nn.init.constant_(m.bias, 0)
    return m


class DownConv(nn.Module):  # encoder/unet.py:48-72
    def __init__(self, c_in, c_out):
        super().__init__()
        self.conv1 = _conv3x3(c_in, c_out)
        self.conv2 = _conv3x3(c_out, c_out)


class UpConv(nn.Module):  # encoder/unet.py:75-114 (transpose up-conv, concat merge)
    def __init__(self, c_in, c_out):
        super().__init__()
        # ConvTranspose2d weight is [in, out, 2, 2]; torch computes fan_in from size(1)
        self.upconv = _Affine(c_in, c_out, 2, 2, fan_in=c_out * 4, bias_len=c_out)
        self.conv1 = _conv3x3(2 * c_out, c_out)
        self.conv2 = _conv3x3(c_out, c_out)


class UNet(nn.Module):  # encoder/unet.py:117-239: UNet(32, in_channels=32, depth=3, start_filts=32, concat)
    def __init__(self):
        super().__init__()
        self.down_convs = nn.ModuleList([DownConv(32, 32), DownConv(32, 64), DownConv(64, 128)])
        self.up_convs = nn.ModuleList([UpConv(128, 64), UpConv(64, 32)])
        self.conv_final = _Affine(32, 32, 1, 1, fan_in=32)
        nn.init.xavier_normal_(self.conv_final.weight)
        nn.init.constant_(self.conv_final.bias, 0)


class LocalVoxelEncoder(nn.Module):  # encoder/voxels.py:10-121
    def __init__(self):
        super().__init__()
        self.conv_in = _Affine(32, 1, 3, 3, 3, fan_in=27)
        self.unet = UNet()
        self.c_dim = CDIM
        self.reso_plane = GRID
        self.plane_type = list(PLANES)
        self.padding = 0


class ResnetBlockFC(nn.Module):  # layers.py:6-47
    def __init__(self, size):
        super().__init__()
        self.fc_0 = _linear(size, size)
        self.fc_1 = _linear(size, size)
        nn.init.zeros_(self.fc_1.weight)


class LocalDecoder(nn.Module):  # conv_onet/models/decoder.py:61-206 (concat_feat=True -> c_dim 96)
    def __init__(self, out_dim: int):
        super().__init__()
        self.c_dim = 3 * CDIM
        self.n_blocks = 5
        self.hidden_size = 32
        self.out_dim = out_dim
        self.fc_c = nn.ModuleList([_linear(96, 32) for _ in range(5)])
        self.fc_p = _linear(3, 32)
        self.blocks = nn.ModuleList([ResnetBlockFC(32) for _ in range(5)])
        self.fc_out = _linear(32, out_dim)
        self.__dict__["_owner"] = None   # weakref to the model whose engine evaluates this head
        self.__dict__["_head_bit"] = 0

    def __getstate__(self):   # the owner back-reference is rebound by the owning model (copy.deepcopy / pickle)
        d = self.__dict__.copy()
        d["_owner"] = None
        return d

    def forward(self, p, c_plane, **kwargs):
        """decoder.py:133-176 -- the bare head output (no sigmoid / normalise): (B,N) or (B,N,4)."""
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise _lib.GigaError("LocalDecoder is a parameter container; call it through its ConvolutionalOccupancyNetwork")
        out = owner._decode_heads(p, c_plane, self._head_bit | HEAD_RAW)
        return out[{HEAD_QUAL: 0, HEAD_ROT: 1, HEAD_WIDTH: 2, HEAD_TSDF: 3}[self._head_bit]]


class PlaneFeatures(dict):
    """The dict `encode_inputs` returns ({'xz','xy','yz'} -> (B,32,40,40) views) plus the packed
    channels-last buffer [3][B][40][40][32] the views alias (what the decoder kernels read)."""

    packed: Optional[torch.Tensor] = None


# --------------------------------------------------------------------------------------
# engine: one giga_ctx per (module, device)
# --------------------------------------------------------------------------------------
class _Engine:
    def __init__(self, device: torch.device):
        if device.type != "cuda":
            raise _lib.GigaError("giga_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        self.device = device
        h = C.c_void_p()
        check(lib.giga_ctx_create(C.byref(h), device.index if device.index is not None else torch.cuda.current_device()),
              "giga_ctx_create")
        self.h = h
        self.param_key = None
        self.param_names = None
        self.tensors = []
        self.struct_version = -1
        self.commit_mode = "auto"      # "host": always pack on the host (A/B testing of the device-side packer)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib.giga_ctx_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def invalidate(self):
        """Forget the committed parameters: the next call re-packs and re-uploads all of them."""
        self.param_key = None
        self.struct_version = -1

    def sync_params(self, module: nn.Module):
        # fast path (every call): the cached tensor objects still hold the committed storage and version.
        # NOTE: change detection is (storage pointer, tensor._version); in-place writes through `.data` / `.detach()` views
        # made BEFORE the write do not bump the version counter -- call net.invalidate_params() after such writes.
        if self.struct_version == _STRUCT_VERSION[0] and self.param_key == tuple((v.data_ptr(), v._version) for v in self.tensors):
            return
        params = list(module.state_dict(keep_vars=True).items())
        self.tensors = [v for _, v in params]
        self.struct_version = _STRUCT_VERSION[0]
        key = tuple((v.data_ptr(), v._version) for v in self.tensors)
        names = tuple(k for k, _ in params)
        if key == self.param_key and names == self.param_names:
            return
        self.param_names = names
        # parameters that already live on this device (the normal case) are packed BY KERNELS straight from the live tensors
        # (giga_ctx_commit_device: no device -> host -> pack -> device round trip); anything else takes the host packer below
        if self.commit_mode != "host" and "encoder.conv_in.weight" in names and all(
                v.is_cuda and v.device == self.device and v.dtype == torch.float32 and v.is_contiguous() and v.data_ptr() % 16 == 0 for v in self.tensors):
            n = len(params)
            c_names = (C.c_char_p * n)(*[k.encode() for k, _ in params])
            vals = (C.c_void_p * n)(*[v.data_ptr() for v in self.tensors])
            check(lib.giga_ctx_commit_device(self.h, n, c_names, vals, _stream(self.device)), "giga_ctx_commit_device")
            self.__dict__.pop("_train_key", None)      # the commit re-bound the value pointers: the training step binds its gradients again
            self.param_key = key
            return
        # one flattened upload (a training loop re-commits after every optimizer.step(): 164 separate copies cost 20 ms)
        with torch.no_grad():
            flat = torch.cat([v.detach().reshape(-1).float() for _, v in params]).contiguous()
        n = len(params)
        c_names = (C.c_char_p * n)(*[k.encode() for k, _ in params])
        numels = [v.numel() for _, v in params]
        offs, acc = [], 0
        for m in numels:
            offs.append(acc)
            acc += m
        check(lib.giga_ctx_set_params_flat(self.h, n, c_names, (C.c_long * n)(*offs), (C.c_long * n)(*numels), C.c_void_p(flat.data_ptr()),
                                           flat.numel(), int(flat.is_cuda)), "giga_ctx_set_params_flat")
        check(lib.giga_ctx_commit_params(self.h), "giga_ctx_commit_params")
        self.param_key = key

    @property
    def launches(self) -> int:
        return int(lib.giga_ctx_launch_count(self.h))

    def set_option(self, key: str, value: int):
        check(lib.giga_ctx_set_option(self.h, key.encode(), int(value)), f"giga_ctx_set_option({key})")

    def set_timing(self, on: bool):
        check(lib.giga_ctx_set_timing(self.h, int(on)), "giga_ctx_set_timing")

    def timing_report(self):
        """-> {kernel name: (launches, total_ms)} since the last report (CUDA events on the launch stream)."""
        buf = C.create_string_buffer(1 << 16)
        check(int(lib.giga_ctx_timing_report(self.h, buf, len(buf))), "giga_ctx_timing_report")
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.rsplit(" ", 2)
            out[name] = (int(n), float(ms))
        return out


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _prep(t: torch.Tensor, device) -> torch.Tensor:
    if t.device != device:
        raise _lib.GigaError(f"input on {t.device}, model on {device}")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _GigaBase(nn.Module):
    """Shared host logic of the two reference wrappers."""

    _device = None

    # -- engine plumbing ---------------------------------------------------------------
    def _engine(self) -> _Engine:
        dev = next(self.parameters()).device
        eng = self.__dict__.get("_eng")
        if eng is None or eng.device != dev:
            eng = _Engine(dev)
            self.__dict__["_eng"] = eng
        eng.sync_params(self)
        return eng

    def _engine_raw(self) -> _Engine:
        """the engine WITHOUT committing the parameters to the inference operand blobs (the native training step reads the live
        parameter tensors on the device; the next inference call re-commits because the optimizer bumps the version counters)"""
        dev = next(self.parameters()).device
        eng = self.__dict__.get("_eng")
        if eng is None or eng.device != dev:
            eng = _Engine(dev)
            self.__dict__["_eng"] = eng
        return eng

    def _train_active(self, *inputs) -> bool:
        """autograd semantics of an ordinary nn.Module: with gradient mode on and parameters (or query positions) that require
        gradients the outputs are differentiable -- through the native training step (csrc/train_bwd.cuh)."""
        if not torch.is_grad_enabled():
            return False
        return any(t is not None and t.requires_grad for t in inputs) or any(p.requires_grad for p in self.parameters())

    def to(self, device):
        """models/__init__.py:126-134"""
        model = super().to(device)
        model._device = device
        return model

    def invalidate_params(self):
        """Force a re-commit of all parameters on the next call.  Parameter changes are detected through
        (storage pointer, tensor._version): `optimizer.step()`, `load_state_dict`, `p.add_()`, `p.copy_()` are all seen.
        Writes that bypass the version counter -- `p.data.copy_(ema)`, `p.data.clamp_()`, writes through a view taken with
        `.detach()` -- are NOT; call this method after them (or the engine keeps evaluating the previously committed weights)."""
        eng = self.__dict__.get("_eng")
        if eng is not None:
            eng.invalidate()
        return self

    def overflow_count(self, reset: bool = True) -> int:
        """Number of non-finite head outputs the tensor-core decoder has produced since the last reset (an activation or plane
        feature beyond fp16's +-65504 operand range turns the affected outputs into NaN; giga_ctx_overflow_count).  Synchronises."""
        eng = self.__dict__.get("_eng")
        return int(check(int(lib.giga_ctx_overflow_count(eng.h, int(reset))), "giga_ctx_overflow_count")) if eng else 0

    # copy.deepcopy / pickle / torch.save(net): the ctypes engine is per-object device state, never copied; the copy builds its
    # own on first use and its heads are re-bound to it (a copied head must not keep evaluating the original model's weights)
    def __getstate__(self):
        d = self.__dict__.copy()
        for k in ("_eng", "_inflight"):
            d.pop(k, None)
        return d

    def __setstate__(self, state):
        super().__setstate__(state)
        self._bind_heads()

    @property
    def gpu_launches(self) -> int:
        eng = self.__dict__.get("_eng")
        return eng.launches if eng else 0

    # -- encoder -----------------------------------------------------------------------
    def encode_inputs(self, inputs: torch.Tensor) -> PlaneFeatures:
        """models/__init__.py:74-87 -> LocalVoxelEncoder.forward (encoder/voxels.py:89-121)."""
        eng = self._engine()
        x = _prep(inputs, eng.device)
        if x.dim() != 4 or tuple(x.shape[1:]) != (GRID, GRID, GRID):
            raise _lib.GigaError(f"inputs must be (B,{GRID},{GRID},{GRID}), got {tuple(inputs.shape)}")
        B = x.shape[0]
        packed = torch.empty((3, B, GRID, GRID, CDIM), device=eng.device, dtype=torch.float32)
        check(lib.giga_encode(eng.h, C.c_void_p(x.data_ptr()), B, C.c_void_p(packed.data_ptr()), _stream(eng.device)),
              "giga_encode")
        c = PlaneFeatures((k, packed[i].permute(0, 3, 1, 2)) for i, k in enumerate(PLANES))
        c.packed = packed
        return c

    @staticmethod
    def _packed(c: Dict[str, torch.Tensor]) -> torch.Tensor:
        packed = getattr(c, "packed", None)
        if packed is not None and all(c[k].data_ptr() == packed[i].data_ptr() for i, k in enumerate(PLANES)):
            return packed
        # foreign plane tensors (e.g. produced elsewhere): re-pack to channels-last (layout only)
        return torch.stack([c[k].float().permute(0, 2, 3, 1) for k in PLANES]).contiguous()

    # -- decoder -----------------------------------------------------------------------
    def _decode_heads(self, p: torch.Tensor, c, heads: int):
        eng = self._engine()
        planes = self._packed(c)
        pts = _prep(p, eng.device)
        if pts.dim() != 3 or pts.shape[2] != 3 or pts.shape[0] != planes.shape[1]:
            raise _lib.GigaError(f"points must be (B,N,3) with B={planes.shape[1]}, got {tuple(p.shape)}")
        B, N = pts.shape[0], pts.shape[1]
        mk = lambda *s: torch.empty(s, device=eng.device, dtype=torch.float32)
        if heads & 15 & ~int(lib.giga_ctx_heads(eng.h)):
            raise _lib.GigaError("this model has no parameters for a requested decoder head")
        qual = mk(B, N) if heads & HEAD_QUAL else None
        rot = mk(B, N, 4) if heads & HEAD_ROT else None
        width = mk(B, N) if heads & HEAD_WIDTH else None
        occ = mk(B, N) if heads & HEAD_TSDF else None
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        check(lib.giga_decode(eng.h, ptr(planes), B, ptr(pts), N, heads, ptr(qual), ptr(rot), ptr(width), ptr(occ),
                              _stream(eng.device)), "giga_decode")
        return qual, rot, width, occ

    def _bind_heads(self):
        for name, bit in (("decoder_qual", HEAD_QUAL), ("decoder_rot", HEAD_ROT), ("decoder_width", HEAD_WIDTH),
                          ("decoder_tsdf", HEAD_TSDF)):
            dec = self._modules.get(name)
            if dec is not None:
                dec.__dict__["_owner"] = weakref.ref(self)
                dec.__dict__["_head_bit"] = bit

    def decode_occ(self, p, c, **kwargs):
        """models/__init__.py:100-109"""
        logits = self._decode_heads(p, c, HEAD_TSDF)[3]
        return dist.Bernoulli(logits=logits)

    def infer_geo(self, inputs, p_tsdf, **kwargs):
        """models/__init__.py:69-72"""
        c = self.encode_inputs(inputs)
        return self._decode_heads(p_tsdf, c, HEAD_TSDF)[3]

    def sample_feature(self, p, c, mode: str = "concat") -> torch.Tensor:
        eng = self._engine()
        planes = self._packed(c)
        pts = _prep(p, eng.device)
        B, N = pts.shape[0], pts.shape[1]
        out = torch.empty((B, N, 96 if mode == "concat" else 32), device=eng.device, dtype=torch.float32)
        check(lib.giga_sample_feature(eng.h, C.c_void_p(planes.data_ptr()), B, C.c_void_p(pts.data_ptr()), N,
                                      0 if mode == "concat" else 1, C.c_void_p(out.data_ptr()), _stream(eng.device)),
              "giga_sample_feature")
        return out

    def scene_argmax(self, qual: torch.Tensor, out_val: Optional[torch.Tensor] = None, out_idx: Optional[torch.Tensor] = None):
        """Per-scene (max quality, first arg-max) -- the final grasp-score reduction; `out_*` may be
        slices of an all-gather buffer."""
        eng = self._engine()
        q = _prep(qual, eng.device)
        B, N = q.shape
        if out_val is None:
            out_val = torch.empty(B, device=eng.device, dtype=torch.float32)
        if out_idx is None:
            out_idx = torch.empty(B, device=eng.device, dtype=torch.int32)
        assert out_val.is_contiguous() and out_idx.is_contiguous() and out_idx.dtype == torch.int32
        check(lib.giga_scene_argmax(eng.h, C.c_void_p(q.data_ptr()), B, N, C.c_void_p(out_val.data_ptr()),
                                    C.c_void_p(out_idx.data_ptr()), _stream(eng.device)), "giga_scene_argmax")
        return out_val, out_idx

    def select_grasps(self, tsdf, qual, rot, width, params=None, K: int = 512, return_qual_vol: bool = False):
        """process() + bound() + select() of the reference planner (detection_implicit.py:115-143, 87-97, 146-174) on
        DEVICE volumes of B scenes (giga_select_grasps): tsdf (B,40,40,40), qual/width (B,64000), rot (B,64000,4) ->
        count (B,) int32, score (B,K), index (B,K) int32, rot (B,K,4), width (B,K) [, processed quality volume (B,40,40,40)]."""
        from .detection_implicit import select_params
        eng = self._engine()
        t, q, r, w = (_prep(a, eng.device) for a in (tsdf, qual, rot, width))
        B = t.shape[0]
        if tuple(t.shape) != (B, GRID, GRID, GRID) or q.numel() != B * GRID ** 3 or r.numel() != 4 * B * GRID ** 3 or w.numel() != B * GRID ** 3:
            raise _lib.GigaError("select_grasps takes 40^3 volumes: tsdf (B,40,40,40), qual/width (B,64000), rot (B,64000,4)")
        params = params or select_params()
        mk = lambda shape, dt: torch.zeros(shape, device=eng.device, dtype=dt)
        count, score, index = mk((B,), torch.int32), mk((B, K), torch.float32), mk((B, K), torch.int32)
        orot, owidth = mk((B, K, 4), torch.float32), mk((B, K), torch.float32)
        qvol = mk((B, GRID, GRID, GRID), torch.float32) if return_qual_vol else None
        ptr = lambda a: C.c_void_p(a.data_ptr()) if a is not None else C.c_void_p(0)
        check(lib.giga_select_grasps(eng.h, ptr(t), ptr(q), ptr(r), ptr(w), B, C.byref(params), K, ptr(count), ptr(score), ptr(index),
                                     ptr(orot), ptr(owidth), ptr(qvol), _stream(eng.device)), "giga_select_grasps")
        out = (count, score, index, orot, owidth)
        return out + (qvol,) if return_qual_vol else out

    def debug_activation(self, name: str, B: int) -> torch.Tensor:
        """Copy an intermediate activation of the last encode (tests only)."""
        eng = self._engine()
        shapes = {"pre": (32, 40), "d0c1": (32, 40), "d0c2": (32, 40), "p0": (32, 20), "d1c1": (64, 20), "d1c2": (64, 20),
                  "p1": (64, 10), "d2c1": (128, 10), "d2c2": (128, 10), "u0": (64, 20), "u0c1": (64, 20), "u0c2": (64, 20),
                  "u1": (32, 40), "u1c1": (32, 40), "u1c2": (32, 40)}
        ch, hw = shapes[name]
        out = torch.empty((3, B, ch, hw, hw), device=eng.device, dtype=torch.float32)
        n = lib.giga_debug_copy(eng.h, name.encode(), C.c_void_p(out.data_ptr()), out.numel(), _stream(eng.device))
        check(int(n), "giga_debug_copy")
        assert n == out.numel()
        return out

    def forward_host(self, tsdf: torch.Tensor, p: Optional[torch.Tensor], p_tsdf: Optional[torch.Tensor] = None, out=None):
        """End-to-end call with HOST tensors (pinned for full PCIe bandwidth): H2D, encode, decode,
        D2H inside one C-ABI call (giga_forward_host).  Mirrors `predict()`
        (detection_implicit.py:99-113).  Returns host tensors (qual, rot, width[, occ])."""
        eng = self._engine()
        for t in (tsdf, p, p_tsdf):
            if t is not None and (t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
                raise _lib.GigaError("forward_host takes contiguous fp32 host tensors")
        B = tsdf.shape[0]
        Ng = p.shape[1] if p is not None else 0
        No = p_tsdf.shape[1] if p_tsdf is not None else 0
        if out is None:
            pin = dict(dtype=torch.float32, pin_memory=True)
            out = (torch.empty((B, Ng), **pin), torch.empty((B, Ng, 4), **pin), torch.empty((B, Ng), **pin),
                   torch.empty((B, No), **pin))
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None and t.numel() else C.c_void_p(0)
        check(lib.giga_forward_host(eng.h, ptr(tsdf), B, ptr(p), Ng, ptr(p_tsdf), No, ptr(out[0]), ptr(out[1]), ptr(out[2]),
                                    ptr(out[3]), _stream(eng.device)), "giga_forward_host")
        return out if No else out[:3]


    # pipelined host path (serving loops): submit to slot 0/1, wait later; copies overlap the other slot's kernels
    def forward_host_submit(self, slot: int, tsdf: torch.Tensor, p: Optional[torch.Tensor], p_tsdf: Optional[torch.Tensor], out):
        eng = self._engine()
        for t in (tsdf, p, p_tsdf) + tuple(out):
            if t is not None and t.numel() and (t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or not t.is_pinned()):
                raise _lib.GigaError("forward_host_submit takes contiguous pinned fp32 host tensors")
        B = tsdf.shape[0]
        Ng = p.shape[1] if p is not None else 0
        No = p_tsdf.shape[1] if p_tsdf is not None else 0
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None and t.numel() else C.c_void_p(0)
        check(lib.giga_forward_host_submit(eng.h, slot, ptr(tsdf), B, ptr(p), Ng, ptr(p_tsdf), No, ptr(out[0]), ptr(out[1]),
                                           ptr(out[2]), ptr(out[3])), "giga_forward_host_submit")
        self.__dict__.setdefault("_inflight", {})[slot] = (tsdf, p, p_tsdf, out)   # keep the host buffers alive

    def forward_host_wait(self, slot: int):
        eng = self._engine()
        check(lib.giga_forward_host_wait(eng.h, slot), "giga_forward_host_wait")
        return self.__dict__.get("_inflight", {}).pop(slot, (None, None, None, None))[3]


class ConvolutionalOccupancyNetwork(_GigaBase):
    """Drop-in for conv_onet/models/__init__.py:15-164 (giga, giga_aff, giga_detach)."""

    def __init__(self, with_tsdf: bool = True, device=None, detach_tsdf: bool = False):
        super().__init__()
        self.decoder_qual = LocalDecoder(1)
        self.decoder_rot = LocalDecoder(4)
        self.decoder_width = LocalDecoder(1)
        if with_tsdf:
            self.decoder_tsdf = LocalDecoder(1)
        self.encoder = LocalVoxelEncoder()
        self._device = device
        self.detach_tsdf = detach_tsdf
        self._bind_heads()
        if device is not None:
            self.to(device)

    def _forward_native(self, inputs, p, p_tsdf=None, best=None):
        """One C-ABI call (giga_forward): encode + grasp heads at p + TSDF head at p_tsdf [+ per-scene arg-max into
        best = (val f32[B], idx i32[B])]; the plane features stay in a library-owned buffer."""
        eng = self._engine()
        x = _prep(inputs, eng.device)
        if x.dim() != 4 or tuple(x.shape[1:]) != (GRID, GRID, GRID):
            raise _lib.GigaError(f"inputs must be (B,{GRID},{GRID},{GRID}), got {tuple(inputs.shape)}")
        B = x.shape[0]
        pts = _prep(p, eng.device)
        if pts.dim() != 3 or pts.shape[2] != 3 or pts.shape[0] != B:
            raise _lib.GigaError(f"points must be (B,N,3) with B={B}, got {tuple(p.shape)}")
        N = pts.shape[1]
        mk = lambda *s: torch.empty(s, device=eng.device, dtype=torch.float32)
        qual, rot, width = mk(B, N), mk(B, N, 4), mk(B, N)
        pt = occ = None
        No = 0
        if p_tsdf is not None:
            if not hasattr(self, "decoder_tsdf"):
                raise AttributeError("this model has no decoder_tsdf")   # as the reference (models/__init__.py:64)
            pt = _prep(p_tsdf, eng.device)
            if pt.dim() != 3 or pt.shape[2] != 3 or pt.shape[0] != B:
                raise _lib.GigaError(f"p_tsdf must be (B,N,3) with B={B}, got {tuple(p_tsdf.shape)}")
            No = pt.shape[1]
            occ = mk(B, No)
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        bv, bi = best if best is not None else (None, None)
        check(lib.giga_forward(eng.h, ptr(x), B, ptr(pts), N, ptr(pt), No, C.c_void_p(0), ptr(qual), ptr(rot), ptr(width), ptr(occ),
                               ptr(bv), ptr(bi), _stream(eng.device)), "giga_forward")
        if p_tsdf is not None:
            return qual, rot, width, occ
        return qual, rot, width

    def forward_with_argmax(self, inputs, p, p_tsdf=None, out_val=None, out_idx=None):
        """forward() plus the final grasp-score reduction (per-scene max quality, first arg-max) in the same call;
        out_val / out_idx may be this rank's slices of the all-gather buffers."""
        dev = next(self.parameters()).device
        B = inputs.shape[0]
        if out_val is None:
            out_val = torch.empty(B, device=dev, dtype=torch.float32)
        if out_idx is None:
            out_idx = torch.empty(B, device=dev, dtype=torch.int32)
        assert out_val.is_contiguous() and out_idx.is_contiguous() and out_idx.dtype == torch.int32
        return self._forward_native(inputs, p, p_tsdf, best=(out_val, out_idx)), (out_val, out_idx)

    def forward(self, inputs, p, p_tsdf=None, sample=True, **kwargs):
        """models/__init__.py:42-67"""
        if self._train_active(p, p_tsdf):
            from .training import native_forward
            return native_forward(self, inputs, p, p_tsdf)
        return self._forward_native(inputs, p, p_tsdf)

    def decode(self, p, c, **kwargs):
        """models/__init__.py:111-124 (sigmoid / normalise fused into the kernel epilogue)."""
        qual, rot, width, _ = self._decode_heads(p, c, HEAD_GRASP)
        return qual, rot, width

    def query_feature(self, p, c):
        """models/__init__.py:89-90 -> LocalDecoder.query_feature (decoder.py:178-191): summed 32-d feature."""
        return self.sample_feature(p, c, mode="sum")

    def decode_feature(self, p, feature):
        """models/__init__.py:92-98.  In the reference this feeds the 32-d *summed* feature into fc_c
        layers that expect 96 inputs (concat_feat=True), i.e. it raises for every shipped GIGA config
        and is commented out of forward (:57-58); mirrored as an explicit error."""
        raise RuntimeError("decode_feature: size mismatch, fc_c expects 96-d features (concat_feat=True); "
                           "same failure as the reference for GIGA configs")

    def grad_refine(self, x, pos, bound_value=0.0125, lr=1e-6, num_step=1):
        """models/__init__.py:136-164: SGD on the query positions to raise the predicted quality (not called by any shipped script).
        d(qual)/d(pos) comes from the native backward kernels (fc_p and grid_sampler's grid gradient, csrc/train_bwd.cuh)."""
        pos_tmp = pos.clone()
        l_bound = pos - bound_value
        u_bound = pos + bound_value
        pos_tmp.requires_grad = True
        optimizer = torch.optim.SGD([pos_tmp], lr=lr)
        self.eval()
        with torch.enable_grad():
            for _ in range(num_step):
                optimizer.zero_grad()
                qual_out, _, _ = self.forward(x, pos_tmp)
                loss = -qual_out.sum()
                loss.backward()
                optimizer.step()
        for prm in self.parameters():   # the reference leaves parameter gradients behind as well; drop ours
            prm.grad = None
        with torch.no_grad():
            pos_tmp = torch.maximum(torch.minimum(pos_tmp, u_bound), l_bound)
            qual_out, rot_out, width_out = self.forward(x, pos_tmp)
        return qual_out, pos_tmp, rot_out, width_out


class ConvolutionalOccupancyNetworkGeometry(_GigaBase):
    """Drop-in for conv_onet/models/__init__.py:166-226 (giga_geo: encoder + decoder_tsdf only)."""

    def __init__(self, device=None):
        super().__init__()
        self.decoder_tsdf = LocalDecoder(1)
        self.encoder = LocalVoxelEncoder()
        self._device = device
        self._bind_heads()
        if device is not None:
            self.to(device)

    def _forward_native(self, inputs, p, p_tsdf=None):
        return (self.infer_geo(inputs, p_tsdf),)

    def forward(self, inputs, p, p_tsdf, sample=True, **kwargs):
        """models/__init__.py:179-195"""
        if self._train_active(p_tsdf):
            from .training import native_forward
            return native_forward(self, inputs, None, p_tsdf)[0]
        return self.infer_geo(inputs, p_tsdf)
