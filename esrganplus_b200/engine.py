"""ctypes driver of the native RRDBNet engine (esrp_rrdbnet_* in include/esrp.h).

PyTorch is plumbing here: it owns the fp32 Parameters (the state_dict wire format), the workspace
allocation and the CUDA stream.  The packed bf16 weight cache inside the native handle is a derived
cache, rebuilt whenever a Parameter's storage pointer or in-place version counter changes
(load_state_dict, optimizer.step(), init_weights, .to(device)); it never appears in a state_dict.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch

from . import _lib


class GeneratorEngine:
    """One native handle per (module, device)."""

    def __init__(self, in_nc: int, out_nc: int, nf: int, nb: int, gc: int, upscale: int, device: torch.device):
        self.lib = _lib.load()
        self.device = device
        self.cfg = (in_nc, out_nc, nf, nb, gc, upscale)
        self.out_nc, self.upscale = out_nc, upscale
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib.esrp_rrdbnet_create(in_nc, out_nc, nf, nb, gc, upscale, C.byref(h)),
                       "esrp_rrdbnet_create")
        self.handle = h
        n = self.lib.esrp_rrdbnet_num_tensors(self.handle)
        self.keys: List[str] = [self.lib.esrp_rrdbnet_tensor_key(self.handle, i).decode() for i in range(n)]
        self._sig: Tuple = ()
        self._ws: Dict[Tuple[int, int, int], torch.Tensor] = {}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.esrp_rrdbnet_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- weights ---------------------------------------------------------------------------------
    def sync_weights(self, named_params: Dict[str, torch.Tensor], epoch: int = 0) -> bool:
        """Repack if any source tensor changed. `named_params`: state_dict-style key -> fp32 tensor
        on this device; `epoch`: the module's invalidation counter (writes through ``.data`` bump no tensor
        version, see architecture._NativeWeights). Returns True when a repack was launched."""
        cached = getattr(self, "_param_cache", None)
        if cached is not None and cached[0] is named_params:
            tensors = cached[1]   # same dict object as last time (callers cache it per module): the tensors were validated then
        else:
            tensors = []
            for k in self.keys:
                t = named_params.get(k)
                if t is None:
                    raise RuntimeError(f"RRDBNet engine: parameter {k!r} missing from the module")
                if t.device != self.device or t.dtype != torch.float32:
                    raise RuntimeError(f"RRDBNet engine: {k} must be fp32 on {self.device}, got {t.dtype} on {t.device}")
                tensors.append(t if t.is_contiguous() else t.contiguous())
            if all(t is named_params[k] for t, k in zip(tensors, self.keys)):   # (a contiguous() copy must be re-made every time)
                self._param_cache = (named_params, tensors)
        sig = (epoch, tuple([(t.data_ptr(), t._version) for t in tensors]))
        if sig == self._sig:
            return False
        ptrs = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        _lib.check(self.lib.esrp_rrdbnet_load_weights(self.handle, ptrs, len(tensors),
                                                      torch.cuda.current_stream(self.device).cuda_stream),
                   "esrp_rrdbnet_load_weights")
        self._sig = sig
        self._keep = tensors
        return True

    # -- forward ---------------------------------------------------------------------------------
    def workspace(self, n: int, h: int, w: int) -> torch.Tensor:
        key = (n, h, w)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = self.lib.esrp_rrdbnet_workspace_bytes(self.handle, n, h, w)
            if nbytes < 0:
                raise RuntimeError("esrp_rrdbnet_workspace_bytes failed")
            if len(self._ws) >= 4:  # keep the cache bounded: ragged test images come in many sizes
                self._ws.clear()
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def forward(self, x: torch.Tensor, training: bool, seed: int) -> torch.Tensor:
        if x.dim() != 4 or x.dtype != torch.float32 or x.device != self.device:
            raise RuntimeError(f"RRDBNet forward expects an fp32 NCHW tensor on {self.device}, got {x.dtype} {tuple(x.shape)} on {x.device}")
        x = x.contiguous()
        n, c, h, w = x.shape
        if c != self.cfg[0]:
            raise RuntimeError(f"RRDBNet forward: expected {self.cfg[0]} input channels, got {c}")
        y = torch.empty((n, self.out_nc, h * self.upscale, w * self.upscale), dtype=torch.float32, device=self.device)
        ws = self.workspace(n, h, w)
        base = ws.data_ptr()
        aligned = (base + 1023) // 1024 * 1024
        with torch.cuda.device(self.device):
            _lib.check(self.lib.esrp_rrdbnet_forward(self.handle, x.data_ptr(), y.data_ptr(), n, h, w,
                                                     aligned, ws.numel() - (aligned - base), int(bool(training)),
                                                     seed & 0xFFFFFFFFFFFFFFFF,
                                                     torch.cuda.current_stream(self.device).cuda_stream),
                       "esrp_rrdbnet_forward")
        return y

    def forward_u8(self, img: torch.Tensor, bgr: bool = True) -> torch.Tensor:
        """uint8 HWC images [n,h,w,in_nc] (cv2 order when bgr) -> uint8 HWC [n,s*h,s*w,out_nc]: the /255, channel swap,
        clamp, x255, round of test_image/test.py:31-40 run on the device (esrp_rrdbnet_forward_u8)."""
        if img.dim() != 4 or img.dtype != torch.uint8 or img.device != self.device or img.shape[3] != self.cfg[0]:
            raise RuntimeError(f"forward_u8 expects a uint8 [n,h,w,{self.cfg[0]}] tensor on {self.device}")
        img = img.contiguous()
        n, h, w, _ = img.shape
        key = ("u8", n, h, w)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = self.lib.esrp_rrdbnet_workspace_bytes_u8(self.handle, n, h, w)
            if nbytes < 0:
                raise RuntimeError("esrp_rrdbnet_workspace_bytes_u8 failed")
            if len(self._ws) >= 4:
                self._ws.clear()
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        base = ws.data_ptr()
        aligned = (base + 1023) // 1024 * 1024
        y = torch.empty((n, h * self.upscale, w * self.upscale, self.out_nc), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.esrp_rrdbnet_forward_u8(self.handle, img.data_ptr(), y.data_ptr(), n, h, w, aligned,
                                                        ws.numel() - (aligned - base), int(bool(bgr)),
                                                        torch.cuda.current_stream(self.device).cuda_stream),
                       "esrp_rrdbnet_forward_u8")
        return y

    @property
    def num_launches(self) -> int:
        return self.lib.esrp_rrdbnet_num_launches(self.handle)

    @property
    def num_chained_convs(self) -> int:
        """Conv launches that the persistent chains of the last planned forward replaced."""
        return self.lib.esrp_rrdbnet_num_chained_convs(self.handle)

    @property
    def num_pair_launches(self) -> int:
        """Conv launches of the last planned forward that run as CTA pairs (cta_group::2; ESRP_PAIR=1 only)."""
        return self.lib.esrp_rrdbnet_num_pair_launches(self.handle)

    def set_timing(self, enable: bool) -> None:
        """Bracket the dense-block convs of the trunk of every following forward with CUDA events (esrp_rrdbnet_set_timing)."""
        _lib.check(self.lib.esrp_rrdbnet_set_timing(self.handle, int(bool(enable))), "esrp_rrdbnet_set_timing")

    def trunk_times_ms(self, last: int = 64) -> List[float]:
        """Device milliseconds of the trunk's dense-block convs in the last timed forwards (synchronises the device)."""
        torch.cuda.synchronize(self.device)
        buf = (C.c_float * last)()
        n = self.lib.esrp_rrdbnet_get_timing(self.handle, buf, last)
        if n < 0:
            raise RuntimeError("esrp_rrdbnet_get_timing failed: " + self.lib.esrp_last_error().decode())
        return [float(buf[i]) for i in range(n)]

    def set_chain(self, enable: bool) -> None:
        """One persistent launch for the dense-block convs of the trunk (opt-in, see include/esrp.h) or one launch per conv (default)."""
        _lib.check(self.lib.esrp_rrdbnet_set_chain(self.handle, int(bool(enable))), "esrp_rrdbnet_set_chain")

    # -- training --------------------------------------------------------------------------------
    def _check_input(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or x.dtype != torch.float32 or x.device != self.device:
            raise RuntimeError(f"RRDBNet forward expects an fp32 NCHW tensor on {self.device}, got {x.dtype} {tuple(x.shape)} on {x.device}")
        if x.shape[1] != self.cfg[0]:
            raise RuntimeError(f"RRDBNet forward: expected {self.cfg[0]} input channels, got {x.shape[1]}")
        return x.contiguous()

    def train_forward(self, x: torch.Tensor, noise: bool, seed: int):
        """Forward that keeps the activations the backward needs (esrp_rrdbnet_train_forward).  Returns
        (y, token); the token names this forward's saved state: a later train_forward on the same shape
        reuses the workspace and invalidates older tokens."""
        x = self._check_input(x)
        n, _, h, w = x.shape
        key = ("train", n, h, w)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = self.lib.esrp_rrdbnet_train_workspace_bytes(self.handle, n, h, w)
            if nbytes < 0:
                raise RuntimeError("esrp_rrdbnet_train_workspace_bytes failed")
            for k in [k for k in self._ws if k[0] == "train"]:
                del self._ws[k]
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        base = ws.data_ptr()
        aligned = (base + 1023) // 1024 * 1024
        y = torch.empty((n, self.out_nc, h * self.upscale, w * self.upscale), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.esrp_rrdbnet_train_forward(self.handle, x.data_ptr(), y.data_ptr(), n, h, w, aligned,
                                                           ws.numel() - (aligned - base), int(bool(noise)),
                                                           seed & 0xFFFFFFFFFFFFFFFF,
                                                           torch.cuda.current_stream(self.device).cuda_stream),
                       "esrp_rrdbnet_train_forward")
        self._train_token = getattr(self, "_train_token", 0) + 1
        self._train_ws = (ws, aligned)
        return y, self._train_token

    def _grad_layout(self):
        lay = getattr(self, "_glay", None)
        if lay is None:
            import numpy as np
            dims = (C.c_int32 * 4)()
            shapes, offs, off = [], [], 0
            for i in range(len(self.keys)):
                _lib.check(self.lib.esrp_rrdbnet_tensor_shape(self.handle, i, dims), "esrp_rrdbnet_tensor_shape")
                shp = tuple(int(d) for d in dims if d > 0)
                shapes.append(shp)
                offs.append(off)
                numel = 1
                for d in shp:
                    numel *= d
                off += (numel + 3) // 4 * 4  # 16-byte aligned slices of one flat buffer
            lay = self._glay = (shapes, np.asarray(offs, dtype=np.uint64), off)
        return lay

    def backward(self, dy: torch.Tensor, token: int, needs: List[bool]):
        """dL/dy -> list of per-tensor gradients (views into ONE flat fp32 buffer, in key order; None where
        `needs` is False) plus the flat buffer itself (what a data-parallel all-reduce operates on)."""
        import numpy as np
        if token != getattr(self, "_train_token", None):
            raise RuntimeError("RRDBNet backward: the saved activations of this forward were overwritten by a later "
                               "training forward of the same module (one forward/backward pair at a time)")
        if dy.dtype != torch.float32 or dy.device != self.device:
            raise RuntimeError("RRDBNet backward expects an fp32 gradient on the module's device")
        dy = dy.contiguous()
        shapes, offs, total = self._grad_layout()
        flat = torch.empty(total, dtype=torch.float32, device=self.device)
        ptrs = offs * np.uint64(4) + np.uint64(flat.data_ptr())
        if not all(needs):
            ptrs = np.where(np.asarray(needs, dtype=bool), ptrs, np.uint64(0))
        ptrs = np.ascontiguousarray(ptrs)
        ws, aligned = self._train_ws
        with torch.cuda.device(self.device):
            _lib.check(self.lib.esrp_rrdbnet_backward(self.handle, dy.data_ptr(),
                                                      ptrs.ctypes.data_as(C.POINTER(C.c_void_p)), len(shapes), aligned,
                                                      torch.cuda.current_stream(self.device).cuda_stream),
                       "esrp_rrdbnet_backward")
        self._train_token += 1  # consumed
        # one split call instead of 771 slice ops; slices are padded to 16 bytes, only odd-sized tensors need a narrow
        sizes, numels = self._grad_split_sizes()
        parts = flat.split(sizes)
        grads = []
        for part, shp, numel, need in zip(parts, shapes, numels, needs):
            if not need:
                grads.append(None)
            elif part.numel() == numel:
                grads.append(part.view(shp))
            else:
                grads.append(part[:numel].view(shp))
        return grads, flat

    def backward_flat(self, dy: torch.Tensor, token: int):
        """Like `backward` with every tensor needed, but into a PERSISTENT flat buffer whose per-tensor views are created
        once (`grad_views`): no per-step allocation, split or view creation for the 771 tensors.  Returns the flat buffer."""
        import numpy as np
        if token != getattr(self, "_train_token", None):
            raise RuntimeError("RRDBNet backward: the saved activations of this forward were overwritten by a later "
                               "training forward of the same module (one forward/backward pair at a time)")
        if dy.dtype != torch.float32 or dy.device != self.device:
            raise RuntimeError("RRDBNet backward expects an fp32 gradient on the module's device")
        dy = dy.contiguous()
        if getattr(self, "_flat_grad", None) is None:
            shapes, offs, total = self._grad_layout()
            self._flat_grad = torch.zeros(total, dtype=torch.float32, device=self.device)
            ptrs = np.ascontiguousarray(offs * np.uint64(4) + np.uint64(self._flat_grad.data_ptr()))
            self._flat_ptrs = ptrs
            sizes, numels = self._grad_split_sizes()
            self.grad_views = [part[:n].view(shp) if part.numel() != n else part.view(shp)
                               for part, shp, n in zip(self._flat_grad.split(sizes), shapes, numels)]
        ws, aligned = self._train_ws
        with torch.cuda.device(self.device):
            _lib.check(self.lib.esrp_rrdbnet_backward(self.handle, dy.data_ptr(),
                                                      self._flat_ptrs.ctypes.data_as(C.POINTER(C.c_void_p)), len(self.grad_views), aligned,
                                                      torch.cuda.current_stream(self.device).cuda_stream),
                       "esrp_rrdbnet_backward")
        self._train_token += 1  # consumed
        return self._flat_grad

    def _grad_split_sizes(self):
        c = getattr(self, "_gsplit", None)
        if c is None:
            shapes, offs, total = self._grad_layout()
            o = offs.tolist() + [total]
            sizes = [int(o[i + 1] - o[i]) for i in range(len(shapes))]
            numels = []
            for shp in shapes:
                n = 1
                for d in shp:
                    n *= d
                numels.append(n)
            c = self._gsplit = (sizes, numels)
        return c

    def train_launches(self):
        return (self.lib.esrp_rrdbnet_train_num_launches(self.handle, 0),
                self.lib.esrp_rrdbnet_train_num_launches(self.handle, 1))
