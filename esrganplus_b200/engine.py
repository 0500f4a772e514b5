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
    def sync_weights(self, named_params: Dict[str, torch.Tensor]) -> bool:
        """Repack if any source tensor changed. `named_params`: state_dict-style key -> fp32 tensor
        on this device. Returns True when a repack was launched."""
        tensors = []
        for k in self.keys:
            t = named_params.get(k)
            if t is None:
                raise RuntimeError(f"RRDBNet engine: parameter {k!r} missing from the module")
            if t.device != self.device or t.dtype != torch.float32:
                raise RuntimeError(f"RRDBNet engine: {k} must be fp32 on {self.device}, got {t.dtype} on {t.device}")
            tensors.append(t if t.is_contiguous() else t.contiguous())
        sig = tuple((t.data_ptr(), t._version) for t in tensors)
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

    @property
    def num_launches(self) -> int:
        return self.lib.esrp_rrdbnet_num_launches(self.handle)
