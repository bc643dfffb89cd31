"""Python mirror of include/hevcb.h: a Context object whose methods call the C ABI.

Device-resident methods take/return torch CUDA tensors (torch is used only for memory and streams);
host methods take/return numpy arrays and go through the `_host` entry points, i.e. they include the
host<->device copies.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import HevcbError, ScanSummary, load_library


@dataclass
class ScanResult:
    n_nals: int
    n_terminated: int
    last_rc: int
    last_start: int
    last_end: int
    rbsp_bytes: int
    n_epb: int
    nal_start: object  # torch tensor (device API) or numpy array (host API), length >= n_nals
    nal_end: object
    rbsp_off: object
    rbsp_end: object
    rbsp: object  # EPB-free image or None


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One libhevcb200 context bound to a CUDA device (hevcb_create / hevcb_destroy)."""

    def __init__(self, device: int = 0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.hevcb_create(device, C.byref(h))
        if rc != 0:
            raise HevcbError(rc, self._L.hevcb_last_error(None).decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.hevcb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise HevcbError(rc, self._L.hevcb_last_error(self._h).decode())

    @property
    def launch_count(self) -> int:
        return int(self._L.hevcb_launch_count(self._h))

    @property
    def sm_count(self) -> int:
        return int(self._L.hevcb_sm_count(self._h))

    # ---- scan + strip -----------------------------------------------------------------------
    def scan_strip_device_raw(self, d_buf, size, d_start, d_end, cap, d_rbsp, d_off, d_rend, d_summary, stream_ptr):
        """Direct call of hevcb_scan_strip_device with raw device pointers (ints)."""
        self._check(self._L.hevcb_scan_strip_device(self._h, d_buf, size, d_start, d_end, cap, d_rbsp, d_off, d_rend, d_summary, stream_ptr))

    def scan_strip_device(self, buf, size=None, cap_nals=None, want_rbsp=True, out=None, sync=True):
        """buf: torch.uint8 CUDA tensor.  Returns ScanResult with torch tensors (device resident)."""
        import torch

        assert buf.is_cuda and buf.dtype == torch.uint8 and buf.is_contiguous()
        size = int(buf.numel() if size is None else size)
        if cap_nals is None:
            cap_nals = size // 3 + 8
        dev = buf.device
        if out is None:
            out = dict(
                nal_start=torch.empty(cap_nals, dtype=torch.int64, device=dev),
                nal_end=torch.empty(cap_nals, dtype=torch.int64, device=dev),
                rbsp_off=torch.empty(cap_nals, dtype=torch.int64, device=dev),
                rbsp_end=torch.empty(cap_nals, dtype=torch.int64, device=dev),
                rbsp=torch.empty(size + 16, dtype=torch.uint8, device=dev) if want_rbsp else None,
                summary=torch.zeros(8, dtype=torch.int64, device=dev),
            )
        stream = torch.cuda.current_stream(dev).cuda_stream
        self.scan_strip_device_raw(
            buf.data_ptr(), size, out["nal_start"].data_ptr(), out["nal_end"].data_ptr(), cap_nals,
            out["rbsp"].data_ptr() if out.get("rbsp") is not None else None,
            out["rbsp_off"].data_ptr(), out["rbsp_end"].data_ptr(), out["summary"].data_ptr(), stream,
        )
        if not sync:
            return out
        s = out["summary"].cpu().numpy()
        n_nals, n_term = int(s[0]), int(s[1])
        last_rc = int(np.int32(s[2] & 0xFFFFFFFF))
        overflow = int(s[2] >> 32)
        if overflow:
            raise HevcbError(-104, f"{n_nals} NALs exceed cap_nals {cap_nals}")
        return ScanResult(n_nals, n_term, last_rc, int(s[3]), int(s[4]), int(s[5]), int(s[6]),
                          out["nal_start"], out["nal_end"], out["rbsp_off"], out["rbsp_end"], out.get("rbsp"))

    def scan_strip_host(self, buf: np.ndarray, size=None, cap_nals=None, want_rbsp=True) -> ScanResult:
        """buf: numpy uint8 array (host).  Includes H2D/D2H copies (hevcb_scan_strip_host)."""
        assert buf.dtype == np.uint8
        size = int(buf.size if size is None else size)
        if cap_nals is None:
            cap_nals = size // 3 + 8
        ns = np.empty(cap_nals, dtype=np.int64)
        ne = np.empty(cap_nals, dtype=np.int64)
        ro = np.empty(cap_nals, dtype=np.int64)
        re = np.empty(cap_nals, dtype=np.int64)
        rb = np.empty(size + 16, dtype=np.uint8) if want_rbsp else None
        sm = ScanSummary()
        self._check(self._L.hevcb_scan_strip_host(self._h, _np_ptr(buf), size, _np_ptr(ns), _np_ptr(ne), cap_nals,
                                                 _np_ptr(rb), _np_ptr(ro), _np_ptr(re), C.byref(sm)))
        n = sm.n_nals
        return ScanResult(n, sm.n_terminated, sm.last_rc, sm.last_start, sm.last_end, sm.rbsp_bytes, sm.n_epb,
                          ns[:n], ne[:n], ro[:n], re[:n], rb[: sm.rbsp_bytes] if rb is not None else None)
