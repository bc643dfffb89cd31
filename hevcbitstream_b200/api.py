"""Python mirror of include/hevcb.h: a Context object whose methods call the C ABI.

Device-resident methods take/return torch CUDA tensors (torch is used only for memory and streams);
host methods take/return numpy arrays and go through the `_host` entry points, i.e. they include the
host<->device copies.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import BsOp, EditRule, EditSet, HevcbError, InsertSummary, ParseBuffers, ParseSummary, ScanSummary, StreamIndex, load_library


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

    # ---- the device bit reader / writer on their own (hevcb_bs_read_host / hevcb_bs_write_host) ----------
    BS_KINDS = {"u": 0, "f": 0, "u1": 1, "u8": 2, "ue": 3, "se": 4, "skip": 5}

    def bs_read(self, data: bytes, ops):
        """ops: list of ("u", n) / ("u1",) / ("u8",) / ("ue",) / ("se",) / ("skip", n).  Returns (values, bitpos, overrun) lists."""
        n = len(ops)
        arr = (BsOp * max(n, 1))()
        for i, op in enumerate(ops):
            arr[i].kind = self.BS_KINDS[op[0]]
            arr[i].n = int(op[1]) if len(op) > 1 else 0
        buf = np.frombuffer(bytes(data), np.uint8).copy()
        vals = np.zeros(max(n, 1), np.int32)
        pos = np.zeros(max(n, 1), np.int64)
        ovr = np.zeros(max(n, 1), np.int32)
        self._check(self._L.hevcb_bs_read_host(self._h, _np_ptr(buf) if buf.size else None, int(buf.size), arr, n, _np_ptr(vals), _np_ptr(pos), _np_ptr(ovr)))
        return vals[:n].tolist(), pos[:n].tolist(), ovr[:n].tolist()

    def bs_write(self, ops, cap: int):
        """ops: list of ("u", n, v) / ("u1", v) / ("u8", v) / ("ue", v) / ("se", v).  Returns (bytes, bits written, overrun)."""
        n = len(ops)
        arr = (BsOp * max(n, 1))()
        for i, op in enumerate(ops):
            arr[i].kind = self.BS_KINDS[op[0]]
            arr[i].n = int(op[1]) if op[0] in ("u", "f") else 0
            v = int(op[-1])
            arr[i].value = v - (1 << 32) if v >= (1 << 31) else v
        out = np.zeros(max(cap, 1), np.uint8)
        bits = C.c_int64(0)
        ovr = C.c_int32(0)
        self._check(self._L.hevcb_bs_write_host(self._h, arr, n, _np_ptr(out), int(cap), C.byref(bits), C.byref(ovr)))
        return out[:cap].tobytes(), int(bits.value), int(ovr.value)

    def trace_name(self, kind: int, code: int) -> str:
        """Text the reference prints for a trace record (hevcb_trace_name)."""
        b = C.create_string_buffer(160)
        n = self._L.hevcb_trace_name(int(kind), int(code) & 0xFFFFFFFF, b, 160)
        if n < 0:
            raise HevcbError(-101, f"unknown trace code {code:#x} for kind {kind}")
        return b.value.decode()

    # ---- scan + strip -----------------------------------------------------------------------
    def scan_strip_device_raw(self, d_buf, size, d_start, d_end, cap, d_rbsp, d_off, d_rend, d_summary, stream_ptr):
        """Direct call of hevcb_scan_strip_device with raw device pointers (ints)."""
        self._check(self._L.hevcb_scan_strip_device(self._h, d_buf, size, d_start, d_end, cap, d_rbsp, d_off, d_rend, d_summary, stream_ptr))

    def scan_strip_device(self, buf, size=None, cap_nals=None, want_rbsp=True, out=None, sync=True, _grow=False):
        """buf: torch.uint8 CUDA tensor.  Returns ScanResult with torch tensors (device resident)."""
        import torch

        assert buf.is_cuda and buf.dtype == torch.uint8 and buf.is_contiguous()
        size = int(buf.numel() if size is None else size)
        auto_cap = (cap_nals is None or _grow) and out is None and sync
        if cap_nals is None:
            # a realistic bound (one NAL per 64 bytes: 0.5 bytes of arrays per input byte); when a stream holds more NALs the
            # summary reports the true count and the call is repeated with it (only when the arrays are ours and we synchronise)
            cap_nals = size // 64 + 1024
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
            if auto_cap and cap_nals < size // 3 + 8:
                del out
                return self.scan_strip_device(buf, size=size, cap_nals=min(size // 3 + 8, max(n_nals + 8, 4 * cap_nals)), want_rbsp=want_rbsp, _grow=True)
            raise HevcbError(-104, f"{n_nals} NALs exceed cap_nals {cap_nals}")
        return ScanResult(n_nals, n_term, last_rc, int(s[3]), int(s[4]), int(s[5]), int(s[6]),
                          out["nal_start"], out["nal_end"], out["rbsp_off"], out["rbsp_end"], out.get("rbsp"))

    def scan_strip_host(self, buf: np.ndarray, size=None, cap_nals=None, want_rbsp=True, _grow=False) -> ScanResult:
        """buf: numpy uint8 array (host).  Includes H2D/D2H copies (hevcb_scan_strip_host)."""
        assert buf.dtype == np.uint8
        size = int(buf.size if size is None else size)
        auto_cap = cap_nals is None or _grow
        if cap_nals is None:
            cap_nals = size // 64 + 1024  # see scan_strip_device
        ns = np.empty(cap_nals, dtype=np.int64)
        ne = np.empty(cap_nals, dtype=np.int64)
        ro = np.empty(cap_nals, dtype=np.int64)
        re = np.empty(cap_nals, dtype=np.int64)
        rb = np.empty(size + 16, dtype=np.uint8) if want_rbsp else None
        sm = ScanSummary()
        rc = self._L.hevcb_scan_strip_host(self._h, _np_ptr(buf), size, _np_ptr(ns), _np_ptr(ne), cap_nals,
                                           _np_ptr(rb), _np_ptr(ro), _np_ptr(re), C.byref(sm))
        if rc == -104 and auto_cap and cap_nals < size // 3 + 8:  # HEVCB_E_CAPACITY: grow (the summary's count is a lower bound)
            return self.scan_strip_host(buf, size=size, cap_nals=min(size // 3 + 8, max(int(sm.n_nals) + 8, 4 * cap_nals)), want_rbsp=want_rbsp,
                                        _grow=True)
        self._check(rc)
        n = sm.n_nals
        return ScanResult(n, sm.n_terminated, sm.last_rc, sm.last_start, sm.last_end, sm.rbsp_bytes, sm.n_epb,
                          ns[:n], ne[:n], ro[:n], re[:n], rb[: sm.rbsp_bytes] if rb is not None else None)

    # ---- EPB insertion (rbsp_to_nal) ----------------------------------------------------------
    def insert_device(self, rbsp, rbsp_off, rbsp_end, n_nals=None, start_code_len=0, out_cap=None, sync=True):
        """rbsp_to_nal for every segment rbsp[rbsp_off[k]:rbsp_end[k]] (torch CUDA tensors).  Returns dict(out, out_off,
        summary) plus out_bytes / n_inserted when sync."""
        import torch

        n = int(rbsp_off.numel() if n_nals is None else n_nals)
        dev = rbsp.device
        if out_cap is None:  # worst case: one 03 per two payload bytes
            payload = int((rbsp_end[:n] - rbsp_off[:n]).clamp(min=0).sum().item()) if n else 0
            out_cap = payload * 3 // 2 + (start_code_len + 1) * n + 64
        out = dict(out=torch.empty(out_cap + 16, dtype=torch.uint8, device=dev), out_off=torch.empty(n + 1, dtype=torch.int64, device=dev),
                   summary=torch.zeros(4, dtype=torch.int64, device=dev))
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._check(self._L.hevcb_insert_device(self._h, rbsp.data_ptr(), rbsp_off.data_ptr(), rbsp_end.data_ptr(), n, start_code_len,
                                                out["out"].data_ptr(), out_cap, out["out_off"].data_ptr(), out["summary"].data_ptr(), stream))
        if sync:
            s = out["summary"].cpu().numpy()
            out["out_bytes"], out["n_inserted"] = int(s[1]), int(s[2])
            if int(s[3]) & 0xFFFFFFFF:
                raise HevcbError(-104, f"{out['out_bytes']} output bytes exceed out_cap {out_cap}")
        return out

    def insert_host(self, rbsp: np.ndarray, rbsp_off: np.ndarray, rbsp_end: np.ndarray, start_code_len=0, out_cap=None):
        """Host arrays in, (out bytes, out_off[n+1], n_inserted) out; includes the copies (hevcb_insert_host)."""
        assert rbsp.dtype == np.uint8
        rbsp_off = np.ascontiguousarray(rbsp_off, dtype=np.int64)
        rbsp_end = np.ascontiguousarray(rbsp_end, dtype=np.int64)
        n = int(rbsp_off.size)
        if out_cap is None:  # worst case: one 03 per two payload bytes
            payload = int(np.maximum(rbsp_end - rbsp_off, 0).sum()) if n else 0
            out_cap = payload * 3 // 2 + (start_code_len + 1) * n + 64
        out = np.empty(out_cap, dtype=np.uint8)
        out_off = np.empty(n + 1, dtype=np.int64)
        sm = InsertSummary()
        self._check(self._L.hevcb_insert_host(self._h, _np_ptr(rbsp), int(rbsp.size), _np_ptr(rbsp_off), _np_ptr(rbsp_end), n, start_code_len,
                                              _np_ptr(out), out_cap, _np_ptr(out_off), C.byref(sm)))
        return out[: sm.out_bytes], out_off, int(sm.n_inserted)

    # ---- length-prefixed framing <-> Annex-B ------------------------------------------------------
    def reframe_device(self, buf, nal_start, nal_end, n_nals=None, start_code_len=0, len_size=0, out_cap=None, sync=True):
        """Copies the NAL units buf[nal_start[k]:nal_end[k]] behind a start code (start_code_len 3 / 4) or a big-endian length
        of len_size bytes (hevcb_reframe_device).  Returns dict(out, out_off, summary [, out_bytes])."""
        import torch

        n = int(nal_start.numel() if n_nals is None else n_nals)
        dev = buf.device
        if out_cap is None:
            out_cap = int((nal_end[:n] - nal_start[:n]).clamp(min=0).sum().item()) + (start_code_len + len_size) * n + 64 if n else 64
        out = dict(out=torch.empty(out_cap + 16, dtype=torch.uint8, device=dev), out_off=torch.empty(n + 1, dtype=torch.int64, device=dev),
                   summary=torch.zeros(4, dtype=torch.int64, device=dev))
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._check(self._L.hevcb_reframe_device(self._h, buf.data_ptr(), nal_start.data_ptr(), nal_end.data_ptr(), n, start_code_len, len_size,
                                                 out["out"].data_ptr(), out_cap, out["out_off"].data_ptr(), out["summary"].data_ptr(), stream))
        if sync:
            s = out["summary"].cpu().numpy()
            out["out_bytes"] = int(s[1])
            if int(s[3]) & 0xFFFFFFFF:
                raise HevcbError(-104, f"{out['out_bytes']} output bytes exceed out_cap {out_cap}")
        return out

    def lenpref_index_device(self, buf, size=None, len_size=4, sample_off=None, cap_nals=None):
        """NAL extents of length-prefixed data (hevcb_lenpref_index_device).  sample_off: int64 CUDA tensor of n_samples + 1
        boundaries or None.  Returns (nal_start, nal_end, n_nals, n_bad_samples)."""
        import torch

        size = int(buf.numel() if size is None else size)
        dev = buf.device
        if cap_nals is None:
            cap_nals = size // (len_size + 1) + 8
        ns = torch.empty(cap_nals, dtype=torch.int64, device=dev)
        ne = torch.empty(cap_nals, dtype=torch.int64, device=dev)
        tot = torch.zeros(2, dtype=torch.int64, device=dev)
        n_samples = int(sample_off.numel() - 1) if sample_off is not None else 1
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._check(self._L.hevcb_lenpref_index_device(self._h, buf.data_ptr(), size, len_size, sample_off.data_ptr() if sample_off is not None else None,
                                                       n_samples, ns.data_ptr(), ne.data_ptr(), cap_nals, tot.data_ptr(), stream))
        t = tot.cpu().numpy()
        if int(t[0]) > cap_nals:
            raise HevcbError(-104, f"{int(t[0])} NALs exceed cap_nals {cap_nals}")
        return ns, ne, int(t[0]), int(t[1])

    # ---- batched header parse -----------------------------------------------------------------
    def parse_device(self, buf, scan: "ScanResult", cap_pairs=None, sync=True, trace=False, aux=False, spec=False):
        """Header parse of every NAL found by scan_strip_device (device resident).  Returns a dict of torch tensors
        (rc, nal_hdr, kind, ubflag, hdr_end, cols[8][n], pair_off[n+1], pair_field, pair_value) and `summary`.
        trace=True: the read_debug variant (include/hevcb.h): the lists hold what read_debug_hevc_nal_unit prints, with
        `pair_pos` = bit position of every record."""
        import torch

        n = int(scan.n_nals)
        dev = buf.device
        if cap_pairs is None:
            cap_pairs = (96 if trace else 64) * n + 4096
        out = dict(
            rc=torch.empty(max(n, 1), dtype=torch.int32, device=dev), nal_hdr=torch.empty(max(n, 1), dtype=torch.int32, device=dev),
            kind=torch.empty(max(n, 1), dtype=torch.uint8, device=dev), ubflag=torch.empty(max(n, 1), dtype=torch.uint8, device=dev),
            hdr_end=torch.empty(max(n, 1), dtype=torch.int32, device=dev), cols=torch.empty((8, max(n, 1)), dtype=torch.int32, device=dev),
            pair_off=torch.empty(n + 1, dtype=torch.int64, device=dev), pair_field=torch.empty(cap_pairs, dtype=torch.int32, device=dev),
            pair_value=torch.empty(cap_pairs, dtype=torch.int32, device=dev), summary=torch.zeros(8, dtype=torch.int64, device=dev),
        )
        pb = ParseBuffers(out["rc"].data_ptr(), out["nal_hdr"].data_ptr(), out["kind"].data_ptr(), out["ubflag"].data_ptr(),
                          out["hdr_end"].data_ptr(), out["cols"].data_ptr(), out["pair_off"].data_ptr(), out["pair_field"].data_ptr(),
                          out["pair_value"].data_ptr(), cap_pairs)
        if trace:
            out["pair_pos"] = torch.empty(cap_pairs, dtype=torch.int32, device=dev)
            pb.pair_pos = out["pair_pos"].data_ptr()
        if aux:  # extension mode: AUD / EOS / EOB / filler / SEI NALs are parsed instead of returning -1 (kind 5)
            pb.flags = 1
        if spec:  # spec-correct mode (HEVCB_PARSE_SPEC, include/hevcb.h): the standard's syntax where the reference departs from it
            pb.flags |= 2
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._check(self._L.hevcb_parse_device(self._h, buf.data_ptr(), scan.nal_start.data_ptr(), scan.nal_end.data_ptr(), scan.rbsp.data_ptr(),
                                               scan.rbsp_off.data_ptr(), scan.rbsp_end.data_ptr(), n, C.byref(pb), out["summary"].data_ptr(), stream))
        out["n"] = n
        out["cap_pairs"] = cap_pairs
        if sync:
            s = out["summary"].cpu().numpy()
            out["n_ok"], out["n_pairs"], out["n_vps"], out["n_sps"], out["n_pps"], out["n_slices"] = (int(x) for x in s[1:7])
            if int(s[7]) & 0xFFFFFFFF:
                raise HevcbError(-104, f"{out['n_pairs']} syntax elements exceed cap_pairs {cap_pairs}")
        return out

    # ---- header rewrite -------------------------------------------------------------------------
    def rewrite_device(self, buf, scan: "ScanResult", parsed: dict, edits=(), size=None, out_cap=None, sync=True):
        """read -> edit -> write_hevc_nal_unit -> rbsp_to_nal for every NAL (hevcb_rewrite_device); must directly follow
        parse_device(buf, scan).  edits: iterable of (kind, field index or name, op, arg).  Returns dict(out, out_start,
        out_end, summary) + out_bytes / n_rewritten / n_inserted when sync."""
        import torch

        n = int(scan.n_nals)
        size = int(buf.numel() if size is None else size)
        dev = buf.device
        es = EditSet()
        for i, (kind, field, op, arg) in enumerate(edits):
            if isinstance(field, str):
                idx = int(self._L.hevcb_field_index(kind, field.encode()))
                if idx < 0:
                    raise HevcbError(-102, f"unknown field {field!r}")
                field = idx
            es.e[i] = EditRule(kind, field, op, arg)
            es.n = i + 1
        if out_cap is None:
            out_cap = size + size // 2 + 64 * n + 4096
        out = dict(out=torch.empty(out_cap + 16, dtype=torch.uint8, device=dev), out_start=torch.empty(max(n, 1), dtype=torch.int64, device=dev),
                   out_end=torch.empty(max(n, 1), dtype=torch.int64, device=dev), summary=torch.zeros(5, dtype=torch.int64, device=dev))
        pb = ParseBuffers(parsed["rc"].data_ptr(), parsed["nal_hdr"].data_ptr(), parsed["kind"].data_ptr(), parsed["ubflag"].data_ptr(),
                          parsed["hdr_end"].data_ptr(), parsed["cols"].data_ptr(), parsed["pair_off"].data_ptr(), parsed["pair_field"].data_ptr(),
                          parsed["pair_value"].data_ptr(), parsed["cap_pairs"])
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._check(self._L.hevcb_rewrite_device(self._h, buf.data_ptr(), size, scan.nal_start.data_ptr(), scan.nal_end.data_ptr(), scan.rbsp.data_ptr(),
                                                 scan.rbsp_off.data_ptr(), scan.rbsp_end.data_ptr(), n, C.byref(pb), C.byref(es), out["out"].data_ptr(),
                                                 out_cap, out["out_start"].data_ptr(), out["out_end"].data_ptr(), out["summary"].data_ptr(), stream))
        if sync:
            s = out["summary"].cpu().numpy()
            out["n_rewritten"], out["out_bytes"], out["n_inserted"] = int(s[1]), int(s[2]), int(s[3])
            if int(s[4]) & 0xFFFFFFFF:
                raise HevcbError(-104, f"{out['out_bytes']} output bytes exceed out_cap {out_cap}")
        return out

    def index_host(self, buf: np.ndarray, size=None, cap_nals=None, cap_pairs=None, want_rbsp=True, flags=0) -> "HostIndex":
        """Annex-B bytes in host memory -> full index (hevcb_index_host: scan + strip + parse, copies included)."""
        assert buf.dtype == np.uint8
        size = int(buf.size if size is None else size)
        if cap_nals is None and cap_pairs is None:
            # a realistic bound first (one NAL per 64 bytes: 0.6 bytes of arrays per input byte); a denser stream reports
            # HEVCB_E_CAPACITY and the call repeats with the worst case (a start code every 3 bytes: ~11 bytes per input byte)
            try:
                cn = size // 64 + 1024
                return HostIndex(self, buf, size, cn, 64 * cn + 4096, want_rbsp, flags)
            except HevcbError as e:
                if e.code != -104:
                    raise
        if cap_nals is None:
            cap_nals = size // 3 + 8
        if cap_pairs is None:
            cap_pairs = 64 * min(cap_nals, size // 4 + 8) + 4096
        return HostIndex(self, buf, size, cap_nals, cap_pairs, want_rbsp, flags)


class HostIndex:
    """Owns the host arrays of a hevcb_stream_index and exposes hevcb_materialize."""

    def __init__(self, ctx: Context, buf, size, cap_nals, cap_pairs, want_rbsp, flags=0):
        self._L = ctx._L
        a = self.arrays = dict(
            nal_start=np.zeros(cap_nals, np.int64), nal_end=np.zeros(cap_nals, np.int64), rbsp_off=np.zeros(cap_nals, np.int64),
            rbsp_end=np.zeros(cap_nals, np.int64), rbsp=np.zeros(size + 16, np.uint8) if want_rbsp else None,
            rc=np.zeros(cap_nals, np.int32), nal_hdr=np.zeros(cap_nals, np.int32), kind=np.zeros(cap_nals, np.uint8),
            ubflag=np.zeros(cap_nals, np.uint8), hdr_end=np.zeros(cap_nals, np.int32), cols=np.zeros(8 * cap_nals, np.int32),
            pair_off=np.zeros(cap_nals + 1, np.int64), pair_field=np.zeros(cap_pairs, np.uint32), pair_value=np.zeros(cap_pairs, np.int32),
        )
        pb = ParseBuffers(_np_ptr(a["rc"]), _np_ptr(a["nal_hdr"]), _np_ptr(a["kind"]), _np_ptr(a["ubflag"]), _np_ptr(a["hdr_end"]), _np_ptr(a["cols"]),
                          _np_ptr(a["pair_off"]), _np_ptr(a["pair_field"]), _np_ptr(a["pair_value"]), cap_pairs)
        pb.flags = int(flags)  # HEVCB_PARSE_AUX = 1, HEVCB_PARSE_SPEC = 2
        self.idx = StreamIndex(cap_nals, _np_ptr(a["nal_start"]), _np_ptr(a["nal_end"]), _np_ptr(a["rbsp_off"]), _np_ptr(a["rbsp_end"]),
                               _np_ptr(a["rbsp"]), pb, ScanSummary(), ParseSummary())
        ctx._check(self._L.hevcb_index_host(ctx._h, _np_ptr(buf), size, C.byref(self.idx)))
        self.n = int(self.idx.scan.n_nals)
        self.scan = self.idx.scan
        self.parse = self.idx.parse
        n = self.n
        self.cols = a["cols"][: 8 * n].reshape(8, n) if n else np.zeros((8, 0), np.int32)

    def __getattr__(self, name):
        a = self.__dict__.get("arrays", {})
        if name in a:
            v = a[name]
            if v is None:
                return None
            if name in ("pair_field", "pair_value"):
                return v[: int(self.idx.parse.n_pairs)]
            if name == "pair_off":
                return v[: self.n + 1]
            if name == "rbsp":
                return v[: int(self.idx.scan.rbsp_bytes)]
            return v[: self.n]
        raise AttributeError(name)

    def materialize(self, k: int, nal, vps, sps, pps, sh) -> int:
        """Applies NAL k to numpy int32 struct images (any may be None).  Returns the reference's rc."""
        return int(self._L.hevcb_materialize(C.byref(self.idx), k, _np_ptr(nal), _np_ptr(vps), _np_ptr(sps), _np_ptr(pps), _np_ptr(sh)))
