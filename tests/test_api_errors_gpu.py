"""GPU tests of the C ABI's error behaviour: misaligned device pointers, capacities that are too small (the call reports the
size that would have been needed), invalid shard geometry.  Errors are codes + hevcb_last_error text, never silent truncation."""
import ctypes as C

import numpy as np
import pytest

import hevcbitstream_b200 as hb
from hevcbitstream_b200 import shard as hs
from hevcbitstream_b200._lib import ScanSummary
from oracle import ref
from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]


def test_misaligned_device_buffer_is_rejected(ctx):
    import torch

    d = torch.zeros(4096 + 64, dtype=torch.uint8, device="cuda")
    a = [torch.zeros(64, dtype=torch.int64, device="cuda") for _ in range(4)]
    summ = torch.zeros(8, dtype=torch.int64, device="cuda")
    with pytest.raises(hb.HevcbError) as e:
        ctx.scan_strip_device_raw(d.data_ptr() + 1, 4096, a[0].data_ptr(), a[1].data_ptr(), 64, None, a[2].data_ptr(), a[3].data_ptr(), summ.data_ptr(), 0)
    assert e.value.code == -103 and "aligned" in str(e.value)
    with pytest.raises(hb.HevcbError) as e:
        ctx.insert_device(d[1:], a[0][:1], a[1][:1], n_nals=1)
    assert e.value.code == -103


def test_nal_capacity_overflow_reports_the_true_count(ctx):
    s = util.c2_stream(64, 64 * 2000, seed=5)
    size = s.size - ref.PAD
    ns, ne = np.zeros(100, np.int64), np.zeros(100, np.int64)
    sm = ScanSummary()
    rc = ctx._L.hevcb_scan_strip_host(ctx._h, s.ctypes.data_as(C.c_void_p), size, ns.ctypes.data_as(C.c_void_p), ne.ctypes.data_as(C.c_void_p), 100, None, None, None,
                                     C.byref(sm))
    assert rc == -104 and sm.overflow == 1 and sm.n_nals == 2000
    st, en, _ = ref.scan_all_with_tail(s, size)
    assert np.array_equal(ns, st[:100]) and np.array_equal(ne, en[:100])  # the first cap_nals entries are valid
    assert b"exceed cap_nals" in ctx._L.hevcb_last_error(ctx._h)


def test_pair_capacity_and_rewrite_capacity(ctx):
    import torch

    s = ref.gen_stream(seed=1, profile=1, n_slices=500, payload_min=10, payload_max=100, ps_period=50)
    size = s.size - ref.PAD
    d = torch.zeros(size + 32, dtype=torch.uint8, device="cuda")
    d[:size] = torch.from_numpy(s[:size].copy())
    scan = ctx.scan_strip_device(d, size=size)
    with pytest.raises(hb.HevcbError) as e:
        ctx.parse_device(d, scan, cap_pairs=1000)
    assert e.value.code == -104 and "syntax elements" in str(e.value)
    parsed = ctx.parse_device(d, scan)
    with pytest.raises(hb.HevcbError) as e:
        ctx.rewrite_device(d, scan, parsed, [], size=size, out_cap=size // 2)
    assert e.value.code == -104
    with pytest.raises(hb.HevcbError):
        ctx.rewrite_device(d, scan, parsed, [(4, "no_such_field", 0, 1)], size=size)


def test_shard_geometry_is_validated(ctx):
    import torch

    d = torch.zeros(8192 + 64, dtype=torch.uint8, device="cuda")
    with pytest.raises(hb.HevcbError) as e:
        hs.scan_strip_shard(ctx, d, 4096, 1, True, False)  # an inner shard needs >= 3 halo bytes
    assert e.value.code == -102
    with pytest.raises(hb.HevcbError):
        hs.scan_strip_shard(ctx, d, 4096, 16, False, True)  # the last shard has no halo
    rec = hs.scan_strip_shard(ctx, d, 4096, 16, True, False).record
    with pytest.raises(hb.HevcbError):
        hs.stitch([rec])  # a single record that is not marked last: inconsistent
