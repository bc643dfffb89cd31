"""Direct known-answer tests of the DEVICE bit reader and bit writer (hevcb_bits / hevcb_bitwriter, the primitives every parser and
writer kernel is built on) through hevcb_bs_read_host / hevcb_bs_write_host: SURVEY Appendix B's vectors (obtained from the
reference) and random scripts against the oracle's restatement of bs.h (oracle/oracle_port.c, pinned on the reference)."""
import numpy as np
import pytest

from oracle import port

pytestmark = pytest.mark.gpu

H = lambda s: bytes.fromhex(s.replace(" ", ""))


def test_appendix_b_reads(ctx):
    v, pos, ovr = ctx.bs_read(H("A6 42 98 E2 04 8A"), [("ue",)] * 8)
    assert v == [0, 1, 2, 3, 4, 5, 6, 7]
    v, pos, ovr = ctx.bs_read(H("A6 42 98 E2 04 8A"), [("se",)] * 8)
    assert v == [0, 1, -1, 2, -2, 3, -3, 4]
    # ue on 00 00 (all zero, the buffer ends inside the prefix): 32767, and the cursor ends beyond the buffer
    v, pos, ovr = ctx.bs_read(H("00 00"), [("ue",)])
    assert v == [32767] and ovr == [1]
    # ue on a single byte 01 -> 127; 02 -> 63 with overrun 0 (p == end)
    assert ctx.bs_read(H("01"), [("ue",)])[0] == [127]
    v, pos, ovr = ctx.bs_read(H("02"), [("ue",)])
    assert v == [63] and ovr == [0] and pos == [8 + 5]
    # 32 leading zeros: 33 + 32 bits consumed, the result is the 32-bit suffix (1 << 32 evaluates to 1 on the x86 reference)
    v, pos, ovr = ctx.bs_read(H("00 00 00 00 80 00 12 34 80"), [("ue",)])
    assert pos == [65] and v == [0x2469]
    # reads past the end return 0 bits and keep counting
    v, pos, ovr = ctx.bs_read(H("FF"), [("u", 4), ("u", 8), ("u1",), ("u8",), ("skip", 3), ("u", 32)])
    assert v == [15, 0xF0, 0, 0, 0, 0] and pos == [4, 12, 13, 21, 24, 56] and ovr == [0, 0, 0, 1, 1, 1]


def test_appendix_b_writes(ctx):
    ops = [("ue", 0), ("ue", 1), ("ue", 2), ("ue", 255), ("ue", 65535), ("se", -3), ("se", 3), ("u", 32, 0xDEADBEEF)]
    out, bits, ovr = ctx.bs_write(ops, 16)
    assert out[:12] == H("A6 01 00 00 00 80 00 1C DB D5 B7 DD") and bits == 12 * 8 + 3 and ovr == 0
    # what does not fit is dropped, the cursor still advances (bs_write_u1 past the end)
    out, bits, ovr = ctx.bs_write([("u", 32, 0xFFFFFFFF), ("u", 16, 0xFFFF)], 4)
    assert out == b"\xff\xff\xff\xff" and bits == 48 and ovr == 1


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_scripts_against_the_oracle_reader(ctx, seed):
    rng = np.random.default_rng(seed)
    for it in range(60):
        n = int(rng.integers(1, 64))
        # zero-heavy bytes make long exp-Golomb prefixes; short buffers make the scripts run off the end
        data = rng.integers(0, 256, n).astype(np.uint8)
        data[rng.random(n) < [0.0, 0.5, 0.85][it % 3]] = 0
        ops = []
        for _ in range(int(rng.integers(1, 80))):
            k = int(rng.integers(0, 4))
            ops.append([("u", int(rng.integers(1, 33))), ("u1",), ("ue",), ("se",)][k])
        v, pos, ovr = ctx.bs_read(data.tobytes(), ops)
        want, byte, bits_left, eof, overrun = port.read_syntax(data.tobytes(), [(o[0], o[1]) if len(o) > 1 else (o[0],) for o in ops])
        want = [w - (1 << 32) if w >= (1 << 31) else w for w in want]
        assert v == want, (seed, it)
        assert pos[-1] == byte * 8 + (8 - bits_left) and ovr[-1] == overrun, (seed, it)


@pytest.mark.parametrize("seed", [4, 5])
def test_written_bits_read_back_by_the_oracle(ctx, seed):
    rng = np.random.default_rng(seed)
    for it in range(40):
        ops, rops, vals = [], [], []
        for _ in range(int(rng.integers(1, 60))):
            k = int(rng.integers(0, 4))
            if k == 0:
                nb = int(rng.integers(1, 33))
                val = int(rng.integers(0, 1 << nb))
                ops.append(("u", nb, val)); rops.append(("u", nb)); vals.append(val)
            elif k == 1:
                val = int(rng.integers(0, 2))
                ops.append(("u1", val)); rops.append(("u1",)); vals.append(val)
            elif k == 2:
                val = int(rng.integers(0, 65535))
                ops.append(("ue", val)); rops.append(("ue",)); vals.append(val)
            else:
                val = int(rng.integers(-30000, 30000))
                ops.append(("se", val)); rops.append(("se",)); vals.append(val)
        out, bits, ovr = ctx.bs_write(ops, 1024)
        assert ovr == 0
        got, byte, bits_left, eof, overrun = port.read_syntax(out[: (bits + 7) // 8], rops)
        assert got == vals, (seed, it)
        assert byte * 8 + (8 - bits_left) == bits
