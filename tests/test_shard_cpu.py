"""CPU tests of the byte-range sharding: cut planning + stitch (the product's host code, libhevcb200.so) over shard
records produced by the TEST-ONLY host build of the kernel logic (tests/hostsim); the stitched result must equal the
oracle's whole-stream result.  The world_size-2 test runs the record exchange over gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ref
from tests import shard_check, util

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("alphabet", [0, 1, 2, 3])
def test_adversarial_small_streams(hostsim, alphabet):
    run = shard_check.hostsim_shard_runner(hostsim)
    rng = np.random.default_rng(300 + alphabet)
    for it in range(600):
        size = int(rng.integers(0, 400))
        buf = util.adversarial(rng, size, alphabet, density=[1.0, 0.5, 0.1][it % 3])
        for g in (1, 2, 3, 5):
            shard_check.check_sharded(buf, size, g, run, tag=f"a{alphabet}-{it}")


def test_every_cut_position(hostsim):
    """two shards, the cut at every legal position of short streams (any p with b[p-1] >= 2 and a last shard >= 64 B)"""
    run = shard_check.hostsim_shard_runner(hostsim)
    rng = np.random.default_rng(77)
    n_cuts = 0
    for it in range(60):
        size = int(rng.integers(130, 260))
        buf = util.adversarial(rng, size, it, density=0.6)
        for p in range(1, size - 64):
            if buf[p - 1] >= 2:
                shard_check.check_sharded(buf, size, 2, run, tag=f"cut{it}@{p}", bounds=np.array([0, p, size], np.int64))
                n_cuts += 1
    assert n_cuts > 1000


def test_nal_spanning_several_shards(hostsim):
    """one NAL that covers whole shards (no event in them), with and without a nal_to_rbsp error in the middle"""
    run = shard_check.hostsim_shard_runner(hostsim)
    rng = np.random.default_rng(5)
    for err_at in (None, 700, 1500):
        body = rng.integers(4, 256, 3000, dtype=np.uint8)
        body[100:103] = [0, 0, 3]  # a removable EPB in the first shard
        body[1800:1803] = [0, 0, 3]
        if err_at is not None:
            body[err_at:err_at + 3] = [0, 0, 2]
        s = np.concatenate([np.array([0, 0, 1, 0x40, 1], np.uint8), body, np.array([0, 0, 1, 0x42, 1, 9, 9, 0x80], np.uint8),
                            rng.integers(4, 256, 200, dtype=np.uint8)])
        buf = util.padded(s)
        for g in (2, 4, 6, 8):
            shard_check.check_sharded(buf, s.size, g, run, tag=f"span-{err_at}")


@pytest.mark.parametrize("seed", [1, 2])
def test_generated_streams(hostsim, seed):
    run = shard_check.hostsim_shard_runner(hostsim)
    s = ref.gen_stream(seed=seed, profile=1, n_slices=800, payload_min=1, payload_max=300, zero_heavy_pct=30, extra_zero_pct=20, ps_period=40,
                       unsupported_pct=5)
    size = s.size - ref.PAD
    for g in (2, 3, 8):
        for cut in (0, 3, 7):
            buf = util.padded(s[: size - cut])
            n = shard_check.check_sharded(buf, size - cut, g, run, tag=f"gen{seed}-{cut}")
            assert n > 800


def test_zero_length_nal_stops_the_loop(hostsim):
    """00 00 01 00 00 01: find_nal_unit returns 0 and the reference loop ends there, wherever the cuts fall"""
    run = shard_check.hostsim_shard_runner(hostsim)
    rng = np.random.default_rng(9)
    for pos in (300, 640, 1000):
        a = rng.integers(4, 256, 1400, dtype=np.uint8)
        for k in range(0, 1300, 100):
            a[k:k + 4] = [0, 0, 1, 0x40]
        a[pos:pos + 6] = [0, 0, 1, 0, 0, 1]
        buf = util.padded(a)
        for g in (2, 3, 4, 7):
            shard_check.check_sharded(buf, a.size, g, run, tag=f"stop{pos}")


def test_plan_shards_degenerate():
    from hevcbitstream_b200 import shard as hs

    z = np.zeros(5000, np.uint8)  # no legal cut anywhere: everything ends up in the first shard
    b = hs.plan_shards(z, 4)
    assert b.tolist() == [0, 5000, 5000, 5000, 5000]
    e = hs.plan_shards(np.zeros(0, np.uint8), 3)
    assert e.tolist() == [0, 0, 0, 0]
    r = np.full(1000, 7, np.uint8)
    b = hs.plan_shards(r, 4)
    assert b.tolist() == [0, 250, 500, 750, 1000]
    b = hs.plan_shards(r[:100], 4)  # last shard must keep >= 64 bytes
    assert b[-1] == 100 and (100 - b[b < 100].max()) >= 64


def test_world_size_2_gloo():
    """two processes, gloo: each scans its shard (host build of the kernel logic), all_gather of the records, stitch"""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           "29533", os.path.join(ROOT, "tests", "shard_gloo_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "SHARD_GLOO_OK" in out.stdout
