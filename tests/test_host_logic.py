"""CPU tests of the host-side logic added in round 2 (no GPU, no compute calls): the node fingerprint, the page-locked
allocator's fallback, the slab partition arithmetic, and bench.py's byte model / config contract."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

import common
from nfft_b200 import cabi
from nfft_b200.dist import shard_range, slab_partition

sys.path.insert(0, common.ROOT)
import bench  # noqa: E402


def _fp(a):
    a = np.ascontiguousarray(a)
    return int(cabi.lib().nfftcu_fingerprint(C.c_void_p(a.ctypes.data), a.nbytes))


def test_fingerprint_detects_single_bit_and_order_changes():
    rng = np.random.default_rng(1)
    x = rng.random(3_000_001)            # 24 MB: the threaded path (more than four 1 MiB chunks)
    base = _fp(x)
    assert base == _fp(x.copy())
    for pos in (0, 1, 131071, 131072, 1_500_000, x.size - 1):
        y = x.copy()
        y.view(np.uint64)[pos] ^= np.uint64(1)          # flip the lowest mantissa bit
        assert _fp(y) != base
    y = x.copy()
    y[[5, 2_000_000]] = y[[2_000_000, 5]]               # the same values in different places (different chunks)
    assert _fp(y) != base
    y = x.copy()
    y[[7, 8]] = y[[8, 7]]                               # ... and within one 32-byte lane group
    assert _fp(y) != base
    assert _fp(x[:-1]) != base                          # length is part of the fingerprint
    assert _fp(np.zeros(0)) == _fp(np.zeros(0))
    small = rng.random(1000)                            # the single-threaded path
    assert _fp(small) == _fp(small.copy()) and _fp(small) != _fp(small[::-1].copy())


def test_fingerprint_does_not_depend_on_the_thread_count():
    code = ("import sys, ctypes as C, numpy as np; sys.path.insert(0, %r); from nfft_b200 import cabi;"
            "x = np.random.default_rng(2).random(2_500_000);"
            "print(int(cabi.lib().nfftcu_fingerprint(C.c_void_p(x.ctypes.data), x.nbytes)))" % common.ROOT)
    outs = set()
    for threads in ("1", "3", "8"):
        env = dict(os.environ, NFFT_B200_HASH_THREADS=threads)
        outs.add(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.strip())
    assert len(outs) == 1


def test_host_alloc_without_a_device_uses_the_heap_and_frees():
    L = cabi.lib()
    for nbytes in (1, 100, 1 << 18, 1 << 22):
        p = cabi.host_alloc(nbytes)
        assert p % 64 == 0
        (C.c_char * nbytes).from_address(p)[nbytes - 1] = b"x"
        cabi.host_free(p)
    L.nfftcu_pool_trim()


def test_slab_partition_with_a_given_order_is_a_partition_in_sorted_order():
    rng = np.random.default_rng(3)
    M = 1003
    order = rng.permutation(M)
    x = rng.random((M, 3)) - 0.5
    got = [slab_partition([8] * 3, [16] * 3, 2, x, r, 4, order=order) for r in range(4)]
    assert np.array_equal(np.concatenate(got), order)
    assert [len(g) for g in got] == [shard_range(M, r, 4)[1] - shard_range(M, r, 4)[0] for r in range(4)]


def test_byte_model_matches_survey_8d():
    """cfg3 fp64 A_trafo 2581.0 MB, A_adj 2346.2 MB, A_pair 4927.2 MB; B / B^T launch 668.4 MB."""
    bm = bench.byte_model(bench.CFGS["cfg3"], 16, 10_000_000)
    assert round(bm["trafo"] / 1e6, 1) == 2581.0
    assert round(bm["adjoint"] / 1e6, 1) == 2346.2
    assert round(bm["pair"] / 1e6, 1) == 4927.2
    assert round(bm["spread"] / 1e6, 1) == 668.4 and bm["spread"] == bm["interp"]
    bm4 = bench.byte_model(bench.CFGS["cfg4"], 16, 100_000_000)
    assert round(bm4["pair"] / 1e6, 1) == 41017.6


def test_both_bench_arms_describe_the_same_workload():
    cfg = bench.CFGS["cfg3"]
    a = bench.workload_config(cfg, 1, "weak")
    assert a == bench.workload_config(dict(cfg), 1, "weak")
    assert "M=10000000" in a["workload"] and "N=128^3" in a["workload"]
    assert bench.metric_name(cfg, "double") == bench.METRIC
    assert bench.METRIC.endswith("rel l2 err")
    x, fh, f = bench.synth(dict(cfg, M=1000), "double", 0)
    x2, fh2, f2 = bench.synth(dict(cfg, M=1000), "double", 0)
    assert np.array_equal(x, x2) and np.array_equal(f, f2) and x.min() >= -0.5 and x.max() < 0.5
    xf, _, ff = bench.synth(dict(cfg, M=1000), "float", 0)
    assert xf.dtype == np.float32 and xf.max() < 0.5 and abs(float(ff.mean())) < 0.05   # zero-mean fp32 data
