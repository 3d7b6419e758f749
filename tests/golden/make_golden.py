#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ (run HERE, in the build container).

Needs /root/reference (read-only) and the reference build oracle/_ref (``make -C oracle``).
Neither exists on the GPU box; the .npz files written by this script are what travels.

Outputs
-------
ndft_fixtures.npz   the reference's own known-answer files tests/data/nfft_*.txt and
                    nfft_adjoint_*.txt (exact NDFT computed by Mathematica at 64 digits,
                    tests/check_nfft.m:23-55), re-encoded as float64 arrays.  Keys
                    "<name>/N", "/x", "/f_hat", "/f" (+ "/kind" = trafo|adjoint).
bessel_i0.npz       the I0(j), j = 0..99 table of tests/bessel.c:29-133, float64.
ref_outputs.npz     outputs of the REFERENCE ITSELF (oracle/_ref/libnfft3_ref.so and the
                    single-precision twin) on small seeded inputs: f, f_hat and the
                    NFFT_SORT_NODES permutation index_x, for 1/2/3-D, fp64 and fp32.
                    Inputs are regenerated from the seed by tests/common.py::make_case.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


def parse_fixture(path: str, adjoint: bool):
    tok = open(path).read().split()
    it = iter(tok)
    d = int(next(it))
    N = [int(next(it)) for _ in range(d)]
    M = int(next(it))
    x = np.array([float(next(it)) for _ in range(M * d)], dtype=np.float64).reshape(M, d)
    NN = int(np.prod(N))
    # one reader for both kinds (tests/nfft.c:381-448 setup_file): x, then f_hat[N_total], then
    # f[M]; in nfft_adjoint_* files f is the input and f_hat the exact adjoint NDFT of it.
    fh = np.array([float(next(it)) for _ in range(2 * NN)]).view(np.complex128)
    f = np.array([float(next(it)) for _ in range(2 * M)]).view(np.complex128)
    return np.array(N, dtype=np.int64), x, fh, f


def make_ndft_fixtures():
    out = {}
    for path in sorted(glob.glob(f"{REF}/tests/data/nfft_*.txt")):
        name = os.path.basename(path)[:-4]
        adj = name.startswith("nfft_adjoint_")
        N, x, fh, f = parse_fixture(path, adj)
        out[f"{name}/N"] = N
        out[f"{name}/x"] = x
        out[f"{name}/f_hat"] = fh
        out[f"{name}/f"] = f
        out[f"{name}/kind"] = np.array("adjoint" if adj else "trafo")
    np.savez_compressed(os.path.join(HERE, "ndft_fixtures.npz"), **out)
    print("ndft_fixtures.npz:", len(out) // 5, "cases")


def make_bessel():
    src = open(f"{REF}/tests/bessel.c").read()
    body = src[src.index("static const R r[]"):src.index("};")]
    vals = [float(v) for v in re.findall(r"K\(([0-9.eE+-]+)\)", body)]
    assert len(vals) == 100 and vals[0] == 1.0, len(vals)
    np.savez_compressed(os.path.join(HERE, "bessel_i0.npz"), i0=np.array(vals, dtype=np.float64))
    print("bessel_i0.npz:", len(vals), "values")


def make_ref_outputs():
    from common import REF_CASES, make_case, ref_api  # noqa: E402
    from nfft_b200.plan import Plan
    out = {}
    for name, spec in REF_CASES.items():
        for prec in ("double", "float"):
            api = ref_api(prec)
            x, fh, f = make_case(spec, prec)
            p = Plan.init_guru(spec["d"], spec["N"], spec["M"], spec["n"], spec["m"], spec["flags"],
                               api=api)
            p.x[:] = x
            p.precompute_one_psi()
            p.f_hat[:] = fh
            p.trafo()
            out[f"{name}/{prec}/f"] = p.f.copy()
            p.f[:] = f
            p.adjoint()
            out[f"{name}/{prec}/f_hat"] = p.f_hat.copy()
            out[f"{name}/{prec}/index_x"] = p.index_x[:, 1].astype(np.int32).copy()
            p.finalize()
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
    print("ref_outputs.npz:", len(out), "arrays")


if __name__ == "__main__":
    make_ndft_fixtures()
    make_bessel()
    make_ref_outputs()
