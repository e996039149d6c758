"""GPU: the two checking builds of the kernel sources (hybridmc_b200/build.py VARIANTS).

* verify — the single-precision partner pre-filter is the one place where bit-exactness
  rests on an error-bound argument (tests/test_prefilter_margins.py): this build predicts
  every pair the filter would drop anyway and raises 0x40000000 if one of them has an event.
* bounds — every index into the shared-memory replica image is range-checked (the
  substitute for compute-sanitizer, which the GPU pool does not allow)."""
import json
import os
import subprocess
import sys

import pytest

from hybridmc_b200.build import VARIANTS, build_variants

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
PREFILTER_VIOLATION = 0x40000000


def _run(variant):
    lib = VARIANTS[variant][1]
    if not os.path.exists(lib):
        build_variants()
    r = subprocess.run([sys.executable, os.path.join(HERE, "gpu_debug_worker.py")],
                       env=dict(os.environ, HMC_LIB=lib), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_prefilter_never_drops_a_pair_with_an_event():
    out = _run("verify")
    assert out["events"] > 5e7
    assert out["err_or"] & PREFILTER_VIOLATION == 0
    assert out["err_or"] & ~32 == 0, hex(out["err_or"])
    assert out["bounds_violations"] == -1  # not the bounds build


def test_no_out_of_range_index_into_the_replica_image():
    out = _run("bounds")
    assert out["events"] > 5e7
    assert out["bounds_violations"] == 0
    assert out["err_or"] & ~32 == 0, hex(out["err_or"])


def test_production_library_is_not_a_checking_build():
    from hybridmc_b200.engine import load_library
    assert load_library().hmc_debug_bounds_violations() == -1
