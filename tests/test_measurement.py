"""The ncu numbers bench.py reports (roofline.traffic, roofline_secondary) must belong to the kernels that are benched: the capture under
profiles/ is stamped with the identity of the stock kernels' machine code (sha256 of their SASS, __graft_entry__.sass_id), and the library in
the tree must still be that build.  Changing a stock kernel without taking a new capture (tools/profile_bench.sh + tools/ncu_extract.py)
fails here instead of silently reporting `traffic: null` at round end."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_capture_belongs_to_the_built_kernels():
    import __graft_entry__ as ge

    if not os.path.exists(ge.CUBIN):
        pytest.skip("kernels not built yet (python -c 'import __graft_entry__ as g; g.build()')")
    bid = ge.sass_id()
    if bid is None:
        pytest.skip("neither cuobjdump nor the SASS id file is available")
    prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert prof["build_id"] == bid, "profiles/traffic.json is a capture of build %s, the tree holds %s: take a new capture" % (prof["build_id"], bid)
    assert prof["kernel"] == "fe_env_step_kernel" and prof["dram_bytes_per_launch"] > 0
    import bench

    assert bench.build_id() == bid
