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


def test_ik_step_restates_the_stock_step_blocks_verbatim():
    """csrc/fe_ik.h carries its own copy of the control-mapping block and of everything that follows the simulation in fe_env_step_one
    (fe_env.h is frozen as the profiled build, see fe_ik.h's header).  The copies may differ from the original only where they say so:
    the policy action's length is a parameter, and the reset after a failed simulation is decided by the caller."""
    csrc = os.path.join(ROOT, "furniture_b200", "csrc")
    env = open(os.path.join(csrc, "fe_env.h")).read()
    ik = open(os.path.join(csrc, "fe_ik.h")).read()
    body = env[env.index("FE_FN void fe_env_step_one("):env.index("// per-env context set-up shared by the CUDA kernels")]
    ctrl = body[body.index("  LANES_BEGIN\n    for (int u = lane; u < m->nu; u += 32) { // _setup_action"):body.index("  for (int i = 0; i < cfg->nsub; ++i) fe_substep_lockstep(w); // _do_simulation")]
    after = body[body.index("  FE_SYNC;\n  if (fail) {"):body.rindex("}")]
    assert ctrl in ik, "fe_ik_controls no longer equals the _setup_action block of fe_env_step_one"
    want = after.replace("sc->act_dim", "act_dim").replace("  FE_SYNC;\n  if (fail) {", "  FE_SYNC;\n  if (reset_now) {", 1)
    assert want in ik, "fe_ik_finish no longer equals the post-simulation part of fe_env_step_one"
