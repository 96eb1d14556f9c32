"""Guard of the measured ptxas schedule of the headline kernels (CPU test; needs the built csrc/build/accel.o and cuobjdump).

The packed fp32x2 pair loop is limited by register-file reads of the three-operand FFMA2s, and ptxas' instruction order /
register assignment decides how many of them hit the operand-reuse cache: builds whose loops contain the same 116
instructions differ by 2-4 % in run time, and an unrelated edit elsewhere in accel.cu can flip the schedule.  The numbers in
profiles/r02 belong to the schedule fingerprinted in profiles/r02/hot_loop_fingerprint.json; this test fails when the built
library no longer has it, so that the change is noticed, re-measured on a B200 (tools/time_splits.py) and the fingerprint
refreshed with `python tools/sass_loops.py --write`."""
import json
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_headline_kernel_schedule_is_the_measured_one():
    import sass_loops
    if not os.path.exists(sass_loops.OBJ) or shutil.which("cuobjdump") is None:
        pytest.skip("csrc/build/accel.o or cuobjdump not available (run __graft_entry__.build() first)")
    want = json.load(open(sass_loops.OUT))["loops"]
    got = sass_loops.hot_loops()
    for key, rec in want.items():
        assert key in got, f"kernel {rec['what']} is no longer built"
        assert got[key]["instructions"] == rec["instructions"], (rec["what"], got[key]["instructions"], rec["instructions"])
        assert got[key]["sha256"] == rec["sha256"], (
            f"the ptxas schedule of '{rec['what']}' changed (reuse flags {got[key]['reuse_flags']} vs {rec['reuse_flags']}): "
            f"re-measure on a B200 and refresh profiles/r02/hot_loop_fingerprint.json (python tools/sass_loops.py --write)")
