"""Innermost pair loops of the force kernels from `cuobjdump -sass` (used by tests/test_sass_fingerprint.py and to refresh
profiles/r02/hot_loop_fingerprint.json:  python tools/sass_loops.py --write)."""
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "nbody_cosmological_simulation_b200", "csrc", "build", "accel.o")
OUT = os.path.join(ROOT, "profiles", "r02", "hot_loop_fingerprint.json")
KERNELS = {   # mangled-name fragment -> description
    "accel_kernelINS_8ForceF32ILi3ELi0ELi2ELi256ELb1ELi4ELb0EEELi0E": "fp32 FLOAT32 D=3 uniform masses (benchmark kernel)",
    "accel_kernelINS_8ForceF32ILi2ELi0ELi2ELi256ELb1ELi4ELb0EEELi0E": "fp32 FLOAT32 D=2 uniform masses (disk galaxies)",
    "accel_kernelINS_8ForceF64ILi3ELi4ELi2ELi256ELb1ELi2ELb0EEELi0E": "fp64 FLOAT64 D=3 uniform masses",
}


def hot_loops(obj=OBJ):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out = {}
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name = f.split("\n", 1)[0].strip()
        key = [k for k in KERNELS if k in name]
        if not key:
            continue
        lines = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", ln).rstrip() for ln in f.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
        addr = lambda ln: int(re.search(r"/\*([0-9a-f]{4})\*/", ln).group(1), 16)  # noqa: E731
        best = None
        for i, ln in enumerate(lines):
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", ln)
            if m and int(m.group(1), 16) < addr(ln):
                j = [k for k, x in enumerate(lines) if addr(x) == int(m.group(1), 16)]
                if j:
                    body = lines[j[0]: i + 1]
                    if sum("MUFU" in x for x in body) >= 4 and (best is None or len(body) < len(best)):
                        best = body
        text = "\n".join(re.sub(r"/\*[0-9a-f]{4}\*/\s*", "", ln).strip() for ln in best)
        out[key[0]] = {"what": KERNELS[key[0]], "instructions": len(best), "reuse_flags": text.count(".reuse"),
                       "sha256": hashlib.sha256(text.encode()).hexdigest()}
    return out


if __name__ == "__main__":
    loops = hot_loops()
    print(json.dumps(loops, indent=1))
    if "--write" in sys.argv:
        note = ("ptxas schedule of the innermost pair loops of the build that was MEASURED (profiles/r02/README.md).  The packed "
                "fp32x2 loop is register-read limited; otherwise equivalent schedules differ by 2-4 % (388 vs 397 vs 406 ms at "
                "N = 2^20), and unrelated edits to accel.cu change the schedule.  If this fingerprint changes, re-measure "
                "(tools/time_splits.py on a B200) before updating it.")
        json.dump({"note": note, "loops": loops}, open(OUT, "w"), indent=1)
        print("wrote", OUT)
