"""Stage the reference's own CALLERS of the env next to the repo, unmodified, for tests/test_gpu_reference_callers.py.

`/root/reference` does not exist on the GPU box, the CUDA env does not run in the build container: the two meet through
`baseline/_ref/` (git-ignored, travels with the gpurun snapshot -- the place the task reserves for the unmodified
reference).  Copies agents/imitation_runners.py (traj_segment_generator) and run.py (test loop) byte for byte; nothing
from there is imported by the product.

    python scripts/stage_reference_callers.py [/root/reference]
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
dst = os.path.join(ROOT, "baseline", "_ref", "OpenRoboRL")
for rel in ("agents/imitation_runners.py", "run.py"):
    src = os.path.join(ref, "OpenRoboRL", rel)
    out = os.path.join(dst, rel)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    shutil.copyfile(src, out)
    print("staged", out)
