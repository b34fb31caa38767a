"""Stage the handful of reference modules that tests/test_gpu_boundary.py drives into baseline/_ref
(git-ignored, never committed; it travels to the GPU box with the gpurun snapshot because
/root/reference does not exist there).  `--clean` removes the copy again.

  python tools/stage_reference.py [--clean]
"""
import shutil
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
DST = REPO / "baseline" / "_ref"
FILES = ["trainer.py", "world.py", "parse.py", "utils.py", "metric.py", "negative_sample.py", "dataloader.py"]

if "--clean" in sys.argv:
    shutil.rmtree(DST, ignore_errors=True)
    print(f"removed {DST}")
elif not REF.exists():
    print(f"{REF} is not present: nothing staged")
else:
    DST.mkdir(parents=True, exist_ok=True)
    for f in FILES:
        shutil.copy2(REF / f, DST / f)
    print(f"staged {len(FILES)} unmodified reference modules into {DST}")
