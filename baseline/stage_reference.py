"""Stages the files of the reference that the measure stage needs into baseline/_ref/ (git-ignored, but shipped to the
GPU box by gpurun), so that GPU-side tests and bench.py's CPU arm can run the reference's OWN code there.

    python baseline/stage_reference.py            (build container: copies from /root/reference)

Nothing is committed: baseline/_ref/ is in .gitignore ("the reference install is not product source").  The
reference is a Python application, not a package -- there is nothing to pip-install (no setup.py / pyproject), and
`ultralytics` (requirements.txt:13, unpinned) is not in the offline wheelhouse, so the files are staged as they are and
imported with the stub recipe of oracle/ref_verbatim.py."""
import os
import shutil
import sys

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["measurement.py", "config.py", "hardware_utils.py", "Utils/check_stitch_distance.py", "main.py",
         "camera_calibration.json", "extrinsics.json", "camera_extrinsics.json"]


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"{SRC} not present: nothing staged (the GPU box uses what was staged in the build container)")
        return 0
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        if not os.path.exists(s):
            print("missing upstream:", f)
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
    print("staged", len(FILES), "files into", DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())
