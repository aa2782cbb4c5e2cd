#!/usr/bin/env python3
"""Stage the files of the REFERENCE that the end-to-end oracle executes into baseline/_ref/  (test infrastructure).

/root/reference does not exist on the GPU box; only /root/repo travels there.  baseline/_ref/ is git-ignored (no
reference source ever enters the history) but NOT gpurun-ignored, so an unmodified copy of the reference's own
`utils/common.py` (process_kenburns / process_inpaint / process_shift / render_pointcloud / fill_disocclusion /
generate_mask and their CUDA kernel strings), `utils/pipeline.py`, `utils/utils.py`, `utils/partial_conv.py`,
`utils/helper_math.h` and `models/*.py` can be imported there behind oracle/refshim.py and run on the B200 as the ground
truth of tests/test_gpu_reference_e2e.py.

Files are copied byte for byte (checked by sha1 in the manifest); nothing is patched.

Run:  python oracle/stage_ref.py        (needs /root/reference; a no-op that keeps the staged copy when it is absent)
"""
import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("KB_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "baseline", "_ref")

FILES = [
    "utils/common.py", "utils/pipeline.py", "utils/utils.py", "utils/partial_conv.py", "utils/helper_math.h",
    "models/disparity_estimation.py", "models/disparity_refinement.py", "models/disparity_refinement_pretrained.py",
    "models/pointcloud_inpainting.py", "models/partial_inpainting.py",
]


def main():
    if not os.path.isdir(REF):
        print(f"[stage_ref] {REF} absent: keeping whatever baseline/_ref/ already holds")
        return 0
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha1(f.read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha1": manifest}, f, indent=1)
    print(f"[stage_ref] staged {len(FILES)} reference files into {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
