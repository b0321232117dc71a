"""Recipe for oracle/_ref/: the reference's own Python files for the hot path, placed where bench.py's
reference arm can import them on the GPU box (where /root/reference does not exist).

    python oracle/make_ref.py            (dev container only; needs /root/reference)

TEST / BENCH INFRASTRUCTURE.  Nothing is copied into the repository's history: oracle/_ref/ is listed in
.gitignore (and NOT in .gpurunignore, so it travels with gpurun like the built .so files).  The reference is
pure Python + numba -- there is nothing to compile; "building" it means placing the files unmodified:

    src/overflow/__init__.py, constants.py, flow_direction.py, flow_accumulation.py, util/raster.py

plus a three-line stub `osgeo` package next to them (flow_direction.py:4 and util/raster.py:2 import GDAL at
module top; the tile kernels never touch it, and this image has no GDAL).  bench.py --impl reference then times
the reference's own `flow_direction_for_tile` (numba prange) for the direction half.  Its accumulation
(`single_tile_flow_accumulation`) is O(N^2) in the queue (list.pop(0), flow_accumulation.py:136: 234 s at
1024^2), so that half stays on the C port with the O(1) FIFO -- a FASTER baseline than the reference.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/overflow"
DST = os.path.join(HERE, "_ref")
FILES = ["__init__.py", "constants.py", "flow_direction.py", "flow_accumulation.py", os.path.join("util", "raster.py")]


def make(quiet=False):
    if not os.path.isdir(SRC):
        if not quiet:
            print(f"{SRC} not found: oracle/_ref is only made in the dev container", file=sys.stderr)
        return False
    for rel in FILES:
        out = os.path.join(DST, "overflow", rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), out)
    stub = os.path.join(DST, "osgeo")
    os.makedirs(stub, exist_ok=True)
    with open(os.path.join(stub, "__init__.py"), "w") as f:
        f.write("# stub: the reference imports GDAL at module top; the tile kernels never use it\nfrom . import gdal  # noqa: F401\n")
    with open(os.path.join(stub, "gdal.py"), "w") as f:
        f.write("def UseExceptions():\n    pass\n\n\nclass Band:\n    pass\n")
    return True


if __name__ == "__main__":
    print("oracle/_ref ready" if make() else "oracle/_ref not made")
