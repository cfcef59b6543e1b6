"""Stages the UNMODIFIED reference package into baseline/_ref/ (git-ignored, travels to the GPU box with gpurun) so
that scripts/bench_vs_reference_gpu.py can run the reference's own Triton kernels next to ours on the same B200.

`pip install --no-index --target baseline/_ref /root/reference` fails here (the build backend, hatchling, is not in
the image), and the package is pure Python, so this does by hand what the wheel would: the package directory plus a
dist-info with the metadata `msda_triton/__init__.py` asks importlib.metadata for.  Nothing under baseline/_ref is
imported by the product, the tests or bench.py.  The reference's tests are staged next to it: they import
`msda_triton.frontend`, so with `PYTHONPATH=msda-triton_b200` they exercise THIS package through the reference's own
assertions (the drop-in check of INTEGRATION.md).

    python scripts/stage_reference.py            # needs /root/reference (this container only)
"""
import re
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference")
DST = ROOT / "baseline" / "_ref"


def main():
    if not (SRC / "src" / "msda_triton").is_dir():
        print("no /root/reference here: nothing staged")
        return 1
    version = re.search(r'^version\s*=\s*"([^"]+)"', (SRC / "pyproject.toml").read_text(), re.M).group(1)
    if DST.exists():
        shutil.rmtree(DST)
    DST.mkdir(parents=True)
    shutil.copytree(SRC / "src" / "msda_triton", DST / "msda_triton")
    # the reference's own test-suite, to be run UNMODIFIED against this repository's package on the GPU box:
    #   PYTHONPATH=msda-triton_b200 python -m pytest baseline/_ref/reference_tests -q
    shutil.copytree(SRC / "tests", DST / "reference_tests")
    info = DST / f"msda_triton-{version}.dist-info"
    info.mkdir()
    (info / "METADATA").write_text(f"Metadata-Version: 2.1\nName: msda_triton\nVersion: {version}\n")
    (info / "INSTALLER").write_text("scripts/stage_reference.py\n")
    print(f"staged reference msda_triton {version} -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
