"""Copy the reference's Python sources UNMODIFIED from /root/reference into baseline/_ref/ (git-ignored, NOT gpurun-ignored, so it
travels to the GPU box) — the place `bench.py --impl reference`, `tests/test_reference_entrypoints.py` and INTEGRATION.md
expect the reference at.  The reference has no setup.py / pyproject.toml, so
`pip install --no-index --target baseline/_ref /root/reference` fails ("neither 'setup.py' nor 'pyproject.toml' found");
a plain copy of its sources, lists and config is the documented fallback.  Nothing is copied into the tracked tree."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
KEEP = (".py", ".yaml", ".txt", ".md")


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"{SRC} is absent (GPU box): keeping the prebuilt {DST}" if os.path.isdir(DST) else f"{SRC} is absent", file=sys.stderr)
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for d, dirs, files in os.walk(SRC):
        dirs[:] = [x for x in dirs if x not in ("__pycache__", ".git", "artifact_distibution", "pretrained_weights")]
        for f in files:
            if f.endswith(KEEP):
                rel = os.path.relpath(os.path.join(d, f), SRC)
                os.makedirs(os.path.dirname(os.path.join(DST, rel)) or DST, exist_ok=True)
                shutil.copy2(os.path.join(d, f), os.path.join(DST, rel))
                n += 1
    print(f"reference: {n} files -> {DST}", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
