"""Install the UNMODIFIED reference package for the benchmark's reference arm (build container only).

TEST / MEASUREMENT INFRASTRUCTURE.  The reference is pure Python without packaging metadata (no setup.py / pyproject),
so `pip install --target baseline/_ref /root/reference` has nothing to build; this recipe does what that install would:
it places the reference's own `src/` package files, byte for byte, under `baseline/_ref/` -- a directory that is
git-ignored (never part of the history) but travels to the GPU box with the working tree, like a built `.so`.
`bench.py --impl reference` imports the live classes from there (two import stubs for `librosa.filters` / `einx`, see
oracle/ref_loader.py) and times the reference's own `decoder(units)` on the box's host cores.  Only the files the
unit-to-speech path imports are installed (src/flow_matching/{__init__,configs,models}.py + modules/, src/hifigan/
{__init__,data}.py).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")

FILES = [
    "src/flow_matching/__init__.py",
    "src/flow_matching/configs.py",
    "src/flow_matching/models.py",
    "src/flow_matching/modules/__init__.py",
    "src/flow_matching/modules/fourier_embed.py",
    "src/flow_matching/modules/norm.py",
    "src/flow_matching/modules/transformer.py",
    "src/flow_matching/modules/fastspeech/__init__.py",
    "src/flow_matching/modules/fastspeech/modules.py",
    "src/hifigan/__init__.py",
    "src/hifigan/data.py",
    "LICENSE",
]


def install() -> bool:
    """Returns True when baseline/_ref holds the reference (installed now or before); False when /root/reference is absent
    and nothing was installed earlier."""
    if not os.path.isdir(os.path.join(SRC, "src", "flow_matching")):
        return os.path.exists(os.path.join(DST, "INSTALLED.json"))
    digest = {}
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.exists(src):
            if rel.endswith("__init__.py"):
                # namespace-style directory in the reference: an empty marker keeps the import path a regular package
                os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                open(os.path.join(DST, rel), "a").close()
                continue
            raise FileNotFoundError(src)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        digest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    if not os.path.exists(os.path.join(DST, "src", "__init__.py")) and os.path.exists(os.path.join(SRC, "src", "__init__.py")):
        shutil.copyfile(os.path.join(SRC, "src", "__init__.py"), os.path.join(DST, "src", "__init__.py"))
    with open(os.path.join(DST, "INSTALLED.json"), "w") as f:
        json.dump({"source": SRC, "files": digest}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = install()
    print("baseline/_ref:", "installed" if ok else "reference tree absent, nothing installed")
    sys.exit(0 if ok else 1)
