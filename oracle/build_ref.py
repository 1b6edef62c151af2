"""Recipe for oracle/_ref/: the reference's OWN implementation of the hot path, taken unmodified
from the read-only checkout so that it can travel to the GPU box (which has no /root/reference).
*** TEST / BASELINE INFRASTRUCTURE ONLY ***

The reference is pure Python: there is nothing to compile; "building" it
means placing the three modules of the hot path where oracle/ref_shim.py can import them:
    models/dynamic_adapter.py  models/vision_transformer_IN21K.py  models/model_speed_test.py
oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but not
gpurun-ignored.  Run by __graft_entry__.build() whenever /root/reference is present; bench.py's
`--impl reference` arm and `cpu_baseline` leg time these files through the import shim
(`kind: "reference"`); without them they fall back to the oracle port (`kind: "port"`).

    python oracle/build_ref.py [--src /root/reference]
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ("models/dynamic_adapter.py", "models/vision_transformer_IN21K.py", "models/model_speed_test.py")


def build(src: str = "/root/reference") -> bool:
    if not all(os.path.isfile(os.path.join(src, f)) for f in FILES):
        return False
    manifest = {"source": src, "files": {}}
    head = os.path.join(src, ".git", "HEAD")
    for f in FILES:
        out = os.path.join(DST, f)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src, f), out)
        manifest["files"][f] = hashlib.sha256(open(out, "rb").read()).hexdigest()
    init = os.path.join(src, "models", "__init__.py")
    if os.path.isfile(init):
        shutil.copyfile(init, os.path.join(DST, "models", "__init__.py"))
    if os.path.isfile(head):
        manifest["git_head"] = open(head).read().strip()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    return True


if __name__ == "__main__":
    src = sys.argv[2] if len(sys.argv) > 2 and sys.argv[1] == "--src" else "/root/reference"
    ok = build(src)
    print("oracle/_ref:", "built from " + src if ok else "reference checkout not found, nothing done")
