"""TEST INFRASTRUCTURE — recipe that stages the UNMODIFIED reference next to the oracle so that it travels to the GPU box.

    python -m oracle.stage_ref            # /root/reference -> oracle/_ref/   (git-ignored, NOT gpurun-ignored)

`/root/reference` exists only in the build container.  The reference is pure Python (no native code, no setup.py /
pyproject: `pip install --target baseline/_ref /root/reference` fails with "neither setup.py nor pyproject.toml found",
recorded in DESIGN.md §6), so "building" it means placing its importable tree where `oracle/ref_harness.py` finds it:
byte-identical copies of `modeling/`, `data/`, `g2vlm_utils.py`, `inference_recon.py` and the first 8 sorted frames of
`examples/dl3dv` (BASELINE configs[0]) go to `oracle/_ref/`, together with `MANIFEST.json` (sha256 per file, so a test can
show that what ran on the B200 is the unmodified source).  Nothing under `oracle/_ref/` is tracked by git; the product
package never imports it (only tests/, smoke() and bench.py's reference legs do, through ref_harness).
`__graft_entry__.build()` calls `stage()` whenever `/root/reference` is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("G2VLM_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
TREES = ("modeling", "data")
FILES = ("g2vlm_utils.py", "inference_recon.py", "inference_chat.py", "LICENSE")
EXAMPLE_DIR, EXAMPLE_FRAMES = os.path.join("examples", "dl3dv"), 8


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def _wanted():
    """Relative paths of every file that is staged."""
    rel = []
    for t in TREES:
        for d, _, fs in os.walk(os.path.join(SRC, t)):
            if "__pycache__" in d:
                continue
            rel += [os.path.relpath(os.path.join(d, f), SRC) for f in fs if not f.endswith(".pyc")]
    rel += [f for f in FILES if os.path.exists(os.path.join(SRC, f))]
    ex = os.path.join(SRC, EXAMPLE_DIR)
    if os.path.isdir(ex):
        rel += [os.path.join(EXAMPLE_DIR, f) for f in sorted(os.listdir(ex))[:EXAMPLE_FRAMES]]
    return sorted(rel)


def stage(force: bool = False) -> str | None:
    """Copies the reference tree to oracle/_ref (no-op when it is already there with matching hashes).  Returns the
    destination, or None when /root/reference does not exist (GPU box: the staged copy from the snapshot is used)."""
    if not os.path.isdir(os.path.join(SRC, "modeling", "g2vlm")):
        return None
    manifest_path = os.path.join(DST, "MANIFEST.json")
    rel = _wanted()
    if not force and os.path.exists(manifest_path):
        old = json.load(open(manifest_path))
        if sorted(old["files"]) == rel and all(os.path.exists(os.path.join(DST, r)) for r in rel):
            return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    files = {}
    for r in rel:
        dst = os.path.join(DST, r)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, r), dst)
        files[r] = _sha(dst)
    json.dump(dict(source=SRC, files=files), open(manifest_path, "w"), indent=0, sort_keys=True)
    return DST


def verify() -> int:
    """Re-hashes the staged files against MANIFEST.json; returns the number of files checked (raises on a mismatch)."""
    man = json.load(open(os.path.join(DST, "MANIFEST.json")))
    for r, h in man["files"].items():
        if _sha(os.path.join(DST, r)) != h:
            raise RuntimeError(f"staged reference file differs from the manifest: {r}")
    return len(man["files"])


if __name__ == "__main__":
    d = stage(force=True)
    print(f"staged {verify()} files under {d}" if d else f"{SRC} not present: nothing staged")
