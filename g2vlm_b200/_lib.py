"""Loader / builder for the C-ABI kernel library ``libg2vlm_b200.so``.

The library is built IN-TREE with nvcc for sm_100a (``build()``; also called by
``__graft_entry__.build``) and loaded with ctypes.  There is deliberately no fallback: if the
shared object is missing or a symbol declared in ``include/g2vlm_b200.h`` cannot be resolved the
import of the op layer fails loudly (the product path must never silently run on the oracle or on
plain PyTorch).
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
HEADER = REPO_ROOT / "include" / "g2vlm_b200.h"
LIB_PATH = Path(os.environ.get("G2VLM_B200_LIB", PKG_DIR / "libg2vlm_b200.so"))

NVCC_FLAGS = [
    "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
]
OBJ_DIR = PKG_DIR / "build"


def sources():
    return sorted(CSRC.glob("*.cu"))


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HEADER]
    return any(d.stat().st_mtime > t for d in deps)


def _header_abi_version() -> int:
    m = re.search(r"#define\s+G2VLM_ABI_VERSION\s+(\d+)", HEADER.read_text())
    if not m:
        raise RuntimeError(f"G2VLM_ABI_VERSION not found in {HEADER}")
    return int(m.group(1))


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source of the package for sm_100a into ``libg2vlm_b200.so``: one object per source
    (only the out-of-date ones, in parallel), then one link."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    OBJ_DIR.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + [HEADER]
    hdr_t = max(h.stat().st_mtime for h in headers)

    def compile_one(src: Path):
        obj = OBJ_DIR / (src.stem + ".o")
        if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr_t):
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", str(obj), str(src)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    if verbose:
        print("".join(log for _, log in results))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH),
           *[str(o) for o, _ in results]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return LIB_PATH


def declared_symbols() -> list[str]:
    """Every ``extern \"C\"`` function declared in include/g2vlm_b200.h."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(g2vlm_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load() -> ctypes.CDLL:
    """Load the kernel library; raise if it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA extension is mandatory; there is no CPU / PyTorch fallback)")
    lib = ctypes.CDLL(str(LIB_PATH))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    if missing:
        raise RuntimeError(f"{LIB_PATH} does not export {missing}; rebuild it")
    lib.g2vlm_last_error.restype = ctypes.c_char_p
    lib.g2vlm_abi_version.restype = ctypes.c_int
    have, want = lib.g2vlm_abi_version(), _header_abi_version()
    if have != want:   # a left-over library would silently misread the argument structs
        raise RuntimeError(f"{LIB_PATH} implements ABI v{have} but include/g2vlm_b200.h declares v{want}; rebuild it")
    if "G2VLM_B200_LIB" not in os.environ and _stale():
        import warnings
        warnings.warn(f"{LIB_PATH} is older than its sources; run __graft_entry__.build()")
    _lib = lib
    return lib
