"""Puts the UNMODIFIED reference package where bench.py's reference arm can import it on the GPU box.

    python baseline/install_reference.py            # /root/reference/audiblelight -> baseline/_ref/audiblelight

The prescribed `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
/root/reference` fails in this image (the reference's build backend, poetry-core, is not installed and there is no
index), also with --no-deps; the reference is pure Python, so its package directory is copied verbatim instead
(SURVEY.md App. E). baseline/_ref/ is git-ignored (reference sources never enter the history) and is NOT
gpurun-ignored, so the copy travels to the GPU box. MANIFEST.sha256 lists every copied file with the hash of its
source, and `verify()` re-checks the copy against it before the arm runs.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_DEFAULT = "/root/reference"


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def install(src_root: str = SRC_DEFAULT) -> bool:
    src = os.path.join(src_root, "audiblelight")
    if not os.path.isdir(src):
        return False
    dst = os.path.join(DEST, "audiblelight")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(DEST, exist_ok=True)
    lines = []
    for dirpath, dirnames, filenames in os.walk(src):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, src_root)
        os.makedirs(os.path.join(DEST, rel), exist_ok=True)
        for fn in sorted(filenames):
            if fn.endswith((".pyc", ".pyo")):
                continue
            s, d = os.path.join(dirpath, fn), os.path.join(DEST, rel, fn)
            shutil.copyfile(s, d)
            lines.append(f"{_sha(s)}  {os.path.join(rel, fn)}")
    # worldstate.py:37 reads <project root>/resources/mp3d_material_config.json at import time
    for rel in ("resources/mp3d_material_config.json",):
        sfile = os.path.join(src_root, rel)
        if os.path.exists(sfile):
            os.makedirs(os.path.dirname(os.path.join(DEST, rel)), exist_ok=True)
            shutil.copyfile(sfile, os.path.join(DEST, rel))
            lines.append(f"{_sha(sfile)}  {rel}")
    with open(os.path.join(DEST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(sorted(lines)) + "\n")
    return True


def verify() -> bool:
    """True when baseline/_ref holds the files of the manifest, byte for byte."""
    man = os.path.join(DEST, "MANIFEST.sha256")
    if not os.path.exists(man):
        return False
    for line in open(man):
        line = line.strip()
        if not line:
            continue
        digest, rel = line.split("  ", 1)
        path = os.path.join(DEST, rel)
        if not os.path.exists(path) or _sha(path) != digest:
            return False
    return True


if __name__ == "__main__":
    ok = install(sys.argv[1] if len(sys.argv) > 1 else SRC_DEFAULT)
    print("installed" if ok else "reference source not found", "-> verify:", verify())
