"""Recipe: make the reference's own package importable on the GPU box (TEST / BASELINE INFRASTRUCTURE).

``/root/reference`` exists only in the build container.  The reference (davidnabergoj/nfmc) is pure Python, so "building"
it means putting its package where ``bench.py --impl reference`` and the CPU-baseline leg can import it later:

    python oracle/build_ref.py        # /root/reference/nfmc/**/*.py  ->  oracle/_ref/nfmc/   (git-ignored, NOT gpurun-ignored)

Nothing under ``oracle/_ref`` is committed and no reference source enters the repository's history; the directory travels
to the GPU box with the snapshot exactly like the built ``.so`` files.  The third-party packages the reference imports and
that are absent everywhere (``torchflows``, ``potentials``) are satisfied by ``oracle/shim`` -- the oracle's RealNVP
restatement, so the flow arithmetic of this baseline stays "parity unpinned" (oracle/__init__.py).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/nfmc"
DST = os.path.join(HERE, "_ref", "nfmc")


def build_ref() -> bool:
    """Returns True if oracle/_ref/nfmc is present afterwards."""
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)              # GPU box: use what travelled with the snapshot
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    digest = hashlib.sha256()
    n_files = 0
    for root, _dirs, files in os.walk(SRC):
        rel = os.path.relpath(root, SRC)
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            src = os.path.join(root, f)
            shutil.copyfile(src, os.path.join(DST, rel, f))
            digest.update(open(src, "rb").read())
            n_files += 1
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": n_files, "sha256_of_sources": digest.hexdigest()}, fh)
    return True


def import_reference():
    """Put oracle/_ref and oracle/shim on sys.path; returns True if ``import nfmc`` (the reference) works."""
    ref = os.path.join(HERE, "_ref")
    if not os.path.isdir(os.path.join(ref, "nfmc")):
        return False
    for p in (os.path.join(HERE, "shim"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import nfmc  # noqa: F401
        return True
    except Exception:
        return False


if __name__ == "__main__":
    ok = build_ref()
    print("oracle/_ref/nfmc present:", ok)
