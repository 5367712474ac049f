"""Package the UNMODIFIED reference sources of the hot path so they can travel to the GPU box (TEST / BENCH INFRASTRUCTURE).

    python oracle/make_ref.py          (authoring container only: needs /root/reference; also run by __graft_entry__.build())

The reference is pure Python, so there is nothing to compile: its own files

    src/caco_torch/**            (model: caco.py, audio_models/mae.py, text_models/roberta.py)
    src/eval/eval_caco_torch.py  (frontend + evaluation drivers)      src/eval/eval_utils.py
    src/eval/dataset_processors.py, src/eval/eval_dataset_configs.py  (imported by eval_caco_torch.py)

are stored byte-for-byte in ONE archive, ``oracle/_ref/caco_reference_src.zip`` (git-ignored: never in history, but shipped
to the GPU box with the snapshot like the built .so), next to a manifest of their SHA-256 digests.  ``oracle/ref_loader.py``
imports them straight from the archive (zipimport) — no reference source file is ever copied into the tree.  Only
``bench.py`` (``--impl reference`` and the ``gpu_library_baseline`` leg) and ``tests/`` may use it; the product never does.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
OUT_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(OUT_DIR, "caco_reference_src.zip")
MANIFEST = os.path.join(OUT_DIR, "MANIFEST.json")
FILES = ["src/eval/eval_caco_torch.py", "src/eval/eval_utils.py", "src/eval/dataset_processors.py",
         "src/eval/eval_dataset_configs.py"]


def _model_files():
    out = []
    for d, _, fs in os.walk(os.path.join(REF_ROOT, "src", "caco_torch")):
        for f in sorted(fs):
            if f.endswith(".py"):
                out.append(os.path.relpath(os.path.join(d, f), REF_ROOT))
    return sorted(out)


def build(verbose: bool = True) -> bool:
    """True if the archive was (re)built, False if /root/reference is absent (GPU box: the prebuilt archive is used)."""
    if not os.path.isdir(os.path.join(REF_ROOT, "src", "caco_torch")):
        return False
    os.makedirs(OUT_DIR, exist_ok=True)
    files = _model_files() + FILES
    manifest = {}
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        # `src` and `src/eval` have no __init__.py in the reference (namespace packages): zipimport only resolves those
        # through explicit directory entries
        for d in sorted({os.path.dirname(f) for f in files} | {"src"}):
            z.writestr(zipfile.ZipInfo(d + "/", date_time=(2020, 1, 1, 0, 0, 0)), b"")
        for rel in files:
            data = open(os.path.join(REF_ROOT, rel), "rb").read()
            manifest[rel] = hashlib.sha256(data).hexdigest()
            info = zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0))      # fixed timestamps: reproducible archive
            info.compress_type = zipfile.ZIP_DEFLATED
            z.writestr(info, data)
    json.dump({"source": REF_ROOT, "files": manifest}, open(MANIFEST, "w"), indent=1, sort_keys=True)
    if verbose:
        print(f"{ARCHIVE}: {len(files)} unmodified reference files")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
