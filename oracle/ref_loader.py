"""Import the unmodified reference (gzhu06/Cacophony ``src.caco_torch`` + ``src.eval.eval_caco_torch``) from the archive
``oracle/make_ref.py`` wrote (TEST / BENCH INFRASTRUCTURE — never imported by the product package).

``load()`` returns ``(create_caco_model, eval_module)`` or raises ``ReferenceUnavailable``.  The two modules the reference's
``eval_utils.py`` imports but the hot path never calls (``astropy.stats.jackknife``, ``soundfile``) are not installed here and
are stubbed with empty modules, exactly as ``oracle/make_golden.py`` does; nothing of the reference itself is altered (the
manifest's SHA-256 digests are re-checked against the archive's members on load).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = os.path.join(HERE, "_ref", "caco_reference_src.zip")
MANIFEST = os.path.join(HERE, "_ref", "MANIFEST.json")


class ReferenceUnavailable(RuntimeError):
    pass


_loaded = None


def load():
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.exists(ARCHIVE):
        raise ReferenceUnavailable(f"{ARCHIVE} missing: run `python oracle/make_ref.py` where /root/reference is mounted")
    manifest = json.load(open(MANIFEST))["files"]
    with zipfile.ZipFile(ARCHIVE) as z:
        for rel, digest in manifest.items():
            if hashlib.sha256(z.read(rel)).hexdigest() != digest:
                raise ReferenceUnavailable(f"{rel} in {ARCHIVE} does not match its manifest digest")
    for m in ("astropy", "astropy.stats", "soundfile"):          # off-path imports of src/eval/eval_utils.py:2-3
        sys.modules.setdefault(m, types.ModuleType(m))
    if not hasattr(sys.modules["astropy.stats"], "jackknife"):
        sys.modules["astropy.stats"].jackknife = None
    if ARCHIVE not in sys.path:
        sys.path.insert(0, ARCHIVE)
    for name in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        mod = sys.modules[name]
        if ARCHIVE not in str(getattr(mod, "__file__", "") or getattr(mod, "__path__", "")):
            del sys.modules[name]                                  # a different `src` package is not the reference
    from src.caco_torch import create_caco_model
    from src.eval import eval_caco_torch as E
    if ARCHIVE not in (E.__file__ or ""):
        raise ReferenceUnavailable("`src.eval.eval_caco_torch` resolved outside the reference archive")
    _loaded = (create_caco_model, E)
    return _loaded
