"""Stage the UNMODIFIED reference for the CPU arm of bench.py -- test / measurement infrastructure, not product code.

The reference (tarepan/QuickVC-official) is a flat directory of Python scripts with a poetry `pyproject.toml`; the
prescribed offline install
    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
fails in this image (build backend `poetry-core` is not installed and cannot be fetched), and the tree has no package
to install anyway: `convert.py` imports `models`, `utils` ... from its own directory.  So "installing" it means
putting the files `infer` needs on a path: this script copies them, byte for byte, from /root/reference into
`baseline/_ref/` -- a git-IGNORED directory (never committed, not product source) that `gpurun` ships to the GPU box
like the built `.so`.  `bench.py --impl reference` and `bench.py`'s `cpu_baseline` leg then time the reference's own
`SynthesizerTrn.infer` (kind "reference"); when `baseline/_ref` is absent they fall back to the oracle port (kind
"port").  `__graft_entry__.build()` runs this whenever /root/reference is present.

Files: the import closure of `models.SynthesizerTrn.infer` -- models.py, modules.py, commons.py, pqmf.py (imported at
module level, never instantiated for ms_istft_vits) -- plus configs/quickvc.json.
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("models.py", "modules.py", "commons.py", "pqmf.py", os.path.join("configs", "quickvc.json"))


def stage(reference: str = "/root/reference") -> bool:
    """Copy the files; returns False (and stages nothing) when the reference tree is not mounted."""
    if not os.path.isdir(reference):
        return False
    for rel in FILES:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(reference, rel), dst)
    return True


def staged() -> bool:
    return all(os.path.isfile(os.path.join(DEST, rel)) for rel in FILES)


def load_reference_models(path: str = DEST):
    """Import the reference's `models` module from `path` (SciPy shim of SURVEY.md section 8c: pqmf.py:13 imports
    `scipy.signal.kaiser`, which SciPy >= 1.13 moved to scipy.signal.windows)."""
    import scipy.signal
    import scipy.signal.windows
    if not hasattr(scipy.signal, "kaiser"):
        scipy.signal.kaiser = scipy.signal.windows.kaiser
    if path not in sys.path:
        sys.path.insert(0, path)
    import models  # noqa: PLC0415  (the reference's models.py)
    return models


if __name__ == "__main__":
    ok = stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("staged" if ok else "reference tree not present; nothing staged", DEST)
