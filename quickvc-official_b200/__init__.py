"""quickvc-official_b200: QuickVC's `SynthesizerTrn.infer` on B200 (sm_100a).

The directory name carries the reference's hyphen; import it as `quickvc_official_b200`
(the sibling shim package) -- `from quickvc_official_b200 import SynthesizerTrn`.
"""
from .models import SynthesizerTrn  # noqa: F401
from .capi import QvcError, launch_count  # noqa: F401

__all__ = ["SynthesizerTrn", "QvcError", "launch_count"]
