"""Moved to oracle/reference_loader.py (shared by the tests and by bench.py's baseline legs)."""
from oracle.reference_loader import *  # noqa: F401,F403
from oracle.reference_loader import (HOT_PREFIXES, RecordingEncoder, ReplayDecoder, _Inject, build_reference_net,  # noqa: F401
                                     inject_latents, load_reference_dcae_module, reference_available, reference_file)
