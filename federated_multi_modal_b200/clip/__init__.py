from . import model  # noqa: F401
from .model import build_model  # noqa: F401
from ..synth import synthetic_tokenize as tokenize  # noqa: F401  (stand-in for clip.tokenize; no BPE vocab offline)
