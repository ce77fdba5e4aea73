"""Drop-in for the reference's `samd_sam_only` package (samd_sam_only/__init__.py:1-5)."""
from .samd_config import SamdConfig
from .samd_model import SamdModel
from .utils import SamdGenerationConfig
from .sam import build_sam, load_sam, dump_sam
from .draft import DraftModel
