"""`samd_sam_only.cache` is the same cache as `samd.cache` (the reference's two files differ only in
comments)."""
from samd.cache import SamdStaticCache, SamdCache  # noqa: F401
