"""samd_sam_only drop-in package (filled in below)."""
