"""samd drop-in package (filled in below)."""
