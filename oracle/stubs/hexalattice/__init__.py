"""Import-time stand-in for hexalattice (absent; only the Honeycomb element uses it)."""
