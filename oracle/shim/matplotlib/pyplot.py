"""Empty stand-in (oracle test infrastructure)."""
