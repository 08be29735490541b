"""Host-side utilities either side of the hot path."""
