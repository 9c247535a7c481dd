"""Host-side callers of the hot path: the reference's data loaders, re-stated to feed the device at line rate."""
