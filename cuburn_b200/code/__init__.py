"""Host-side generators for the device code: genome packing and per-genome kernels."""
