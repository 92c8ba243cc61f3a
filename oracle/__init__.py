"""
ORACLE -- test infrastructure, not product code.

A CPU restatement of the reference renderer's hot path (stevenrobertson/cuburn),
used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under cuburn_b200/ imports this package.

Parity status.  The reference as a whole cannot be executed here (Python 2 +
PyCUDA + CUDA-4-era texture references), but large parts of
it can, and the oracle is pinned against every one of them:

  * the reference's own DEVICE code compiled for the CPU from the sources in place
    (oracle/build_ref.py -> oracle/_ref/): all 95 variation bodies, catmull_rom /
    catmull_rom_mag with the knot search, the camera / affine / variation precalc
    hunks, precalc_densities (3 and 6 xforms), the YUV helpers, interp_palette_flat, every filter kernel (run through a
    serial CUDA shim with clamp-to-edge textures) and all six f32_to_* pixel-format
    kernels -- tests/test_reference_code.py
  * the reference's own HOST code executed under Python 3 (tests/golden/make_*_golden.py):
    SplineEval, profile enumeration, make_seeds, flam3 conversion and blending
  * the MWC recurrence model (code/mwc.py:90-129) and the multiplier table
    (code/primes.bin, regenerated and compared byte for byte)
  * the pixel-format assertions of code/tests/test_output.py and the profile tests

Not covered by reference-generated vectors: the `iter` kernel's own control flow
(weighted xform choice, inter-warp shuffle, binning through `cvt.rni`, packed-u64
accumulation and `flush_atom` -- nested tempita control flow plus inline PTX).  There
the oracle is a line-by-line restatement of the cited sources: **parity unpinned** for
that part only.
"""
