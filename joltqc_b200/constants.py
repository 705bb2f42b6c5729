"""Constants of the J/K path; same names and values as jqc/constants.py:21-38."""
LMAX = 4            # highest angular momentum (g)
NPRIM_MAX = 3       # primitives per kernel-side shell
BASIS_STRIDE = 12   # doubles per packed shell record
TILE = 4            # shells per screening tile edge / group padding
MAX_SMEM = 227 * 1024   # B200: opt-in dynamic shared memory per CTA (reference caps at 48 KiB)
NSM_B200 = 148
