"""
cuburn_b200 -- a Blackwell (sm_100a) implementation of cuburn's render hot path.

Host-side modules keep the reference's names (``profile``, ``render``,
``filters``, ``output``, ``genome.*``) so that callers written against
stevenrobertson/cuburn keep working; all device work goes through the C-ABI
library ``csrc/libcuburn_b200.so`` (see ``include/cuburn_b200.h``).
"""
__version__ = '0.1.0'
