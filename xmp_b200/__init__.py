"""xmp_b200: B200-native batched Fitch scoring / exact maximum-parsimony branch and bound.

The product is the C-ABI library ``libxmp_b200.so`` (hand-written sm_100a CUDA, ``include/xmp_cuda.h``)
driven by the C host in ``host/`` (``fastdnamp_b200``).  This package is the thin Python mirror of
those two C interfaces used by ``bench.py`` and the tests: ``xmp_b200.hostlib`` wraps the host
preprocessing / tree output library, ``xmp_b200.cuda`` wraps the device library.  There is no CPU
fallback: ``xmp_b200.cuda`` raises if the CUDA library or a GPU is missing.
"""
__version__ = "0.1.0"
