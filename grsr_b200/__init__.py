"""grsr_b200 -- B200 (sm_100a) implementation of GRIMM's unichromosomal reversal-distance and
sorting-by-reversals path behind a C ABI (include/grsr.h).

The package holds the CUDA sources (csrc/), the build recipe (build.py), a ctypes binding of the
C ABI for tests and bench (capi.py), the C host program that mirrors GRIMM's command line
(csrc/grimm_main.c -> bin/grimm) and seeded input generators (gen.py).  There is no CPU
implementation of the path in here: every compute call goes to libgrsr_cuda.so and fails loudly
when that library or a CUDA device is missing.
"""
