"""spring_b200 -- B200-native reorder + encode hot path of the SPRING FASTQ compressor.

Only what the path needs lives here:
  csrc/      CUDA kernels (sm_100a) + the C ABI (include/spring_b200.h)
  capi.py    ctypes binding of the C ABI (the CUDA library is mandatory: no CPU fallback)
  hotpath.py host-side mirror of the reference's call_reorder / call_encoder
  dnaio.py   on-disk record formats of the path's inputs and outputs
  synth.py   seeded synthetic read sets (BASELINE.json configs)
"""
__version__ = "0.1.0"
