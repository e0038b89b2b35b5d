"""cirtorch_b200 -- B200-native (sm_100a) global-descriptor retrieval hot path of cirtorch.

Host side mirrors the reference's module / function API (paths relative to /root/reference):
  cirtorch/modules/pools.py, normalizations.py, heads/global_head.py  -> cirtorch_b200.modules.*
  (upstream cirtorch.layers.{pooling,normalization,functional} aliases  -> cirtorch_b200.layers.*)
  cirtorch/utils/whiten.py                                             -> cirtorch_b200.utils.whiten
  extract_vectors (scripts/test.py:200,236)                            -> cirtorch_b200.extract
  ranking (scripts/train_globalF.py:729-734)                           -> cirtorch_b200.search
  alpha-QE / DBA (BASELINE.json config 5)                              -> cirtorch_b200.rerank
  hard-negative mining (globalFeatures/tuples_dataset.py:213-350)      -> cirtorch_b200.mining
Every op calls hand-written CUDA kernels through the C-ABI in include/cir_b200.h (ctypes,
cirtorch_b200._lib); there is no CPU fallback.
"""
__version__ = "0.1.0"
