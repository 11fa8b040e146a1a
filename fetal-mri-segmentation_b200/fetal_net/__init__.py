"""fetal_net — B200-native drop-in for the hot path of GalDude33/Fetal-MRI-Segmentation.

Mirrors the reference package layout for the path it replaces:
  fetal_net.model       builders looked up by `getattr(fetal_net.model, config['model_name'])`
                        (fetal/train_fetal.py:31-32)
  fetal_net.metrics     losses looked up by `getattr(fetal_net.metrics, config['loss'])`
  fetal_net.prediction  patch_wise_prediction and friends (fetal_net/prediction.py)
  fetal_net.training    train_model / load_old_model (fetal_net/training.py)
Numerics run in libfetalb200.so (hand-written sm_100a CUDA behind the C ABI of include/fetal_b200.h).
"""
__all__ = ["model", "metrics", "prediction", "training"]
