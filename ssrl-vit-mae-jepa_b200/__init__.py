"""B200-native MAE pre-training hot path (drop-in for giolucasd/ssrl-vit-mae-jepa's ``src/models/mae.py`` +
``src/training/mae.py``).  Import as ``ssrl_vit_mae_jepa_b200`` (alias package; the directory name carries the
reference's hyphens)."""
from .build import build  # noqa: F401
from .mae import MAEEngine, MaskedAutoencoder, num_keep_tokens, sincos_2d  # noqa: F401
from . import training  # noqa: F401
from .training import DevicePrefetcher, FusedTrainer, MAEPretrainModule  # noqa: F401
from .classifier import ClassificationHead, ViTClassifier, ViTClassifierTrainModule, ViTEngine, VisionTransformer  # noqa: F401
from . import data, ops, trainer  # noqa: F401,E402  (ops: registers torch.ops.mae_b200.*)
from .trainer import LiteTrainer, ModelCheckpoint  # noqa: F401,E402
