from .DeepFMs import DeepFMs  # noqa: F401
from .QREmbeddingBag import QREmbeddingBag  # noqa: F401
