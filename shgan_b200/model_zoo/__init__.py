"""Drop-in counterpart of the reference's `lib.model_zoo` for the generator-forward path: same registry
(`get_model`), same `type` names, constructor kwargs and state_dict keys."""
from .common.get_model import get_model, register  # noqa: F401
from . import stylegan, comodgan, shgan  # noqa: F401  (registers the model types)
