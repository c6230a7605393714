"""B200-native SAM ViT image-encoder forward (drop-in for ``sam.modeling.ImageEncoderViT``).

The directory name contains hyphens, so it is not importable by name; ``import iuvl_b200`` (the
loader module at the repo root) registers this package under the alias ``iuvl_b200``.
"""
from .config import EncoderConfig, PRESETS, state_dict_spec  # noqa: F401
from .synthetic import make_state_dict, make_images, rel_l2  # noqa: F401


def __getattr__(name):
    # heavy / native parts are imported lazily so that `import iuvl_b200` works without a GPU
    if name in ("ImageEncoderViT", "build_encoder", "sam_encoder_registry", "install_into_reference"):
        import importlib
        _enc = importlib.import_module(__name__ + ".encoder")
        return getattr(_enc, name)
    raise AttributeError(name)
