from .loader import load_model, pack_model  # noqa: F401
