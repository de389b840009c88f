from .model import HYVideoDiffusionTransformer, get_rotary_pos_embed  # noqa: F401
