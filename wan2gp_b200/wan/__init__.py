from .model import WanModel  # noqa: F401
from .rope import get_rotary_pos_embed  # noqa: F401
from .vae import WanVAE  # noqa: F401
from .any2video import WanAny2V  # noqa: F401,E402
from .t5 import T5Encoder, T5EncoderModel  # noqa: F401,E402
