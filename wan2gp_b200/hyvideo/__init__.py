from .model import HYVideoDiffusionTransformer, get_rotary_pos_embed  # noqa: F401
from .vae import AutoencoderKLConv3D, HYVAEDecoder, HYVAEEncoder  # noqa: F401
from .vae10 import AutoencoderKLCausal3D, HYVAE10Decoder, HYVAE10Encoder  # noqa: F401
from .byt5 import ByT5Encoder  # noqa: F401
from .llm import LlamaLikeTextModel  # noqa: F401
from .hunyuan import HunyuanVideoSampler  # noqa: F401
