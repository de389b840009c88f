"""Import shims that let the UNMODIFIED reference modules under /root/reference
run on CPU in this container (test infrastructure only -- see oracle/README.md).

The reference (deepbeepmeep/Wan2GP) needs third-party packages that are not
installed here (mmgp, diffusers) and imports its whole model zoo from
models/wan/__init__.py.  SURVEY.md section 8c lists the blockers; this module
clears them with stub modules so that

    models/wan/modules/model.py      (WanModel, reference hot path W1-W11)
    models/wan/modules/vae.py        (WanVAE_,  reference hot path V1-V7)
    models/wan/modules/posemb_layers (RoPE tables, W0)

import and execute exactly as written.  Nothing here is arithmetic: every
number that comes out of `load_reference()` is produced by reference code.

/root/reference does not exist on the GPU box, so only oracle/gen_golden.py
(run here, output committed under tests/golden/) and the CPU-only validation
tests that are skipped when the tree is absent may call this.
"""
import functools
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("WAN2GP_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models/wan/modules/model.py"))


class _Stub(types.ModuleType):
    """Module whose every attribute is an empty class (sibling conditioning
    modules imported at model.py:18-27 and never touched on the t2v/i2v2_2 path)."""

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return type(k, (), {})


_loaded = None


def load_reference():
    """Returns a namespace with the reference's WanModel, WanVAE_, get_rotary_pos_embed."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch

    sys.path.insert(0, REFERENCE_ROOT)
    # ---- mmgp (requirements.txt:2, not installed): no arithmetic on this path ----
    mmgp = types.ModuleType("mmgp")
    offload = types.ModuleType("mmgp.offload")
    offload.shared_state = {"_attention": "sdpa"}
    _caches = {}
    offload.get_cache = lambda name: _caches.setdefault(name, {})
    offload.clear_caches = _caches.clear
    mmgp.offload = offload
    st2 = types.ModuleType("mmgp.safetensors2")
    mmgp.safetensors2 = st2
    sys.modules.update({"mmgp": mmgp, "mmgp.offload": offload, "mmgp.safetensors2": st2})

    # ---- diffusers (only base classes / decorator are used at model.py:10-11) ----
    class ConfigMixin:
        pass

    class ModelMixin(torch.nn.Module):
        pass

    def register_to_config(init):
        @functools.wraps(init)
        def w(self, *a, **k):
            return init(self, *a, **k)
        return w

    d = types.ModuleType("diffusers")
    dc = types.ModuleType("diffusers.configuration_utils")
    dm = types.ModuleType("diffusers.models")
    dmu = types.ModuleType("diffusers.models.modeling_utils")
    dc.ConfigMixin = ConfigMixin
    dc.register_to_config = register_to_config
    dmu.ModelMixin = ModelMixin
    dm.ModelMixin = ModelMixin
    sys.modules.update({"diffusers": d, "diffusers.configuration_utils": dc,
                        "diffusers.models": dm, "diffusers.models.modeling_utils": dmu})

    # shared/attention.py:14 probes the CUDA device at import time
    torch.cuda.get_device_capability = lambda *a, **k: (0, 0)

    # bypass models/wan/__init__.py (imports the whole pipeline zoo)
    for name, rel in [("models", "models"), ("models.wan", "models/wan"),
                      ("models.wan.modules", "models/wan/modules")]:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
        sys.modules[name] = m
    for name in ["multitalk", "multitalk.multitalk_utils", "animate", "animate.motion_encoder",
                 "animate.face_blocks", "animate.model_animate", "scail", "scail.model_scail",
                 "scail2", "steadydancer", "steadydancer.small_archs",
                 "steadydancer.mobilenetv2_dcd", "shotplan", "animate2"]:
        sys.modules["models.wan." + name] = _Stub("models.wan." + name)

    from models.wan.modules.model import WanModel, sinusoidal_embedding_1d
    from models.wan.modules.posemb_layers import get_rotary_pos_embed
    from models.wan.modules.vae import WanVAE_

    ns = types.SimpleNamespace(WanModel=WanModel, WanVAE_=WanVAE_,
                               get_rotary_pos_embed=get_rotary_pos_embed,
                               sinusoidal_embedding_1d=sinusoidal_embedding_1d,
                               offload=offload)
    _loaded = ns
    return ns


class Pipe:
    """Stand-in for the pipeline object polled at model.py:1997."""
    _interrupt = False


# ----------------------------------------------------------------------------- Hunyuan Video
_loaded_hy = None


def load_reference_hy():
    """Reference HYVideoDiffusionTransformer (models/hyvideo/modules/models.py) importable on CPU.
    Extra shims: register_to_config must populate self.config (models.py:710), namespace
    packages for models.hyvideo{,.modules,.utils,.text_encoder}; ByT5Mapper is imported from the reference's own
    text_encoder/byT5/__init__.py when its imports resolve, else that module is stubbed (byT5 branch unused)."""
    global _loaded_hy
    if _loaded_hy is not None:
        return _loaded_hy
    import importlib
    import inspect

    import torch
    load_reference()          # mmgp / diffusers stubs, namespace package `models`
    dc = sys.modules["diffusers.configuration_utils"]

    def register_to_config(init):
        @functools.wraps(init)
        def w(self, *a, **k):
            ba = inspect.signature(init).bind(self, *a, **k)
            ba.apply_defaults()
            self.config = types.SimpleNamespace(**{n: v for n, v in ba.arguments.items() if n != "self"})
            return init(self, *a, **k)
        return w
    dc.register_to_config = register_to_config
    sys.modules["diffusers"].ModelMixin = sys.modules["diffusers.models"].ModelMixin
    sys.modules["diffusers"].ConfigMixin = dc.ConfigMixin
    for name, rel in [("models.hyvideo", "models/hyvideo"), ("models.hyvideo.modules", "models/hyvideo/modules"),
                      ("models.hyvideo.utils", "models/hyvideo/utils"), ("models.hyvideo.text_encoder", "models/hyvideo/text_encoder")]:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
        sys.modules[name] = m
    byt5_real = True
    try:
        importlib.import_module("models.hyvideo.text_encoder.byT5")
    except Exception:
        byt5_real = False

        class _NNStub(types.ModuleType):
            def __getattr__(self, k):
                if k.startswith("__"):
                    raise AttributeError(k)
                return type(k, (torch.nn.Module,), {})
        sys.modules["models.hyvideo.text_encoder.byT5"] = _NNStub("models.hyvideo.text_encoder.byT5")
    from models.hyvideo.modules.models import HUNYUAN_VIDEO_CONFIG, HYVideoDiffusionTransformer
    from models.hyvideo.modules.posemb_layers import get_nd_rotary_pos_embed
    _loaded_hy = types.SimpleNamespace(HYVideoDiffusionTransformer=HYVideoDiffusionTransformer, CONFIGS=HUNYUAN_VIDEO_CONFIG,
                                       get_nd_rotary_pos_embed=get_nd_rotary_pos_embed, byt5_real=byt5_real)
    return _loaded_hy


def hook_linear_input_cast(model):
    """The HY blocks hard-cast activations to torch.bfloat16 (models.py:211, 290, 302, 313) so the reference only runs with
    bf16 weights.  To pin the oracle tightly we run it with fp32 weights and let every nn.Linear/Conv cast its INPUT to the
    weight dtype (a forward pre-hook; reference code untouched): the run is then fp32 arithmetic with bf16 roundings exactly
    at the reference's hard-cast points."""
    import torch

    def pre(m, args):
        return (args[0].to(m.weight.dtype),) + tuple(args[1:])
    for mod in model.modules():
        if isinstance(mod, (torch.nn.Linear, torch.nn.Conv3d)):
            mod.register_forward_pre_hook(pre)
    return model


def split_linear_modules(model, split_map):
    """Restatement of mmgp.offload.split_linear_modules (mmgp==3.7.12, requirements.txt:2 -- third-party, NOT in the reference
    tree) as used at models/hyvideo/hunyuan_handler.py:274-278: every sub-module that owns a Linear named <key> gets extra
    Linear children `mapped_modules[i]` holding consecutive row slices (`split_sizes`) of its weight / bias.  No arithmetic."""
    import torch
    for mod in list(model.modules()):
        for key, spec in split_map.items():
            lin = getattr(mod, key, None)
            if not isinstance(lin, torch.nn.Linear):
                continue
            sizes = spec["split_sizes"]
            total = lin.weight.shape[0]
            if sum(sizes) != total:                      # the map is written for hidden_size 3072: rescale for reduced configs
                sizes = [s * total // sum(sizes) for s in sizes]
            off = 0
            for name, n in zip(spec["mapped_modules"], sizes):
                sub = torch.nn.Linear(lin.in_features, n, bias=lin.bias is not None)
                sub.weight = torch.nn.Parameter(lin.weight[off:off + n].detach().clone(), requires_grad=False)
                if lin.bias is not None:
                    sub.bias = torch.nn.Parameter(lin.bias[off:off + n].detach().clone(), requires_grad=False)
                setattr(mod, name, sub)
                off += n
    return model


_loaded_hyvae = None


def load_reference_hyvae():
    """Reference Hunyuan Video 1.5 VAE decoder (models/hyvideo/vae/hunyuanvideo_15_vae.py::Decoder) importable on CPU: extra
    diffusers stubs (BaseOutput, DiagonalGaussianDistribution, AutoencoderKLOutput) -- none of them carries arithmetic."""
    global _loaded_hyvae
    if _loaded_hyvae is not None:
        return _loaded_hyvae
    load_reference_hy()
    for name in ["diffusers.models.autoencoders", "diffusers.models.autoencoders.vae", "diffusers.models.modeling_outputs", "diffusers.utils"]:
        sys.modules.setdefault(name, types.ModuleType(name))

    class BaseOutput(dict):
        pass
    sys.modules["diffusers.models.autoencoders.vae"].BaseOutput = BaseOutput
    class DiagonalGaussianDistribution:          # diffusers' posterior container: only `.parameters` (the encoder's moments) is read here
        def __init__(self, parameters, deterministic=False):
            self.parameters = parameters
    sys.modules["diffusers.models.autoencoders.vae"].DiagonalGaussianDistribution = DiagonalGaussianDistribution
    sys.modules["diffusers.models.modeling_outputs"].AutoencoderKLOutput = type("AutoencoderKLOutput", (BaseOutput,), {})
    m = types.ModuleType("models.hyvideo.vae")
    m.__path__ = [os.path.join(REFERENCE_ROOT, "models/hyvideo/vae")]
    sys.modules["models.hyvideo.vae"] = m
    from models.hyvideo.vae.hunyuanvideo_15_vae import Decoder, Encoder
    _loaded_hyvae = types.SimpleNamespace(Decoder=Decoder, Encoder=Encoder)
    return _loaded_hyvae


_loaded_hyvae10 = None


def load_reference_hyvae10():
    """Reference HunyuanVideo 1.0 VAE decoder (models/hyvideo/vae/vae.py::DecoderCausal3D + unet_causal_3d_blocks.py).
    Its mid block uses `diffusers.models.attention_processor.Attention` -- THIRD-PARTY code (diffusers==0.36.0,
    requirements.txt:4) that is not in the reference tree.  `_Attention` below restates the published behaviour of
    Attention(+AttnProcessor2_0) for the exact constructor call at unet_causal_3d_blocks.py:690-703 (single head,
    GroupNorm over tokens, to_q/to_k/to_v/to_out Linears with bias, residual connection, rescale_output_factor): parity of
    that one block is pinned to this restatement, not to diffusers itself ("parity unpinned" at that boundary)."""
    global _loaded_hyvae10
    if _loaded_hyvae10 is not None:
        return _loaded_hyvae10
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    load_reference_hyvae()

    class _Attention(nn.Module):
        def __init__(self, query_dim, heads=8, dim_head=64, rescale_output_factor=1.0, eps=1e-5, norm_num_groups=None,
                     spatial_norm_dim=None, residual_connection=False, bias=False, upcast_softmax=False,
                     _from_deprecated_attn_block=False, **kw):
            super().__init__()
            inner = heads * dim_head
            self.heads, self.rescale, self.residual = heads, rescale_output_factor, residual_connection
            self.group_norm = nn.GroupNorm(norm_num_groups, query_dim, eps=eps, affine=True) if norm_num_groups else None
            self.to_q, self.to_k, self.to_v = (nn.Linear(query_dim, inner, bias=bias) for _ in range(3))
            self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])

        def forward(self, hidden_states, temb=None, attention_mask=None, **kw):
            residual = hidden_states
            b, n, c = hidden_states.shape
            if self.group_norm is not None:
                hidden_states = self.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
            q, k, v = (f(hidden_states).view(b, n, self.heads, -1).transpose(1, 2) for f in (self.to_q, self.to_k, self.to_v))
            if attention_mask is not None:
                attention_mask = attention_mask.view(b, 1, n, n)
            o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask)
            o = self.to_out[0](o.transpose(1, 2).reshape(b, n, -1))
            if self.residual:
                o = o + residual
            return o / self.rescale

    def mod(name, **attrs):
        m = sys.modules.setdefault(name, types.ModuleType(name))
        for k, v in attrs.items():
            setattr(m, k, v)
        return m
    blank = lambda n: type(n, (nn.Module,), {})                                      # noqa: E731
    logging = types.SimpleNamespace(get_logger=lambda *a, **k: types.SimpleNamespace(warn=print, warning=print, info=print))
    mod("diffusers.utils", is_torch_version=lambda *a: True, logging=logging, BaseOutput=dict)
    mod("diffusers.utils.torch_utils", randn_tensor=lambda *a, **k: torch.randn(*a))
    mod("diffusers.utils.accelerate_utils", apply_forward_hook=lambda f: f)
    mod("diffusers.models.activations", get_activation=lambda n: {"swish": nn.SiLU(), "silu": nn.SiLU()}[n])
    mod("diffusers.models.attention_processor", Attention=_Attention, SpatialNorm=blank("SpatialNorm"), AttentionProcessor=object,
        AttnAddedKVProcessor=object, AttnProcessor=object, ADDED_KV_ATTENTION_PROCESSORS=(), CROSS_ATTENTION_PROCESSORS=())
    mod("diffusers.models.normalization", AdaGroupNorm=blank("AdaGroupNorm"), RMSNorm=blank("RMSNorm"))
    mod("diffusers.loaders", FromOriginalVAEMixin=object, FromOriginalModelMixin=object)
    mod("diffusers.loaders.single_file_model", FromOriginalModelMixin=object)
    load_reference_hy()                              # ConfigMixin / ModelMixin / register_to_config shims
    sys.modules.setdefault("loguru", types.SimpleNamespace(logger=types.SimpleNamespace(warning=print, info=print)))
    from models.hyvideo.vae.vae import DecoderCausal3D
    from models.hyvideo.vae.autoencoder_kl_causal_3d import AutoencoderKLCausal3D
    _loaded_hyvae10 = types.SimpleNamespace(DecoderCausal3D=DecoderCausal3D, AutoencoderKLCausal3D=AutoencoderKLCausal3D)
    return _loaded_hyvae10


_loaded_unipc = None


def load_reference_unipc():
    """Reference FlowUniPCMultistepScheduler (shared/utils/fm_solvers_unipc.py), the default Wan sample solver (any2video.py:518-522).
    diffusers' SchedulerMixin / ConfigMixin only provide config plumbing; register_to_config must populate self.config."""
    global _loaded_unipc
    if _loaded_unipc is not None:
        return _loaded_unipc
    import enum
    import functools
    import importlib.util
    import inspect

    def mod(name, **attrs):
        m = sys.modules.setdefault(name, types.ModuleType(name))
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    def register_to_config(init):
        @functools.wraps(init)
        def wrapped(self, *a, **kw):
            sig = inspect.signature(init)
            bound = sig.bind(self, *a, **kw)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
            self.config = types.SimpleNamespace(**cfg)
            self.register_to_config = lambda **u: self.config.__dict__.update(u)
            init(self, *a, **kw)
        return wrapped

    class Karras(enum.Enum):
        UniPCMultistepScheduler = 1
    mod("diffusers")
    mod("diffusers.configuration_utils", ConfigMixin=type("ConfigMixin", (), {}), register_to_config=register_to_config)
    mod("diffusers.schedulers")
    mod("diffusers.schedulers.scheduling_utils", KarrasDiffusionSchedulers=Karras, SchedulerMixin=type("SchedulerMixin", (), {}),
        SchedulerOutput=lambda prev_sample: types.SimpleNamespace(prev_sample=prev_sample))
    du = mod("diffusers.utils")
    du.deprecate = lambda *a, **k: None
    du.is_scipy_available = lambda: False
    spec = importlib.util.spec_from_file_location("_ref_fm_solvers_unipc", os.path.join(REFERENCE_ROOT, "shared/utils/fm_solvers_unipc.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    du.BaseOutput = dict
    mod("diffusers.utils.torch_utils", randn_tensor=lambda *a, **k: None)
    spec2 = importlib.util.spec_from_file_location("_ref_fm_solvers", os.path.join(REFERENCE_ROOT, "shared/utils/fm_solvers.py"))
    m2 = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(m2)
    du.logging = types.SimpleNamespace(get_logger=lambda *a, **k: types.SimpleNamespace(warning=print, info=print, warn=print))
    mods = {}
    for nm in ("lcm_scheduler", "basic_flowmatch", "../../models/hyvideo/diffusion/schedulers/scheduling_flow_match_discrete"):
        sp = importlib.util.spec_from_file_location("_ref_" + os.path.basename(nm), os.path.normpath(os.path.join(REFERENCE_ROOT, "shared/utils", nm + ".py")))
        mods[nm] = importlib.util.module_from_spec(sp)
        sp.loader.exec_module(mods[nm])
    _loaded_unipc = types.SimpleNamespace(FlowUniPCMultistepScheduler=m.FlowUniPCMultistepScheduler,
                                          LCMScheduler=mods["lcm_scheduler"].LCMScheduler, FlowMatchScheduler=mods["basic_flowmatch"].FlowMatchScheduler,
                                          FlowMatchDiscreteScheduler=mods["../../models/hyvideo/diffusion/schedulers/scheduling_flow_match_discrete"].FlowMatchDiscreteScheduler,
                                          FlowDPMSolverMultistepScheduler=m2.FlowDPMSolverMultistepScheduler,
                                          get_sampling_sigmas=m2.get_sampling_sigmas, retrieve_timesteps=m2.retrieve_timesteps)
    return _loaded_unipc


_loaded_t5 = None


def load_reference_t5():
    """The reference's T5Encoder (models/wan/modules/t5.py), loaded from its file without importing the `models.wan` package (which pulls
    the whole model zoo).  Stubs: `shared.utils.gguf_mapping` (checkpoint-name remapping only) and the sibling `tokenizers` module (ftfy is
    not installed; the encoder takes token ids).  No arithmetic is replaced."""
    global _loaded_t5
    if _loaded_t5 is not None:
        return _loaded_t5
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, "models/wan/modules/t5.py")):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import importlib.util

    for name in ("shared", "shared.utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    gm = types.ModuleType("shared.utils.gguf_mapping")
    gm.has_standard_gguf_tensor_names = lambda sd: False
    gm.remap_state_dict_triplet = lambda *a, **k: a[:3]
    sys.modules["shared.utils.gguf_mapping"] = gm
    pkg = types.ModuleType("_ref_t5pkg")
    pkg.__path__ = []
    tok = types.ModuleType("_ref_t5pkg.tokenizers")
    tok.HuggingfaceTokenizer = type("HuggingfaceTokenizer", (), {})
    sys.modules.update({"_ref_t5pkg": pkg, "_ref_t5pkg.tokenizers": tok})
    spec = importlib.util.spec_from_file_location("_ref_t5pkg.t5", os.path.join(REFERENCE_ROOT, "models/wan/modules/t5.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_ref_t5pkg.t5"] = mod
    import torch
    cur = torch.cuda.current_device           # t5.py:640 evaluates torch.cuda.current_device() as a default argument at import time
    torch.cuda.current_device = lambda: 0
    try:
        spec.loader.exec_module(mod)
    finally:
        torch.cuda.current_device = cur
    _loaded_t5 = mod
    return mod
