"""CPU dry-run of the drop-in boundary levels 1-2: the plugin's `plugin_info.json` resolves
to a `family_handler` with the static methods WanGP's registry calls, `load_model` returns `(pipeline_obj, pipe_dict)`, and
`pipeline_obj.generate(**kwargs)` accepts the keyword set wgp.py:7762-7880 passes and honours the return / callback / interrupt
contract of any2video.py:414-503, 1409, 1446, 1741-1746, 1810-1826.  The C ABI is stubbed (no arithmetic: outputs are uninitialised);
the arithmetic behind generate() is covered by the `-m gpu` tests (test_pipeline_gpu.py, test_plugin_gpu.py)."""
import importlib
import json
import os

import pytest
import torch

from tests.test_host_dryrun_cpu import stub_abi  # noqa: F401  (fixture)
from wan2gp_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_handler():
    info = json.load(open(os.path.join(ROOT, "plugin", "plugin_info.json")))
    assert info["type"] == "model" and os.path.isdir(os.path.join(ROOT, "plugin", info["defaults"]))
    mods = info["model_handlers"] if isinstance(info["model_handlers"], list) else [info["model_handlers"]]
    assert all(m.startswith(".") for m in mods)                    # relative to the plugin package root (docs/PLUGINS.md:45)
    return importlib.import_module("plugin" + mods[0])


# the keyword set wgp.py:7762-7880 passes to wan_model.generate for a plain generation (values of an idle t2v / i2v session)
def wgp_kwargs(**over):
    kw = dict(input_prompt="a cat", alt_prompt="", image_start=None, image_end=None, input_frames=None, input_frames2=None, input_ref_images=None,
              input_ref_masks=None, input_masks=None, input_masks2=None, input_video=None, input_faces=None, input_custom=None, video_guide=None,
              video_guide2=None, denoising_strength=0.9, masking_strength=0.1, prefix_frames_count=0, frame_num=5, batch_size=1, height=32,
              width=48, fit_into_canvas=True, shift=12.0, sample_solver="unipc", sampling_steps=4, guide_scale=4.0, guide2_scale=3.0,
              guide3_scale=5.0, switch_threshold=875, switch2_threshold=0, guide_phases=2, model_switch_phase=1, embedded_guidance_scale=6.0,
              n_prompt="", seed=7, callback=None, enable_RIFLEx=False, VAE_tile_size=0, joint_pass=True, perturbation_switch=0,
              perturbation_layers=None, perturbation_start=0.0, perturbation_end=1.0, apg_switch=0, cfg_star_switch=1, cfg_zero_step=-1,
              alt_guide_scale=1.0, audio_cfg_scale=4.0, input_waveform=None, input_waveform_sample_rate=None, audio_guide=None, audio_guide2=None,
              audio_prompt_type="", audio_proj=None, audio_scale=None, audio_context_lens=None, context_scale=None, control_scale_alt=1.0,
              alt_scale=1.0, motion_amplitude=1.0, model_mode=None, causal_block_size=5, causal_attention=True, fps=16, overlapped_latents=None,
              return_latent_slice=None, overlap_noise=0, overlap_size=0, sub_parallel_window_size=0, sub_parallel_window_overlap=0,
              color_correction_strength=1, conditioning_latents_size=0, input_video_is_hdr=False, lora_dir="loras", keep_frames_parsed=[],
              model_filename=["hi.safetensors", "lo.safetensors"], model_type="b200_t2v_2_2", loras_slists=None, NAG_scale=1, NAG_tau=3.5,
              NAG_alpha=0.5, attention_sparsity=0, speakers_bboxes=None, image_mode=0, video_prompt_type="", window_no=1, offloadobj=None,
              set_header_text=lambda *_: None, pre_video_frame=None, prefix_video=None, original_input_ref_images=[], image_refs_relative_size=50,
              outpainting_dims=None, face_arc_embeds=None, custom_settings=None, frame_window_options=None, gen_state={}, temperature=1.0,
              window_start_frame_no=0, input_video_strength=1.0, self_refiner_setting=0, self_refiner_plan="", self_refiner_f_uncertainty=0.0,
              self_refiner_certain_percentage=0.999, duration_seconds=0, pause_seconds=0, top_p=0.9, top_k=50, set_progress_status=None,
              loras_selected=[], frames_relative_positions_list=[], frames_to_inject=[], verbose_level=0, gen_cache={}, vae_upsampler=None,
              save_masks=False)
    kw.update(over)
    return kw


def fake_t5(prompts, device):
    return [torch.randn(5 + len(p) % 7, 128, generator=torch.Generator().manual_seed(len(p))) for p in prompts]


@pytest.fixture
def pipeline(stub_abi, monkeypatch):  # noqa: F811
    h = load_handler()
    monkeypatch.setitem(h.ARCHS, "b200_t2v_2_2", ("tiny", True))             # reduced architectures: the dry-run computes nothing
    monkeypatch.setitem(h.ARCHS, "b200_i2v_2_2", ("tiny_i2v", True))
    fh = h.family_handler

    def make(arch):
        cfg = synth.WAN_CONFIGS[h.ARCHS[arch][0]]
        sds = [synth.make_wan_state_dict(cfg, s) for s in (0, 1)]
        vsd = synth.make_vae_state_dict(synth.VAE_CFG_TINY, 0, encoder=True)
        model_def = json.load(open(os.path.join(ROOT, "plugin", "defaults", arch + ".json")))["model"]
        return fh.load_model(["hi", "lo"], arch, arch, model_def, dtype=torch.bfloat16, VAE_dtype=torch.float32, submodel_no_list=[1, 2],
                             text_encoder_filename=None, profile=1, lm_decoder_engine="legacy", text_encoder=fake_t5, state_dicts=sds,
                             vae_state_dict=vsd, vae_cfg=synth.VAE_CFG_TINY, device="cpu")
    return fh, make


def test_family_handler_contract():
    fh = load_handler().family_handler
    for name in ("query_supported_types", "query_family_maps", "query_model_family", "query_family_infos", "query_model_def",
                 "query_model_files", "load_model", "fix_settings", "update_default_settings", "validate_generative_settings",
                 "set_cache_parameters", "get_lora_dir", "register_lora_cli_args", "get_rgb_factors"):
        assert callable(getattr(fh, name)), name                    # wan_handler.py:72-1500 / wgp.py:2717-2735, 3153, 3213, 3659, 4069
    types = fh.query_supported_types()
    assert set(types) == {"b200_t2v_1.3B", "b200_t2v_2_2", "b200_i2v_2_2"}
    for t in types:                                                  # every type has a defaults/*.json whose architecture names it
        d = json.load(open(os.path.join(ROOT, "plugin", "defaults", t + ".json")))
        assert d["model"]["architecture"] == t and all("quanto" not in u for u in d["model"]["URLs"])
        md = fh.query_model_def(t, d["model"])
        assert md["multiple_submodels"] == ("URLs2" in d["model"]) and ("unipc", "unipc") in md["sample_solvers"]
    assert fh.query_model_family() in fh.query_family_infos()
    ui = {"sample_solver": ""}
    fh.fix_settings("b200_t2v_2_2", 2.0, {}, ui)
    assert ui["sample_solver"] == "unipc"
    assert fh.validate_generative_settings("b200_t2v_2_2", {}, {"NAG_scale": 2}) is not None
    assert fh.validate_generative_settings("b200_t2v_2_2", {}, {}) is None


def test_generate_t2v_contract(pipeline):
    fh, make = pipeline
    pipe_obj, pipe = make("b200_t2v_2_2")
    assert set(pipe) == {"transformer", "transformer2", "vae"} and all(isinstance(m, torch.nn.Module) for m in pipe.values())
    assert pipe_obj.model is pipe["transformer"] and pipe_obj.model2 is pipe["transformer2"] and pipe_obj._interrupt is False
    events = []

    def callback(step=-1, latents=None, force=True, read_state=False, override_num_inference_steps=-1, pass_no=-1, denoising_extra=""):
        """signature of the progress callback wgp.py builds (build_callback, wgp.py:4172-4178)."""
        kw = {"override_num_inference_steps": override_num_inference_steps, "denoising_extra": denoising_extra, "read_state": read_state}
        events.append((step, None if latents is None else tuple(latents.shape), force, kw))
    out = pipe_obj.generate(**wgp_kwargs(callback=callback, return_latent_slice=slice(0, 1)))
    assert set(out) == {"x", "latent_slice"}
    x = out["x"]
    assert x.dtype == torch.uint8 and x.device.type == "cpu" and tuple(x.shape) == (3, 5, 32, 48)
    assert tuple(out["latent_slice"].shape) == (1, 16, 1, 4, 6)
    # callback protocol: init (-1, None, True) twice, per-block polls (-1, None, False, True) from the model, one (i, latent, False) per step
    assert events[0][:3] == (-1, None, True) and events[1][3]["override_num_inference_steps"] == 4
    steps = [e for e in events if e[0] >= 0 and e[1] is not None]
    assert [e[0] for e in steps] == [0, 1, 2, 3] and all(e[1] == (16, 2, 4, 6) for e in steps)
    assert any("Low Noise" in e[3].get("denoising_extra", "") for e in events)            # expert switch at t <= 875 was announced
    # second call with the full kwarg set and no callback: the contract holds without progress reporting
    assert pipe_obj.generate(**wgp_kwargs(seed=-1, sample_solver="euler"))["x"].shape == (3, 5, 32, 48)


def test_generate_interrupt_and_errors(pipeline):
    fh, make = pipeline
    pipe_obj, _ = make("b200_t2v_2_2")
    n = {"polls": 0}

    def callback(step=-1, latents=None, force=True, read_state=False, **kw):
        n["polls"] += 1
        if n["polls"] == 6:
            pipe_obj._interrupt = True                              # the UI thread's abort (wgp.py:1628)
    assert pipe_obj.generate(**wgp_kwargs(callback=callback)) is None
    pipe_obj._interrupt = False
    for bad in (dict(input_frames=torch.zeros(3, 5, 32, 48)), dict(NAG_scale=2.0), dict(apg_switch=1), dict(image_end=torch.zeros(3, 32, 48)),
                dict(video_prompt_type="GUV"), dict(image_start=torch.zeros(3, 32, 48))):
        with pytest.raises(NotImplementedError):
            pipe_obj.generate(**wgp_kwargs(**bad))
    with pytest.raises(NotImplementedError):
        pipe_obj.generate(**wgp_kwargs(sample_solver="unipc_hf"))


def test_generate_i2v_contract(pipeline):
    fh, make = pipeline
    pipe_obj, pipe = make("b200_i2v_2_2")
    assert pipe_obj.i2v
    img = torch.rand(3, 32, 48) * 2 - 1
    out = pipe_obj.generate(**wgp_kwargs(model_type="b200_i2v_2_2", image_start=img, shift=5.0, guide_scale=3.5, guide2_scale=3.5,
                                         switch_threshold=900, joint_pass=False))
    assert out["x"].shape == (3, 5, 32, 48) and out["latent_slice"] is None
    # continue-video conditioning: 5 prefix frames -> 2 conditioned latent frames
    out = pipe_obj.generate(**wgp_kwargs(model_type="b200_i2v_2_2", input_video=torch.rand(3, 5, 32, 48) * 2 - 1, frame_num=9))
    assert out["x"].shape == (3, 9, 32, 48)
    y, h, w = pipe_obj._i2v_condition(img, None, 9, 32, 48, 0, 1.0)
    assert tuple(y.shape) == (20, 3, 4, 6) and (h, w) == (32, 48)
    assert float(y[:4, 0].min()) == 1.0 and float(y[:4, 1:].abs().max()) == 0.0          # mask: first latent frame given, rest to generate
