"""Host-side logic of the drop-in package against fixtures produced by the unmodified reference
(tests/golden/prompts.json, samplers.npz) + pure-python checks.  CPU only."""
import json
import os
import types

import numpy as np
import pytest
import torch


def test_prompt_builder_matches_reference(golden_dir):
    from visualcla import modeling_utils as mu

    class Tok:
        bos_token, img_start_token, img_end_token, img_token = "<s>", "<img>", "</img>", "<img_token>"

        def __call__(self, text, return_tensors=None, add_special_tokens=None):
            assert add_special_tokens is False
            return text

    for case in json.load(open(os.path.join(golden_dir, "prompts.json"))):
        got = mu.encoding_text(case["history"], case["text"], case["num_patch"], Tok())
        assert got == case["prompt"]
    with pytest.raises(ValueError):
        mu.encoding_text([{"type": "system", "value": "x"}], "t", 4, Tok())


def test_samplers_match_reference(golden_dir):
    from visualcla import modeling_utils as mu
    g = np.load(os.path.join(golden_dir, "samplers.npz"))
    logits = torch.from_numpy(g["logits"])
    tfs = mu.TailFreeLogitsWarper(tfs=0.9)(None, logits.clone())
    topa = mu.TopALogitsWarper(top_a=0.2)(None, logits.clone())
    assert np.array_equal(torch.isfinite(tfs).numpy(), np.isfinite(g["tfs_0p9"]))
    assert np.array_equal(torch.isfinite(topa).numpy(), np.isfinite(g["top_a_0p2"]))
    assert np.allclose(np.nan_to_num(tfs.numpy(), neginf=-1e9), np.nan_to_num(g["tfs_0p9"], neginf=-1e9))
    with pytest.raises(ValueError):
        mu.TailFreeLogitsWarper(tfs=1.5)


def test_default_generation_config_matches_reference():
    from visualcla.modeling_utils import DEFAULT_GENERATION_CONFIG as g
    # ref: models/visualcla/modeling_utils.py:36-47
    assert (g.max_new_tokens, g.do_sample, g.top_p, g.top_k, g.temperature, g.no_repeat_ngram_size, g.repetition_penalty) == \
        (512, True, 0.9, 40, 0.5, 15, 1.1)


def test_config_roundtrip(tmp_path):
    from visualcla import VisualCLAConfig
    from visualcla.engine import path_config_7b
    p = path_config_7b()
    cfg = VisualCLAConfig.from_path_config(p)
    cfg.save_pretrained(str(tmp_path))
    back = VisualCLAConfig.from_pretrained(str(tmp_path))
    q = back.to_path_config()
    for k in p:
        assert q[k] == pytest.approx(p[k]), k
    assert back.visual_resampler_config["num_query_tokens"] == 64 and back.use_visual_resampler
    cfg.use_visual_resampler = False
    with pytest.raises(NotImplementedError):
        cfg.to_path_config()


def test_public_surface():
    import visualcla
    for name in ("VisualCLAModel", "VisualCLAConfig", "VisualCLAProcessor", "get_model_and_tokenizer_and_processor", "chat",
                 "chat_in_stream", "hijack_samplers"):
        assert hasattr(visualcla, name)
    from visualcla.modeling_utils import DEFAULT_GENERATION_CONFIG  # noqa: F401  (gradio_demo.py:2 imports it)
    m = visualcla.VisualCLAModel
    for meth in ("from_pretrained", "from_merged_pretrained", "from_vision_text_pretrained", "forward", "generate",
                 "get_input_embeddings", "set_input_embeddings", "get_output_embeddings", "set_output_embeddings",
                 "resize_token_embeddings", "state_dict", "eval", "float", "half", "to"):
        assert hasattr(m, meth), meth
    visualcla.hijack_samplers()
    from transformers import GenerationConfig
    assert GenerationConfig().tfs == 1.0


def test_shard_bounds_cover_batch():
    from visualcla.dp import shard_bounds
    for B in (1, 7, 8, 64, 65):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_iteratorize_streams_and_stops():
    from visualcla.modeling_utils import Iteratorize

    def producer(callback=None, n=5):
        for i in range(n):
            callback(i)
        return "done"

    with Iteratorize(producer, {"n": 5}) as it:
        assert list(it) == [0, 1, 2, 3, 4]
    seen = []
    with Iteratorize(producer, {"n": 1000}) as it:
        for v in it:
            seen.append(v)
            if v == 3:
                break
    assert seen == [0, 1, 2, 3]


def test_lora_key_plan_and_fold():
    """Key normalisation of the reference's adapter_model.bin layout (ref convert_ckpt_for_tgwebui.py:46-71) + the fold math."""
    from visualcla import lora
    keys = ["base_model.model.text_model.model.layers.0.self_attn.q_proj.lora_A.weight",
            "base_model.model.text_model.model.layers.0.self_attn.q_proj.lora_B.weight",
            "base_model.model.vision_model.vision_model.encoder.layers.3.mlp.fc1.lora_A.default.weight",
            "base_model.model.vision_model.vision_model.encoder.layers.3.mlp.fc1.lora_B.default.weight",
            "base_model.model.text_model.model.embed_tokens.modules_to_save.default.weight",
            "base_model.model.text_model.lm_head.weight",
            "base_model.model.visual_resampler.query_embeddding",
            "base_model.model.image_projection_layer.bias"]
    pairs, full = lora.plan(keys)
    assert set(pairs) == {"text_model.model.layers.0.self_attn.q_proj.weight", "vision_model.vision_model.encoder.layers.3.mlp.fc1.weight"}
    assert set(full) == {"text_model.model.embed_tokens.weight", "text_model.lm_head.weight", "visual_resampler.query_embeddding",
                         "image_projection_layer.bias"}
    with pytest.raises(ValueError):
        lora.plan(keys[:1])
    g = torch.Generator().manual_seed(0)
    W, A, B = torch.randn(6, 5, generator=g), torch.randn(2, 5, generator=g), torch.randn(6, 2, generator=g)
    out = lora.fold(W.bfloat16(), A, B, scaling=4.0)
    assert torch.allclose(out, W.bfloat16().float() + 4.0 * B @ A, atol=1e-6)
