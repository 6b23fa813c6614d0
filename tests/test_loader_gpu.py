"""The reference's loader entry point end to end (ref models/visualcla/modeling_utils.py:83-141, scripts/inference/inference.py:55-64):
get_model_and_tokenizer_and_processor on a merged-layout directory with a REAL LlamaTokenizer (a sentencepiece model trained here: no
tokenizer files exist offline) and CLIPImageProcessor, then chat() from a PIL image -- the exact call sequence of inference.py -- on
the CUDA path.  Also the unmerged constructor (base text + vision checkpoints) and its coverage check."""
import os
import shutil

import pytest
import torch

import visualcla_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def merged_dir(tmp_path_factory):
    import sentencepiece as spm
    import visualcla
    from transformers import CLIPImageProcessor, LlamaTokenizer
    root = tmp_path_factory.mktemp("merged")
    tokdir = root / "tok"
    tokdir.mkdir()
    corpus = "\n".join(["the quick brown fox jumps over the lazy dog", "a picture of a cat sitting on a mat", "describe the image in detail please",
                        "what colour is the car", "图片里有什么", "请描述这张图片", "hello world this is a test of the tokenizer"] * 20)
    (tokdir / "corpus.txt").write_text(corpus)
    spm.SentencePieceTrainer.train(input=str(tokdir / "corpus.txt"), model_prefix=str(tokdir / "tokenizer"), vocab_size=400, model_type="bpe",
                                   character_coverage=1.0, bos_id=1, eos_id=2, unk_id=0, pad_id=-1, byte_fallback=True, minloglevel=2)
    tok = LlamaTokenizer.from_pretrained(str(tokdir))
    tok.add_special_tokens({"additional_special_tokens": ["<img>", "</img>", "<pad>", "<img_token>"]})     # ref order, modeling_utils.py:95
    cfg = O.PathConfig(**dict(O.tiny_config().to_dict(), t_vocab=len(tok)))
    model = visualcla.VisualCLAModel.from_synthetic(cfg.to_dict(), seed=3, max_batch=1, max_seq=256)
    out = root / "visualcla-tiny"
    model.save_merged_pretrained(str(out))
    tok.save_pretrained(str(out))
    proc = CLIPImageProcessor(size={"shortest_edge": cfg.v_image}, crop_size={"height": cfg.v_image, "width": cfg.v_image})
    proc.save_pretrained(str(out))
    proc.save_pretrained(str(out / "vision_encoder"))
    return str(out), cfg, model


def _image(seed=0):
    from PIL import Image
    import numpy as np
    rng = np.random.default_rng(seed)
    return Image.fromarray(rng.integers(0, 255, (50, 70, 3), dtype=np.uint8))


def test_reference_loader_and_chat_call_sequence(merged_dir):
    import visualcla
    from transformers import GenerationConfig
    path, cfg, original = merged_dir
    # inference.py:55-64
    model, tokenizer, image_processor = visualcla.get_model_and_tokenizer_and_processor(
        visualcla_model=path, torch_dtype=torch.float16, default_device=None, device_map=None, load_in_8bit=False, max_batch=1, max_seq=256)
    model.float().eval()                                           # inference.py:78-80 on a non-CUDA default; no-ops here
    s0, s1, s2, s3 = O.special_ids(cfg)
    assert (tokenizer.img_start_token_id, tokenizer.img_end_token_id, tokenizer.img_token_id) == (s0, s1, s3)
    assert tokenizer.convert_tokens_to_ids("<pad>") == s2 and model.image_at_head is False and model.num_patch == cfg.r_queries
    sd_a, sd_b = original.state_dict(), model.state_dict()
    assert set(sd_a) == set(sd_b) and all(torch.equal(sd_a[k], sd_b[k]) for k in sd_a)
    # inference.py:110
    gc = GenerationConfig(do_sample=False, max_new_tokens=8, eos_token_id=None, pad_token_id=s2)
    img = _image()
    response, history = visualcla.chat(model, image=img, text="describe the image", history=[], generation_config=gc)
    assert isinstance(response, str) and history[0]["first_instruction"] and history[-1] == {"type": "response", "value": response}
    # the same request assembled by hand gives the same tokens
    from visualcla.modeling_utils import encoding_text
    enc = encoding_text([], "describe the image", model.num_patch, tokenizer)
    px = image_processor(img, return_tensors="pt").pixel_values
    ids = enc.input_ids
    assert int((ids == s3).sum()) == cfg.r_queries and int(ids[0, 0]) == tokenizer.bos_token_id
    out = model.generate(input_ids=ids.cuda(), attention_mask=enc.attention_mask.cuda(), pixel_values=px.cuda().half(), generation_config=gc)
    assert tokenizer.decode(out[0], skip_special_tokens=True) == response
    # and the oracle agrees on the first (decisive) token
    w = {k: v.float() for k, v in sd_b.items()}
    o_tok, o_log = O.generate_greedy(w, cfg, ids, px.float(), 2, image_at_head=False)
    top2 = o_log[0, 0].topk(2).values
    if float(top2[0] - top2[1]) > 0.05 * float(o_log.abs().max()):
        assert int(out[0, 0]) == int(o_tok[0, 0])
    # second turn reuses the history; default (sampling) config runs through the device sampler
    r2, h2 = visualcla.chat(model, image=img, text="and the background?", history=history, generation_config=gc)
    assert len(h2) == 4
    from visualcla.modeling_utils import DEFAULT_GENERATION_CONFIG
    import copy
    dgc = copy.deepcopy(DEFAULT_GENERATION_CONFIG)
    dgc.max_new_tokens = 12
    r3, _ = visualcla.chat(model, image=img, text="again", history=[], generation_config=dgc)
    assert isinstance(r3, str)


def test_unmerged_constructor_and_coverage_check(merged_dir, tmp_path):
    import visualcla
    path, cfg, original = merged_dir
    model, tokenizer, _ = visualcla.get_model_and_tokenizer_and_processor(
        text_model=os.path.join(path, "text_encoder"), vision_model=os.path.join(path, "vision_encoder"), lora_model=path,
        torch_dtype=torch.float16, default_device=None, device_map=None, load_in_8bit=False, max_batch=1, max_seq=64)
    a, b = original.state_dict(), model.state_dict()
    for k in a:
        if k.startswith(("text_model.", "vision_model.")):
            assert torch.equal(a[k], b[k]), k
    assert not torch.equal(a["visual_resampler.query_embeddding"], b["visual_resampler.query_embeddding"]), "resampler: fresh init (ref :245-255)"
    # a base checkpoint that lacks a tensor must fail loudly instead of leaving a layer at its random initialisation
    broken = tmp_path / "text_encoder"
    shutil.copytree(os.path.join(path, "text_encoder"), broken)
    sd = torch.load(broken / "pytorch_model.bin", weights_only=True)
    sd.pop("model.layers.1.mlp.down_proj.weight")
    torch.save(sd, broken / "pytorch_model.bin")
    with pytest.raises(RuntimeError, match="lack"):
        visualcla.VisualCLAModel.from_vision_text_pretrained(os.path.join(path, "vision_encoder"), str(broken), visualcla_config=path,
                                                             max_batch=1, max_seq=64)
    # a checkpoint whose vocabulary disagrees with config.json is an error, not an out-of-bounds read
    sd = torch.load(os.path.join(path, "text_encoder", "pytorch_model.bin"), weights_only=True)
    sd["lm_head.weight"] = sd["lm_head.weight"][:-4].clone()
    torch.save(sd, broken / "pytorch_model.bin")
    with pytest.raises(ValueError, match="shape mismatch"):
        visualcla.VisualCLAModel.from_vision_text_pretrained(os.path.join(path, "vision_encoder"), str(broken), visualcla_config=path,
                                                             max_batch=1, max_seq=64)
