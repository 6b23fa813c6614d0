"""Generate tests/golden/*.npz by running the UNMODIFIED reference (behind oracle/ref_shim.py)
on seeded synthetic weights.  Run in the authoring container only:

    python oracle/gen_golden.py

The fixtures pin oracle/visualcla_oracle.py (tests/test_oracle_golden.py) and, through it and
directly, the CUDA path (tests/test_parity_gpu.py).  Weights are NOT stored: they are
regenerated bit-exactly from (config, seed) by the integer-hash generator
(oracle: hash_normal_bf16; device: csrc/elementwise.cu fill_hash_normal_kernel).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import visualcla_oracle as O  # noqa: E402
from ref_shim import import_reference, RESAMPLER_EXTRA  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def build_reference_model(visualcla, cfg: O.PathConfig, weights):
    from transformers import LlamaConfig
    from transformers.models.clip.modeling_clip import CLIPVisionConfig
    tcfg = LlamaConfig(vocab_size=cfg.t_vocab, hidden_size=cfg.t_hidden, intermediate_size=cfg.t_ffn,
                       num_hidden_layers=cfg.t_layers, num_attention_heads=cfg.t_heads,
                       num_key_value_heads=cfg.t_heads, rms_norm_eps=cfg.t_eps, rope_theta=cfg.rope_theta,
                       max_position_embeddings=2048, tie_word_embeddings=False, attn_implementation="eager",
                       pad_token_id=0, bos_token_id=1, eos_token_id=2)
    vcfg = CLIPVisionConfig(hidden_size=cfg.v_hidden, intermediate_size=cfg.v_ffn, num_hidden_layers=cfg.v_layers,
                            num_attention_heads=cfg.v_heads, image_size=cfg.v_image, patch_size=cfg.v_patch,
                            hidden_act="quick_gelu", layer_norm_eps=cfg.v_eps, attn_implementation="eager")
    rcfg = dict(hidden_size=cfg.r_hidden, num_hidden_layers=cfg.r_layers, num_attention_heads=cfg.r_heads,
                intermediate_size=cfg.r_ffn, hidden_act="gelu", layer_norm_eps=cfg.r_eps,
                num_query_tokens=cfg.r_queries, **RESAMPLER_EXTRA)
    tdict, vdict = tcfg.to_dict(), vcfg.to_dict()
    tdict["attn_implementation"] = "eager"
    vdict["attn_implementation"] = "eager"
    config = visualcla.VisualCLAConfig(text_config=tdict, vision_config=vdict,
                                       use_visual_resampler=True, visual_resampler_config=rcfg)
    model = visualcla.VisualCLAModel(config)
    sd = model.state_dict()
    missing = [k for k in sd if k not in weights and "pooler" not in k and "position_ids" not in k]
    extra = [k for k in weights if k not in sd]
    assert not missing and not extra, (missing, extra)
    with torch.no_grad():
        for k, v in weights.items():
            assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
            sd[k].copy_(v)
    model.float().eval()
    try:
        model.text_model.config._attn_implementation = "eager"
        model.vision_model.config._attn_implementation = "eager"
    except Exception:
        pass
    return model


@torch.no_grad()
def case_tiny(visualcla, name, seed, batch, t_text, n_new):
    cfg = O.tiny_config()
    w = O.make_weights(cfg, seed)
    model = build_reference_model(visualcla, cfg, w)
    pixels, ids = O.make_inputs(cfg, batch, t_text, seed=1234 + seed)
    s0, s1, s2, s3 = O.special_ids(cfg)
    mask = torch.ones_like(ids)

    # stage activations (ref: modeling_visualcla.py:283-288)
    vit = model.vision_model(pixel_values=pixels)[0]
    post = model.vision_model.vision_model.post_layernorm(vit)
    res = model.visual_resampler(encoder_hidden_states=post).last_hidden_state
    proj = model.image_projection_layer(res)

    # forward(), image_at_head=True (constructor default, ref: modeling_visualcla.py:108);
    # labels must be non-None in this layout (ref quirk, :313-315)
    model.image_at_head = True
    out = model(input_ids=ids, pixel_values=pixels, attention_mask=mask, labels=ids, return_dict=True)
    logits_head = out.logits
    loss_head = out.loss                      # labels get -100 over the image block (ref: modeling_visualcla.py:313-315)
    labels_masked = ids.clone()
    labels_masked[:, 3::3] = -100             # some ignored text positions as well
    loss_head_masked = model(input_ids=ids, pixel_values=pixels, attention_mask=mask, labels=labels_masked, return_dict=True).loss

    # placeholder layout (loader default, ref: modeling_utils.py:134)
    nq = cfg.r_queries
    ids_ph = torch.cat([ids[:, :2], torch.full((batch, nq), s3, dtype=torch.long), ids[:, 2:]], dim=1)
    model.image_at_head = False
    model.tokenizer = types.SimpleNamespace(img_start_token_id=s0, img_end_token_id=s1, img_token_id=s3)
    out_ph = model(input_ids=ids_ph, pixel_values=pixels, attention_mask=torch.ones_like(ids_ph), return_dict=True)
    logits_ph = out_ph.logits
    loss_ph = model(input_ids=ids_ph, pixel_values=pixels, attention_mask=torch.ones_like(ids_ph), labels=ids_ph, return_dict=True).loss

    # text only
    out_txt = model(input_ids=ids, pixel_values=None, attention_mask=mask, labels=ids, return_dict=True)

    # greedy generate, eos disabled, returns only new tokens (ref: modeling_visualcla.py:333-392)
    from transformers import GenerationConfig
    gc = GenerationConfig(do_sample=False, max_new_tokens=n_new, eos_token_id=None, pad_token_id=0, bos_token_id=1,
                          output_logits=True, return_dict_in_generate=True)
    model.image_at_head = True
    gen = model.generate(input_ids=ids, pixel_values=pixels, attention_mask=mask, generation_config=gc)
    gen_tokens = gen.sequences
    gen_logits = torch.stack(list(gen.logits), dim=1)
    model.image_at_head = False
    gen_ph = model.generate(input_ids=ids_ph, pixel_values=pixels, attention_mask=torch.ones_like(ids_ph),
                            generation_config=gc)
    assert torch.equal(gen_ph.sequences, gen_tokens), "layout equivalence broken in the reference?"

    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"),
        config=np.array(repr(cfg.to_dict())), seed=np.array(seed), batch=np.array(batch), t_text=np.array(t_text),
        pixel_values=pixels.numpy(), input_ids=ids.numpy(), input_ids_placeholder=ids_ph.numpy(),
        vit_out=vit.numpy(), post_ln=post.numpy(), resampler_out=res.numpy(), projector_out=proj.numpy(),
        logits_at_head=logits_head.numpy(), logits_placeholder=logits_ph.numpy(), logits_text_only=out_txt.logits.numpy(),
        gen_tokens=gen_tokens.numpy(), gen_logits=gen_logits.float().numpy(),
        loss_at_head=loss_head.numpy(), labels_masked=labels_masked.numpy(), loss_at_head_masked=loss_head_masked.numpy(),
        loss_placeholder=loss_ph.numpy(), loss_text_only=out_txt.loss.numpy(),
    )
    print(f"[golden] {name}: logits {tuple(logits_head.shape)} gen {tuple(gen_tokens.shape)} "
          f"max|logit| {logits_head.abs().max():.3f} layout diff {float((logits_head - logits_ph).abs().max()):.2e}")


@torch.no_grad()
def case_resampler_fullwidth(visualcla, seed=3):
    """The only in-repo arithmetic at its real width: 6 layers, 1024 hidden, 16 heads, 64 queries over
    257 image tokens (ref: modeling_visual_resampler.py:609-737)."""
    from ref_shim import RESAMPLER_EXTRA
    cfg = O.PathConfig()
    specs = [s for s in O.weight_specs(cfg) if s[0].startswith("visual_resampler.")]
    w = {n: O.hash_normal_bf16(n, int(np.prod(sh)), std, seed, mean).reshape(sh) for n, sh, std, mean in specs}
    from visualcla.modeling_visual_resampler import VisualResamplerConfig, VisualResamplerModel
    rc = VisualResamplerConfig(hidden_size=cfg.r_hidden, num_hidden_layers=cfg.r_layers, num_attention_heads=cfg.r_heads,
                               intermediate_size=cfg.r_ffn, layer_norm_eps=cfg.r_eps, num_query_tokens=cfg.r_queries,
                               **RESAMPLER_EXTRA)
    m = VisualResamplerModel(rc).float().eval()
    sd = m.state_dict()
    for k, v in w.items():
        sd[k[len("visual_resampler."):]].copy_(v)
    x = O.hash_normal_bf16("resampler_input", 2 * cfg.v_tokens * cfg.r_hidden, 1.0, seed).reshape(2, cfg.v_tokens, cfg.r_hidden)
    y = m(encoder_hidden_states=x).last_hidden_state
    np.savez_compressed(os.path.join(OUT, "resampler_fullwidth.npz"), seed=np.array(seed),
                        out=y.numpy().astype(np.float32))
    print(f"[golden] resampler_fullwidth: out {tuple(y.shape)} absmax {float(y.abs().max()):.3f}")


@torch.no_grad()
def case_tiny_padded(visualcla, seed=2):
    """Left-padded batch of different prompt lengths through the reference's generate() (placeholder layout, what the loader
    configures): HF derives position_ids from the attention mask and masks the pad keys."""
    from transformers import GenerationConfig
    cfg = O.tiny_config()
    w = O.make_weights(cfg, seed)
    model = build_reference_model(visualcla, cfg, w)
    s0, s1, s2, s3 = O.special_ids(cfg)
    B, T, nq, n_new = 3, 14, cfg.r_queries, 6
    pads = [0, 3, 6]
    pixels, raw = O.make_inputs(cfg, B, T, seed=1234 + seed)
    S = T + nq
    ids = torch.full((B, S), s2, dtype=torch.long)          # <pad>
    mask = torch.zeros(B, S, dtype=torch.long)
    for b, p in enumerate(pads):
        body = torch.cat([raw[b, :2], torch.full((nq,), s3, dtype=torch.long), raw[b, 2:T - p]])   # [bos,<img>,8x<img_token>,</img>,text...]
        ids[b, S - body.numel():] = body
        mask[b, S - body.numel():] = 1
        assert S - body.numel() == p
    model.image_at_head = False
    model.tokenizer = types.SimpleNamespace(img_start_token_id=s0, img_end_token_id=s1, img_token_id=s3)
    gc = GenerationConfig(do_sample=False, max_new_tokens=n_new, eos_token_id=None, pad_token_id=s2, bos_token_id=1,
                          output_logits=True, return_dict_in_generate=True)
    gen = model.generate(input_ids=ids, pixel_values=pixels, attention_mask=mask, generation_config=gc)
    fwd = model(input_ids=ids, pixel_values=pixels, attention_mask=mask, return_dict=True).logits
    np.savez_compressed(os.path.join(OUT, "tiny_padded.npz"), config=np.array(repr(cfg.to_dict())), seed=np.array(seed),
                        pixel_values=pixels.numpy(), input_ids=ids.numpy(), attention_mask=mask.numpy(), pads=np.array(pads),
                        gen_tokens=gen.sequences.numpy(), gen_logits=torch.stack(list(gen.logits), 1).float().numpy(),
                        forward_logits=fwd.numpy())
    print(f"[golden] tiny_padded: pads {pads} gen {tuple(gen.sequences.shape)}")


def case_host_logic(visualcla):
    """Prompt strings (ref: modeling_utils.py:49-80) and the extra samplers (ref: :250-320) for the host-side tests."""
    import json
    from visualcla import modeling_utils as mu

    class Tok:
        bos_token, img_start_token, img_end_token, img_token = "<s>", "<img>", "</img>", "<img_token>"

        def __call__(self, text, return_tensors=None, add_special_tokens=None):
            return text

    h0 = []
    h1 = [{"type": "instruction", "value": "What is in the picture?", "first_instruction": True},
          {"type": "response", "value": " A cat."}]
    h2 = h1 + [{"type": "instruction", "value": "What colour?"}, {"type": "response", "value": " Black and white."}]
    cases = []
    for hist, text, n in ((h0, "Describe the image.", 64), (h1, "And the background?", 64), (h2, "Thanks!", 8)):
        cases.append({"history": hist, "text": text, "num_patch": n, "prompt": mu.encoding_text(hist, text, n, Tok())})
    with open(os.path.join(OUT, "prompts.json"), "w") as f:
        json.dump(cases, f, ensure_ascii=False, indent=1)
    g = torch.Generator().manual_seed(11)
    logits = torch.randn(3, 500, generator=g) * 3.0
    tfs = mu.TailFreeLogitsWarper(tfs=0.9)(None, logits.clone())
    topa = mu.TopALogitsWarper(top_a=0.2)(None, logits.clone())
    # the logits-processor chain HF's generate() builds for the reference's DEFAULT_GENERATION_CONFIG (ref: modeling_utils.py:36-47:
    # repetition_penalty 1.1, no_repeat_ngram_size 15, temperature 0.5, top_k 40, top_p 0.9), applied to a generated-token history
    # (with inputs_embeds the processors only ever see the NEW tokens).  Two histories: one with a repeated n-gram (n = 3 so that a
    # short history exercises the ban) and the default n = 15.
    from transformers.generation import logits_process as lp
    V = 2000
    big = torch.randn(4, V, generator=g) * 4.0
    hist = torch.randint(0, V, (4, 24), generator=g)
    hist[0, 10:12] = hist[0, 2:4]; hist[0, 22:24] = hist[0, 2:4]     # row 0: "... a b X ... a b Y ... a b" -> X and Y banned at n = 3
    hist[1, 20:24] = hist[1, 5:9]
    chain = {}
    for n_gram in (3, 15):
        x = big.clone()
        x = lp.RepetitionPenaltyLogitsProcessor(penalty=1.1)(hist, x); chain[f"n{n_gram}_rep"] = x.clone()
        x = lp.NoRepeatNGramLogitsProcessor(n_gram)(hist, x); chain[f"n{n_gram}_ngram"] = x.clone()
        x = lp.TemperatureLogitsWarper(0.5)(hist, x); chain[f"n{n_gram}_temp"] = x.clone()
        x = lp.TopKLogitsWarper(top_k=40, min_tokens_to_keep=1)(hist, x); chain[f"n{n_gram}_topk"] = x.clone()
        x = lp.TopPLogitsWarper(top_p=0.9, min_tokens_to_keep=1)(hist, x); chain[f"n{n_gram}_topp"] = x.clone()
    np.savez_compressed(os.path.join(OUT, "samplers.npz"), logits=logits.numpy(), tfs_0p9=tfs.numpy(), top_a_0p2=topa.numpy(),
                        chain_logits=big.numpy(), chain_history=hist.numpy(), **{k: v.numpy() for k, v in chain.items()})
    print(f"[golden] host logic: {len(cases)} prompts; tfs keeps {int(torch.isfinite(tfs).sum())}, top_a keeps {int(torch.isfinite(topa).sum())}")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    visualcla = import_reference()
    case_tiny(visualcla, "tiny_b2_t12", seed=0, batch=2, t_text=12, n_new=8)
    case_tiny(visualcla, "tiny_b3_t7", seed=1, batch=3, t_text=7, n_new=5)
    case_resampler_fullwidth(visualcla)
    case_tiny_padded(visualcla)
    case_host_logic(visualcla)


if __name__ == "__main__":
    main()
