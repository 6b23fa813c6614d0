"""Pin the CPU oracle (oracle/visualcla_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, made by oracle/gen_golden.py).  CPU only."""
import ast
import os

import numpy as np
import pytest
import torch

import visualcla_oracle as O

CASES = ["tiny_b2_t12", "tiny_b3_t7"]
TOL = 2e-4   # fp32 vs fp32, different op order (fused qkv etc.): observed < 2e-5


def _load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = O.PathConfig(**ast.literal_eval(str(g["config"])))
    return g, cfg


def _close(a, b, tol=TOL):
    a = torch.as_tensor(np.asarray(a)).float()
    b = torch.as_tensor(np.asarray(b)).float()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    ref = max(1.0, b.abs().max().item())
    assert err <= tol * ref, f"max abs err {err:.3e} (ref scale {ref:.3f})"


@pytest.mark.parametrize("name", CASES)
def test_inputs_regenerate(golden_dir, name):
    g, cfg = _load(golden_dir, name)
    px, ids = O.make_inputs(cfg, int(g["batch"]), int(g["t_text"]), seed=1234 + int(g["seed"]))
    assert np.array_equal(px.numpy(), g["pixel_values"])
    assert np.array_equal(ids.numpy(), g["input_ids"])


@pytest.mark.parametrize("name", CASES)
def test_stages_and_logits(golden_dir, name):
    g, cfg = _load(golden_dir, name)
    w = O.make_weights(cfg, int(g["seed"]))
    px = torch.from_numpy(g["pixel_values"])
    ids = torch.from_numpy(g["input_ids"])
    st = {}
    logits = O.forward_logits(w, cfg, ids, px, image_at_head=True, stages=st)
    for k in ("vit_out", "post_ln", "resampler_out", "projector_out"):
        _close(st[k], g[k])
    _close(logits, g["logits_at_head"])
    # placeholder layout == at-head layout
    lp = O.forward_logits(w, cfg, torch.from_numpy(g["input_ids_placeholder"]), px, image_at_head=False)
    _close(lp, g["logits_placeholder"])
    lt = O.forward_logits(w, cfg, ids, None)
    _close(lt, g["logits_text_only"])


@pytest.mark.parametrize("name", CASES)
def test_loss_against_reference(golden_dir, name):
    """forward(labels=...).loss of the unmodified reference (incl. its -100 fill over the image block, ref
    modeling_visualcla.py:313-315) vs the oracle's restatement, in all three layouts and with ignored label positions."""
    g, cfg = _load(golden_dir, name)
    w = O.make_weights(cfg, int(g["seed"]))
    px, ids = torch.from_numpy(g["pixel_values"]), torch.from_numpy(g["input_ids"])
    lg = O.forward_logits(w, cfg, ids, px, image_at_head=True)
    assert abs(float(O.causal_lm_loss(lg, ids, cfg.r_queries)) - float(g["loss_at_head"])) <= 1e-4
    assert abs(float(O.causal_lm_loss(lg, torch.from_numpy(g["labels_masked"]), cfg.r_queries)) - float(g["loss_at_head_masked"])) <= 1e-4
    ids_ph = torch.from_numpy(g["input_ids_placeholder"])
    assert abs(float(O.causal_lm_loss(O.forward_logits(w, cfg, ids_ph, px, image_at_head=False), ids_ph)) - float(g["loss_placeholder"])) <= 1e-4
    assert abs(float(O.causal_lm_loss(O.forward_logits(w, cfg, ids, None), ids)) - float(g["loss_text_only"])) <= 1e-4


@pytest.mark.parametrize("name", CASES)
def test_greedy_generate(golden_dir, name):
    g, cfg = _load(golden_dir, name)
    w = O.make_weights(cfg, int(g["seed"]))
    px = torch.from_numpy(g["pixel_values"])
    ids = torch.from_numpy(g["input_ids"])
    n = g["gen_tokens"].shape[1]
    toks, logits = O.generate_greedy(w, cfg, ids, px, n, image_at_head=True)
    assert np.array_equal(toks.numpy(), g["gen_tokens"])
    _close(logits, g["gen_logits"])


def test_resampler_fullwidth(golden_dir):
    g = np.load(os.path.join(golden_dir, "resampler_fullwidth.npz"))
    seed = int(g["seed"])
    cfg = O.PathConfig()
    w = {n: O.hash_normal_bf16(n, int(np.prod(sh)), std, seed, mean).reshape(sh)
         for n, sh, std, mean in O.weight_specs(cfg) if n.startswith("visual_resampler.")}
    x = O.hash_normal_bf16("resampler_input", 2 * cfg.v_tokens * cfg.r_hidden, 1.0, seed).reshape(2, cfg.v_tokens, cfg.r_hidden)
    y = O.resampler_forward(w, cfg, x)
    _close(y, g["out"])


def test_splice_errors():
    cfg = O.tiny_config()
    w = {"text_model.model.embed_tokens.weight": torch.zeros(cfg.t_vocab, cfg.t_hidden)}
    s0, s1, _, s3 = O.special_ids(cfg)
    ids = torch.tensor([[1, s0, s3, s3, s1, 5]])     # only 2 placeholders for 8 queries
    with pytest.raises((ValueError, IndexError)):
        O.splice(w, cfg, ids, torch.zeros(1, cfg.r_queries, cfg.t_hidden), False, s0, s1, s3)


def _padded_inputs(g, cfg):
    w = O.make_weights(cfg, int(g["seed"]))
    px = torch.from_numpy(g["pixel_values"])
    ids = torch.from_numpy(g["input_ids"])
    pads = torch.from_numpy(g["pads"])
    s0, s1, _, s3 = O.special_ids(cfg)
    img = O.vision_encode(w, cfg, px)
    x = O.splice(w, cfg, ids, img, False, s0, s1, s3)
    return w, x, pads


def test_left_padded_batch_generate_and_forward(golden_dir):
    """Left-padded prompts of different lengths (tests/golden/tiny_padded.npz from the reference's generate()/forward())."""
    g, cfg = _load(golden_dir, "tiny_padded")
    w, x, pads = _padded_inputs(g, cfg)
    # forward(): the reference passes no position_ids -> arange positions, pad keys masked; compare only real rows
    lf = O.llama_forward(w, cfg, x, left_pad=pads, pos_from_mask=False)
    ref = torch.from_numpy(g["forward_logits"])
    for b, p in enumerate(pads.tolist()):
        _close(lf[b, p:], ref[b, p:])
    # generate(): positions from the mask
    n = g["gen_tokens"].shape[1]
    cache = O.KVCache(cfg.t_layers)
    logits = O.llama_forward(w, cfg, x, cache, last_only=True, left_pad=pads)[:, -1]
    toks, logs = [], []
    for step in range(n):
        nxt = logits.argmax(-1)
        toks.append(nxt); logs.append(logits)
        if step == n - 1:
            break
        e = w["text_model.model.embed_tokens.weight"][nxt].unsqueeze(1)
        logits = O.llama_forward(w, cfg, e, cache, last_only=True, left_pad=pads)[:, -1]
    assert np.array_equal(torch.stack(toks, 1).numpy(), g["gen_tokens"])
    _close(torch.stack(logs, 1), g["gen_logits"])
