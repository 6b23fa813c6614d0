"""Device-side sampling stack (csrc/sampler.cu through the C ABI) against the processors HF's generate() builds for the reference's
DEFAULT_GENERATION_CONFIG (ref models/visualcla/modeling_utils.py:36-47): processed scores must have the identical kept set and
equal values; the draw is checked statistically (Philox is not torch's generator) and through its deterministic corners."""
import os

import numpy as np
import pytest
import torch

import visualcla_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAL_TOL = 2e-6


@pytest.fixture(scope="module")
def eng():
    import visualcla
    m = visualcla.VisualCLAModel.from_synthetic(O.tiny_config().to_dict(), seed=0, max_batch=2, max_seq=64)
    return m._engine


def _same_scores(got, want, what):
    got, want = got.float().cpu(), torch.as_tensor(np.asarray(want)).float()
    kg, kw = torch.isfinite(got), torch.isfinite(want)
    assert torch.equal(kg, kw), f"{what}: kept sets differ ({int(kg.sum())} vs {int(kw.sum())} finite entries)"
    err = (got[kg] - want[kw]).abs().max().item() if bool(kg.any()) else 0.0
    assert err <= VAL_TOL * max(1.0, want[kw].abs().max().item()), f"{what}: max abs diff {err:.3e}"


@pytest.mark.parametrize("n_gram", [3, 15])
def test_chain_matches_hf_golden(eng, n_gram):
    g = np.load(os.path.join(ROOT, "tests", "golden", "samplers.npz"))
    logits, hist = torch.from_numpy(g["chain_logits"]), torch.from_numpy(g["chain_history"])
    S = eng.sampler_spec
    _, sc = eng.op_sample(logits, hist, S(do_sample=False, repetition_penalty=1.1))
    _same_scores(sc, g[f"n{n_gram}_rep"], "repetition penalty")
    tok, sc = eng.op_sample(logits, hist, S(do_sample=False, repetition_penalty=1.1, no_repeat_ngram_size=n_gram))
    _same_scores(sc, g[f"n{n_gram}_ngram"], "penalty + no-repeat-ngram")
    assert torch.equal(tok.cpu().long(), torch.from_numpy(g[f"n{n_gram}_ngram"]).argmax(-1)), "greedy over the processed scores"
    _, sc = eng.op_sample(logits, hist, S(do_sample=True, repetition_penalty=1.1, no_repeat_ngram_size=n_gram, temperature=0.5, top_k=40, top_p=1.0, seed=1))
    _same_scores(sc, g[f"n{n_gram}_topk"], "... + temperature + top-k")
    tok, sc = eng.op_sample(logits, hist, S(do_sample=True, repetition_penalty=1.1, no_repeat_ngram_size=n_gram, temperature=0.5, top_k=40, top_p=0.9, seed=1))
    _same_scores(sc, g[f"n{n_gram}_topp"], "full default chain")
    kept = torch.isfinite(torch.from_numpy(g[f"n{n_gram}_topp"]))
    assert bool(kept[torch.arange(4), tok.cpu().long()].all()), "the drawn token belongs to the kept set"
    if n_gram == 3:
        assert not bool(torch.isfinite(sc[0, hist[0, 4]])) and not bool(torch.isfinite(sc[0, hist[0, 12]])), "both continuations of the repeated bigram are banned"


def test_chain_matches_hf_live_at_model_vocab(eng):
    """Fresh random logits at the real vocabulary (49958), long histories with repeats, several knob settings, against transformers'
    own processors run here."""
    from transformers.generation import logits_process as lp
    g = torch.Generator().manual_seed(5)
    V, B, L = 49958, 6, 300
    logits = torch.randn(B, V, generator=g) * 3.0
    hist = torch.randint(0, V, (B, L), generator=g)
    hist[:, 200:230] = hist[:, 20:50]                       # long repeats: n-gram bans fire for n <= 31
    hist[:, -14:] = hist[:, 20:34]                          # ... and the current suffix matches them
    for rp, ng, t, k, p in ((1.1, 15, 0.5, 40, 0.9), (1.3, 4, 1.0, 1, 1.0), (1.0, 0, 0.7, 100, 0.5), (1.2, 2, 1.5, 1000, 0.95), (1.0, 0, 1.0, 5, 0.3)):
        x = logits.clone()
        if rp != 1.0:
            x = lp.RepetitionPenaltyLogitsProcessor(penalty=rp)(hist, x)
        if ng:
            x = lp.NoRepeatNGramLogitsProcessor(ng)(hist, x)
        if t != 1.0:
            x = lp.TemperatureLogitsWarper(t)(hist, x)
        x = lp.TopKLogitsWarper(top_k=k, min_tokens_to_keep=1)(hist, x)
        if p < 1.0:
            x = lp.TopPLogitsWarper(top_p=p, min_tokens_to_keep=1)(hist, x)
        tok, sc = eng.op_sample(logits, hist, eng.sampler_spec(do_sample=True, repetition_penalty=rp, no_repeat_ngram_size=ng, temperature=t, top_k=k, top_p=p, seed=3))
        _same_scores(sc, x, f"rp={rp} ngram={ng} T={t} k={k} p={p}")
        assert bool(torch.isfinite(x)[torch.arange(B), tok.cpu().long()].all())


def test_draw_distribution_and_determinism(eng):
    V, B = 1000, 64
    base = torch.full((V,), -20.0)
    probs = torch.tensor([0.4, 0.25, 0.15, 0.1, 0.06, 0.04])
    ids = torch.tensor([7, 300, 41, 999, 0, 512])
    base[ids] = probs.log()
    logits = base.repeat(B, 1)
    counts = torch.zeros(V)
    n_calls = 60
    for s in range(n_calls):
        tok, _ = eng.op_sample(logits, None, eng.sampler_spec(do_sample=True, top_k=6, seed=1000 + s), return_scores=False)
        counts += torch.bincount(tok.cpu().long(), minlength=V).float()
    n = n_calls * B
    assert counts.sum() == n and counts[ids].sum() == n, "only the top-k tokens are ever drawn"
    freq = counts[ids] / n
    sigma = (probs * (1 - probs) / n).sqrt()
    assert bool(((freq - probs).abs() < 5 * sigma).all()), f"draw frequencies {freq.tolist()} vs probabilities {probs.tolist()}"
    a, _ = eng.op_sample(logits, None, eng.sampler_spec(do_sample=True, top_k=6, seed=42), return_scores=False)
    b, _ = eng.op_sample(logits, None, eng.sampler_spec(do_sample=True, top_k=6, seed=42), return_scores=False)
    c, _ = eng.op_sample(logits, None, eng.sampler_spec(do_sample=True, top_k=6, seed=43), return_scores=False)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert len(set(a.tolist())) > 1, "sequences draw independently (counter = (step, sequence))"
    one, _ = eng.op_sample(torch.randn(5, V), None, eng.sampler_spec(do_sample=True, top_k=1, temperature=0.3, top_p=0.2, seed=9), return_scores=False)
    assert one.shape == (5,)


def _model():
    import visualcla
    cfg = O.PathConfig(v_layers=1, r_layers=1, t_hidden=512, t_heads=4, t_ffn=1408, t_layers=2, t_vocab=2003)
    m = visualcla.VisualCLAModel.from_synthetic(cfg.to_dict(), seed=9, max_batch=4, max_seq=200)
    px, ids = O.make_inputs(cfg, 3, 20, seed=5)
    return m, ids.cuda(), px.cuda()


def test_generate_device_path_equals_host_processor_path(monkeypatch):
    """Deterministic knobs (greedy + repetition penalty + n-gram ban, EOS with padding, top_k=1 'sampling'): the in-graph device
    sampler must produce exactly the tokens of the per-step host path that runs HF's processors on the returned logits."""
    m, ids, px = _model()
    kw = dict(input_ids=ids, pixel_values=px, max_new_tokens=40, pad_token_id=0)
    plain = m.generate(do_sample=False, eos_token_id=None, **kw)
    launches0 = m._engine.kernel_launches(reset=True)
    cases = [dict(do_sample=False, eos_token_id=None, repetition_penalty=1.3, no_repeat_ngram_size=2),
             dict(do_sample=True, top_k=1, temperature=0.7, top_p=0.9, eos_token_id=None, repetition_penalty=1.1, no_repeat_ngram_size=15),
             dict(do_sample=False, eos_token_id=int(plain[0, 5]), repetition_penalty=1.0),
             dict(do_sample=False, eos_token_id=[int(plain[0, 3]), int(plain[1, 9]), int(plain[2, 20])], min_new_tokens=6, repetition_penalty=1.2)]
    for c in cases:
        monkeypatch.delenv("VCLA_HOST_SAMPLER", raising=False)
        dev = m.generate(**c, **kw)
        monkeypatch.setenv("VCLA_HOST_SAMPLER", "1")
        host = m.generate(**c, **kw)
        assert dev.shape == host.shape and torch.equal(dev, host), f"{c}: device {dev.tolist()} vs host {host.tolist()}"
    monkeypatch.delenv("VCLA_HOST_SAMPLER", raising=False)
    assert not torch.equal(m.generate(do_sample=False, eos_token_id=None, repetition_penalty=1.3, no_repeat_ngram_size=2, **kw), plain)
    assert launches0 > 0


def test_generate_default_chat_config_on_device():
    """The reference's DEFAULT_GENERATION_CONFIG runs as graph replays with the sampler inside; seeded runs repeat, the tokens
    respect the no-repeat-ngram constraint, and Mirostat / TFS (host path) still work."""
    from visualcla.modeling_utils import DEFAULT_GENERATION_CONFIG
    m, ids, px = _model()
    gc = DEFAULT_GENERATION_CONFIG
    kw = dict(input_ids=ids, pixel_values=px, generation_config=gc, max_new_tokens=48, eos_token_id=None, pad_token_id=0)
    torch.manual_seed(1)
    a = m.generate(**kw)
    torch.manual_seed(1)
    b = m.generate(**kw)
    torch.manual_seed(2)
    c = m.generate(**kw)
    assert a.shape == (3, 48) and torch.equal(a, b) and not torch.equal(a, c)
    small = m.generate(input_ids=ids, pixel_values=px, do_sample=True, top_k=3, temperature=2.0, no_repeat_ngram_size=2, max_new_tokens=60,
                       eos_token_id=None, pad_token_id=0)
    for row in small.tolist():
        bigrams = list(zip(row, row[1:]))
        assert len(bigrams) == len(set(bigrams)), "no bigram may repeat with no_repeat_ngram_size=2"
    mir = m.generate(input_ids=ids[:1], pixel_values=px[:1], do_sample=True, temperature=0.8, top_k=40, top_p=0.9, mirostat_mode=2, mirostat_tau=5,
                     mirostat_eta=0.1, max_new_tokens=6, eos_token_id=None, pad_token_id=0)
    tfs = m.generate(input_ids=ids, pixel_values=px, do_sample=True, temperature=0.8, top_k=40, tfs=0.9, max_new_tokens=6, eos_token_id=None, pad_token_id=0)
    assert mir.shape == (1, 6) and tfs.shape == (3, 6)
