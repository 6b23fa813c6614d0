"""Host-side logic of VisualCLAModel.generate()/forward() on the CPU, with a fake engine standing in for the native context
(a deterministic toy "model": next token = f(previous token, step)).  Covers what needs no GPU: EOS cut + padding semantics
(HF: finished rows emit pad, output is cut where the last row finished), only-new-tokens return, left-pad validation,
placeholder-layout checks (ref: modeling_visualcla.py:359-367), batch chunking, sampling-knob plumbing."""
import types

import pytest
import torch

import visualcla
from visualcla import _native as N
from visualcla.modeling_visualcla import VisualCLAModel

V, NQ = 50, 4


class FakeEngine:
    device = torch.device("cpu")
    vocab, nq, max_batch, max_seq = V, NQ, 4, 64

    def __init__(self):
        self.prefill_calls = []

    def vision_encode(self, px, return_embeds=False):
        self.px_sum = px.float().sum(dim=(1, 2, 3))

    def _logits(self, tok):
        lg = torch.zeros(tok.shape[0], V)
        lg[torch.arange(tok.shape[0]), (tok.long() * 7 + 3) % V] = 5.0
        return lg

    def prefill(self, ids, mode, rows, all_logits=False, last_logits=True, left_pad=None, pos_from_mask=True):
        self.prefill_calls.append(dict(B=ids.shape[0], mode=mode, rows=rows, left_pad=left_pad, pos_from_mask=pos_from_mask))
        first = (ids[:, -1] % V).to(torch.int32)
        self.hist = [first.clone()]
        S = ids.shape[1] + (NQ if mode == N.IMAGE_AT_HEAD else 0)
        la = torch.zeros(ids.shape[0], S, V) if all_logits else None
        ll = torch.zeros(ids.shape[0], V)
        ll[torch.arange(ids.shape[0]), first.long()] = 5.0
        return (ll if last_logits else None), first, la

    def decode_step(self, tok_in, tok_out, logits=None, use_graph=True):
        lg = self._logits(tok_in)
        if logits is not None:
            logits.copy_(lg)
        tok_out.copy_(lg.argmax(-1).to(torch.int32))
        self.hist.append(tok_out.clone())

    def decode_many(self, tok, n):
        for _ in range(n):
            self.decode_step(tok, tok)

    def read_history(self, B, n):
        return torch.stack(self.hist[:n], 0)


def make_model(image_at_head=False):
    m = object.__new__(VisualCLAModel)
    m._engine = FakeEngine()
    m._tok_buf = {}
    m.image_at_head = image_at_head
    m.tokenizer = types.SimpleNamespace(img_start_token_id=40, img_end_token_id=41, img_token_id=42)
    return m


def chain(first, n):
    out = [int(first)]
    for _ in range(n - 1):
        out.append((out[-1] * 7 + 3) % V)
    return out


def test_generate_returns_only_new_tokens_fixed_length():
    m = make_model()
    ids = torch.tensor([[1, 5, 9], [1, 6, 11]])
    out = m.generate(input_ids=ids, pixel_values=None, do_sample=False, max_new_tokens=6, eos_token_id=None, pad_token_id=0)
    assert out.shape == (2, 6) and out.dtype == torch.int64
    assert out[0].tolist() == chain(9, 6) and out[1].tolist() == chain(11, 6)
    slow = m.generate(input_ids=ids, pixel_values=None, do_sample=False, max_new_tokens=6, eos_token_id=None, pad_token_id=0,
                      output_logits=True, return_dict_in_generate=True)
    assert torch.equal(slow.sequences, out) and len(slow.logits) == 6


def test_eos_cut_and_pad_like_hf():
    m = make_model()
    ids = torch.tensor([[1, 5, 9], [1, 6, 11]])
    full = m.generate(input_ids=ids, do_sample=False, max_new_tokens=12, eos_token_id=None, pad_token_id=0)
    eos = int(full[0, 2])                       # row 0 finishes at step 2
    assert eos not in full[1, :3].tolist()
    out = m.generate(input_ids=ids, do_sample=False, max_new_tokens=12, eos_token_id=eos, pad_token_id=49)
    r0 = out[0].tolist()
    assert r0[:3] == full[0, :3].tolist() and all(t == 49 for t in r0[3:])
    # row 1 keeps generating until it hits eos itself or max_new; the output is cut where the LAST row finished
    if eos in full[1].tolist():
        k = full[1].tolist().index(eos)
        assert out.shape[1] == max(3, k + 1)
    else:
        assert out.shape[1] == 12 and out[1].tolist() == full[1].tolist()


def test_left_pad_validation_and_plumbing():
    m = make_model()
    ids = torch.tensor([[48, 48, 1, 5, 9], [1, 2, 3, 6, 11]])
    mask = torch.tensor([[0, 0, 1, 1, 1], [1, 1, 1, 1, 1]])
    m.generate(input_ids=ids, attention_mask=mask, do_sample=False, max_new_tokens=2, eos_token_id=None, pad_token_id=0)
    call = m._engine.prefill_calls[-1]
    assert call["left_pad"].tolist() == [2, 0] and call["pos_from_mask"] is True
    with pytest.raises(NotImplementedError):      # right padding / holes are not the HF batching convention
        m.generate(input_ids=ids, attention_mask=torch.tensor([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1]]), max_new_tokens=2)
    with pytest.raises(ValueError):
        m.generate(input_ids=ids, attention_mask=torch.tensor([[0, 0, 0, 0, 0], [1, 1, 1, 1, 1]]), max_new_tokens=2)
    m2 = make_model(image_at_head=True)
    with pytest.raises(NotImplementedError):      # the reference's at-head splice ignores the mask
        m2.generate(input_ids=ids, pixel_values=torch.zeros(2, 3, 2, 2), attention_mask=mask, max_new_tokens=2)


def test_placeholder_layout_checks():
    m = make_model()
    px = torch.zeros(1, 3, 2, 2)
    good = torch.tensor([[1, 40, 42, 42, 42, 42, 41, 7, 8]])
    m.generate(input_ids=good, pixel_values=px, do_sample=False, max_new_tokens=1, eos_token_id=None, pad_token_id=0)
    call = m._engine.prefill_calls[-1]
    assert call["mode"] == N.IMAGE_PLACEHOLDER and call["rows"].tolist() == [2]
    bad = torch.tensor([[1, 40, 42, 42, 42, 41, 7, 8, 9]])     # 3 placeholders for 4 query tokens
    with pytest.raises(ValueError, match="Num of patch"):
        m.generate(input_ids=bad, pixel_values=px, max_new_tokens=1)
    noimg = torch.tensor([[1, 5, 6, 7, 8, 9, 10, 11, 12]])      # no <img>: sample carries no image (ref :363-365)
    m.generate(input_ids=noimg, pixel_values=px, do_sample=False, max_new_tokens=1, eos_token_id=None, pad_token_id=0)
    assert m._engine.prefill_calls[-1]["rows"].tolist() == [-1]
    m.tokenizer = None
    with pytest.raises(AttributeError):
        m.generate(input_ids=good, pixel_values=px, max_new_tokens=1)


def test_batch_larger_than_engine_capacity_is_chunked():
    m = make_model()
    ids = torch.arange(1, 1 + 6 * 3).reshape(6, 3)             # 6 requests, engine capacity 4
    out = m.generate(input_ids=ids, do_sample=False, max_new_tokens=4, eos_token_id=None, pad_token_id=0)
    assert out.shape == (6, 4)
    assert [c["B"] for c in m._engine.prefill_calls] == [4, 2]
    for b in range(6):
        assert out[b].tolist() == chain(int(ids[b, -1]) % V, 4)


def test_capacity_and_unsupported_arguments():
    m = make_model()
    ids = torch.tensor([[1, 5, 9]])
    with pytest.raises(ValueError, match="max_seq"):
        m.generate(input_ids=ids, max_new_tokens=100)
    with pytest.raises(NotImplementedError):
        m.generate(input_ids=ids, num_beams=4, max_new_tokens=2)
    with pytest.raises(NotImplementedError):
        m.generate(input_ids=ids, prefix_allowed_tokens_fn=lambda *a: [1], max_new_tokens=2)


def test_sampling_knobs_and_stopping_criteria():
    m = make_model()
    ids = torch.tensor([[1, 5, 9]])
    greedy = m.generate(input_ids=ids, do_sample=False, max_new_tokens=5, eos_token_id=None, pad_token_id=0)
    torch.manual_seed(0)
    samp = m.generate(input_ids=ids, do_sample=True, top_k=1, temperature=0.7, top_p=0.9, max_new_tokens=5, eos_token_id=None, pad_token_id=0)
    assert torch.equal(samp, greedy)              # top_k = 1 collapses sampling onto the argmax
    seen = []

    def stop_after_3(input_ids, scores):
        seen.append(input_ids.shape[1])
        return input_ids.shape[1] >= 3

    out = m.generate(input_ids=ids, do_sample=False, max_new_tokens=9, eos_token_id=None, pad_token_id=0, stopping_criteria=[stop_after_3])
    assert out.shape[1] == 3 and seen == [1, 2, 3]
    # repetition penalty changes the logits path but must keep shapes / dtypes
    rp = m.generate(input_ids=ids, do_sample=False, repetition_penalty=1.3, max_new_tokens=5, eos_token_id=None, pad_token_id=0)
    assert rp.shape == (1, 5)


class SamplerEngine(FakeEngine):
    """FakeEngine + the device-sampler surface: records the spec the host hands to the native side and emulates the in-graph
    loop (EOS flags, pad after EOS) so that generate()'s device path can be followed on the CPU."""

    def __init__(self):
        super().__init__()
        self.spec, self.set_calls, self.fin = None, [], None

    def sampler_supported(self):
        return True

    @staticmethod
    def sampler_spec(**kw):
        return dict(kw)

    def set_sampler(self, spec):
        self.set_calls.append(spec)
        self.spec = spec

    def prefill(self, ids, mode, rows, all_logits=False, last_logits=True, left_pad=None, pos_from_mask=True):
        out = super().prefill(ids, mode, rows, all_logits, last_logits, left_pad, pos_from_mask)
        self.fin = torch.zeros(ids.shape[0], dtype=torch.bool)
        self._mark(self.hist[0])
        return out

    def _mark(self, tok):
        for e in (self.spec or {}).get("eos_token_id", ()):
            self.fin |= tok.long() == e

    def decode_step(self, tok_in, tok_out, logits=None, use_graph=True):
        was = self.fin.clone()
        super().decode_step(tok_in, tok_out, logits, use_graph)
        if self.spec is not None:
            tok_out[was] = self.spec["pad_token_id"]
            self.hist[-1] = tok_out.clone()
            self._mark(torch.where(was, torch.full_like(tok_out, -1), tok_out))

    def read_finished(self, B):
        return self.fin.to(torch.int32)


def make_sampler_model():
    m = make_model()
    m._engine = SamplerEngine()
    return m


def test_device_sampler_is_chosen_exactly_when_the_config_allows(monkeypatch):
    """Host-side decision of VisualCLAModel._device_sampler_spec: the reference's DEFAULT_GENERATION_CONFIG and greedy+EOS /
    greedy+penalties run on the device (one native spec, no per-step host work); everything the fused kernel does not cover keeps
    the host logits-processor path."""
    from visualcla.modeling_utils import DEFAULT_GENERATION_CONFIG
    ids = torch.tensor([[1, 5, 9], [1, 6, 11]])
    m = make_sampler_model()
    torch.manual_seed(3)
    m.generate(input_ids=ids, generation_config=DEFAULT_GENERATION_CONFIG, max_new_tokens=4, eos_token_id=None, pad_token_id=0)
    spec = m._engine.set_calls[0]
    assert m._engine.set_calls[-1] is None, "the sampler is switched off again after the call"
    assert (spec["do_sample"], spec["top_k"], spec["no_repeat_ngram_size"], spec["eos_token_id"]) == (True, 40, 15, [])
    assert abs(spec["temperature"] - 0.5) < 1e-9 and abs(spec["top_p"] - 0.9) < 1e-9 and abs(spec["repetition_penalty"] - 1.1) < 1e-9
    torch.manual_seed(3)
    m2 = make_sampler_model()
    m2.generate(input_ids=ids, generation_config=DEFAULT_GENERATION_CONFIG, max_new_tokens=4, eos_token_id=None, pad_token_id=0)
    assert m2._engine.set_calls[0]["seed"] == spec["seed"], "torch.manual_seed reproduces the Philox seed"

    def used_device(**kw):
        mm = make_sampler_model()
        mm.generate(input_ids=ids[:1], max_new_tokens=4, pad_token_id=0, **kw)     # one row: the reference's Mirostat only handles batch 1
        return len(mm._engine.set_calls) > 0

    assert not used_device(do_sample=False, eos_token_id=None)                                   # plain greedy: the argmax graphs
    assert used_device(do_sample=False, eos_token_id=7)                                          # greedy + EOS: sticky flags on the device
    assert used_device(do_sample=False, eos_token_id=None, repetition_penalty=1.2)
    assert used_device(do_sample=True, top_k=5, eos_token_id=None)
    assert not used_device(do_sample=True, top_k=0, top_p=0.9, eos_token_id=None)               # top-k disabled: host path
    assert not used_device(do_sample=True, top_k=5, tfs=0.9, eos_token_id=None)
    assert not used_device(do_sample=True, top_k=5, mirostat_mode=2, eos_token_id=None)
    assert not used_device(do_sample=True, top_k=5, eos_token_id=None, output_logits=True, return_dict_in_generate=True)
    assert not used_device(do_sample=True, top_k=5, eos_token_id=None, stopping_criteria=[lambda i, s: False])
    assert not used_device(do_sample=False, eos_token_id=[1, 2, 3, 4, 5])                       # more EOS ids than the native spec holds
    monkeypatch.setenv("VCLA_HOST_SAMPLER", "1")
    assert not used_device(do_sample=True, top_k=5, eos_token_id=None)


def test_device_eos_path_matches_host_path():
    """Greedy + EOS through the device path (graphs of 8 steps, finished flags polled between them) returns exactly what the
    per-step host path returns: pad after a row's EOS, cut where the last row finished."""
    ids = torch.tensor([[1, 5, 9], [1, 6, 11]])
    full = make_model().generate(input_ids=ids, do_sample=False, max_new_tokens=20, eos_token_id=None, pad_token_id=0)
    for eos in (int(full[0, 2]), int(full[1, 10]), [int(full[0, 1]), int(full[1, 4])]):
        host = make_model().generate(input_ids=ids, do_sample=False, max_new_tokens=20, eos_token_id=eos, pad_token_id=49)
        dev_m = make_sampler_model()
        dev = dev_m.generate(input_ids=ids, do_sample=False, max_new_tokens=20, eos_token_id=eos, pad_token_id=49)
        assert dev_m._engine.set_calls and dev.shape == host.shape and torch.equal(dev, host), (eos, dev.tolist(), host.tolist())


def test_mirostat_wiring_replays_the_reference_quirk():
    """ref modeling_utils.py:366-371 removes the non-temperature warpers while iterating over them, which skips every second one:
    with temperature, top-k and top-p configured, top-p survives next to Mirostat."""
    from transformers import GenerationConfig
    from transformers.generation import logits_process as lp
    from visualcla import modeling_utils as mu
    gc = GenerationConfig(do_sample=True, temperature=0.7, top_k=40, top_p=0.9, repetition_penalty=1.1)
    gc.mirostat_mode, gc.mirostat_tau, gc.mirostat_eta = 2, 5.0, 0.1
    procs = VisualCLAModel._build_processors(gc, None)
    kinds = [type(p) for p in procs]
    assert kinds == [lp.RepetitionPenaltyLogitsProcessor, lp.TemperatureLogitsWarper, lp.TopPLogitsWarper, mu.MirostatLogitsWarper]
    gc.mirostat_mode = 0
    assert mu.MirostatLogitsWarper not in [type(p) for p in VisualCLAModel._build_processors(gc, None)]
