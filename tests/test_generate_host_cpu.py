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
