"""Paged KV cache, capacity limits and per-context state of the native engine (through the Python API -> C ABI)."""
import ctypes as C

import pytest
import torch

import visualcla_oracle as O

pytestmark = pytest.mark.gpu


def _cfg():
    return O.PathConfig(v_layers=1, r_layers=1, t_hidden=512, t_heads=4, t_ffn=1408, t_layers=2, t_vocab=2003)


def _gen(m, ids, px, n):
    return m.generate(input_ids=ids.cuda(), pixel_values=px.cuda(), do_sample=False, max_new_tokens=n, eos_token_id=None, pad_token_id=0)


def test_pages_are_allocated_on_demand_and_not_contiguous():
    """Sequences get physical pages round-robin as they grow: a sequence's pages are interleaved with the others', the decode
    kernels gather through the page table, and a permuted hand-out order changes the table but not one token."""
    import visualcla
    cfg = _cfg()
    B, T, n_new = 3, 100, 90                      # S = 164 -> 3 pages at prefill, crosses into a 4th page during decode
    m = visualcla.VisualCLAModel.from_synthetic(cfg.to_dict(), seed=21, max_batch=4, max_seq=320)
    eng = m._engine
    pps, total, pt = eng.kv_geometry()
    assert (pps, total, pt) == (5, 20, 64)
    px, ids = O.make_inputs(cfg, B, T, seed=4)
    base = _gen(m, ids, px, n_new)
    table, owned, free, exhausted = eng.kv_pages()
    S_end = T + cfg.r_queries + n_new - 1          # tokens cached after the last decode step
    need = (S_end + 1 + pt - 1) // pt              # + the page reserved for the next token
    assert exhausted == 0 and owned[:B].tolist() == [need] * B and owned[B:].tolist() == [0]
    assert free == total - B * need
    rows = [table[b, :need].tolist() for b in range(B)]
    assert sorted(sum(rows, [])) == list(range(B * need)), "every page handed out exactly once, lowest pages first"
    assert rows[0] == [0, 3, 6, 9], f"round-robin hand-out interleaves the sequences: {rows}"
    # permute the hand-out order: different physical placement, identical tokens
    eng.kv_debug_shuffle(7)
    again = _gen(m, ids, px, n_new)
    table2, owned2, _, _ = eng.kv_pages()
    assert torch.equal(again, base)
    assert owned2.tolist() == owned.tolist() and not torch.equal(table2[:B, :need], table[:B, :need])
    # a smaller batch afterwards starts from a clean pool
    one = _gen(m, ids[:1], px[:1], n_new)
    assert torch.equal(one[0], base[0])
    assert eng.kv_pages()[1].tolist() == [need, 0, 0, 0]


def test_decode_refuses_to_run_past_the_context_capacity():
    import visualcla
    from visualcla import _native as N
    from visualcla.dp import generate_dp
    cfg = O.tiny_config()
    m = visualcla.VisualCLAModel.from_synthetic(cfg.to_dict(), seed=0, max_batch=2, max_seq=40)
    eng = m._engine
    px, ids = O.make_inputs(cfg, 2, 12, seed=3)          # S = 12 + 8 = 20
    with pytest.raises(ValueError):
        _gen(m, ids, px, 21)
    with pytest.raises(ValueError):
        generate_dp(m, ids, px, 21)
    full = _gen(m, ids, px, 20)                            # exactly fills the context
    assert full.shape == (2, 20)
    tok = torch.zeros(2, dtype=torch.int32, device="cuda")
    eng.decode_step(tok, tok, None)                        # feeding the 20th token fills the last slot (index 39)
    with pytest.raises(N.NativeError, match="capacity"):
        eng.decode_step(tok, tok, None)                    # one more step would index past the sequence's pages
    with pytest.raises(N.NativeError, match="capacity"):
        eng.decode_many(tok, 4)
    eng.reset()
    with pytest.raises(N.NativeError, match="prefill"):
        eng.decode_step(tok, tok, None)


def test_weight_shape_is_checked_on_both_sides_of_the_abi():
    import visualcla
    from visualcla import _native as N
    cfg = O.tiny_config()
    m = visualcla.VisualCLAModel.from_synthetic(cfg.to_dict(), seed=0, max_batch=1, max_seq=32)
    eng = m._engine
    name = "text_model.lm_head.weight"
    good = eng.read_weight(name)
    with pytest.raises(ValueError, match="shape mismatch"):
        eng.load_weight(name, good[:-4])                   # e.g. a 49954-row head against a 49958 config
    with pytest.raises(ValueError, match="shape mismatch"):
        eng.load_weight(name, good.t().contiguous())
    with pytest.raises(KeyError):
        eng.load_weight("text_model.no_such.weight", good)
    short = good[:-4].contiguous()
    rc = eng.lib.vcla_load_weight(eng._ctx, name.encode(), N.ptr(short), N.VCLA_BF16, short.numel(), 0, None)
    assert rc != 0 and b"elements" in eng.lib.vcla_last_error()
    eng.load_weight(name, good)                            # the right shape still loads


def test_contexts_do_not_share_rope_tables_or_scratch():
    """A second context with another rope_theta / max_seq must not disturb the first (its captured decode graphs keep
    pointing at its own tables)."""
    import visualcla
    cfg = O.tiny_config()
    a = visualcla.VisualCLAModel.from_synthetic(cfg.to_dict(), seed=5, max_batch=2, max_seq=64)
    px, ids = O.make_inputs(cfg, 2, 10, seed=8)
    before = _gen(a, ids, px, 12)
    other = dict(cfg.to_dict(), rope_theta=500.0)
    b = visualcla.VisualCLAModel.from_synthetic(other, seed=5, max_batch=2, max_seq=512)
    diff = _gen(b, ids, px, 12)
    after = _gen(a, ids, px, 12)
    assert torch.equal(after, before)
    w = O.make_weights(cfg, 5)
    ocfg = O.PathConfig(**other)
    o_tok, o_log = O.generate_greedy(w, ocfg, ids, px, 12, image_at_head=True)
    top2 = o_log.topk(2, -1).values
    decisive = (top2[..., 0] - top2[..., 1]) > 0.05 * o_log.abs().max()
    assert bool(((diff.cpu() == o_tok) | ~decisive)[:, :1].all()), "context b follows ITS theta"
