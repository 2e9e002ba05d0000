"""GPU: the V18 training-loop retriever (rag_snvbert_b200.embedding_rag.EmbeddingRagRetriever) against fixture g9, which
the reference's own EmbeddingRAGDataset.process_batch_retrieval + BERTEmbedding produced (tests/golden/make_golden.py):
outputs rag_emb_h1 / rag_emb_h2 [B, k, L, D] and the gradients of a loss on them w.r.t. every embedding parameter."""
import numpy as np
import pytest

from g9_helpers import _g9_check, _g9_setup

pytestmark = pytest.mark.gpu


def _retriever(g, ref_tokens, ref_af, masks, **kw):
    from rag_snvbert_b200.embedding_rag import EmbeddingRagRetriever

    return EmbeddingRagRetriever(ref_tokens, ref_af, masks, int(g["D"]), mask_index=int(g["mask_index"]), **kw)


@pytest.mark.parametrize("cached", [1, 2])
def test_outputs_and_gradients_match_the_reference(cached):
    g, emb, ref_tokens, ref_af, masks, batch, k = _g9_setup("cuda")
    r = _retriever(g, ref_tokens, ref_af, masks, max_cached_windows=cached)
    out = r.process_batch_retrieval(dict(batch), emb, k)
    assert out["rag_emb_h1"].shape == g["rag_emb_h1"].shape and out["rag_emb_h1"].requires_grad
    _g9_check(g, emb, out["rag_emb_h1"], out["rag_emb_h2"], device="cuda", rtol=1e-4)


def test_cached_call_makes_no_host_synchronisation():
    """with both windows' panels cached, a batch call must not synchronise the host (the reference reads B*k ids with int())"""
    import torch

    g, emb, ref_tokens, ref_af, masks, batch, k = _g9_setup("cuda")
    r = _retriever(g, ref_tokens, ref_af, masks, max_cached_windows=2)
    r.process_batch_retrieval(dict(batch), emb, k)   # builds the two panels (an index build synchronises once, as a JIT rebuild does)
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        out = r.process_batch_retrieval(dict(batch), emb, k)
        loss = out["rag_emb_h1"].sum() + out["rag_emb_h2"].sum()
        loss.backward()
    finally:
        torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()
    np.testing.assert_allclose(out["rag_emb_h1"].detach().cpu().numpy(), g["rag_emb_h1"], rtol=0, atol=2e-6)


def test_mask_refresh_invalidates_the_cached_panel():
    g, emb, ref_tokens, ref_af, masks, batch, k = _g9_setup("cuda")
    r = _retriever(g, ref_tokens, ref_af, masks, max_cached_windows=2)
    r.process_batch_retrieval(dict(batch), emb, k)
    assert sorted(r._cache) == [0, 1]
    r.set_window_mask(0, np.zeros_like(masks[0]))
    assert sorted(r._cache) == [1]
