"""Shared by the CPU and GPU tests of fixture g9 (V18 training retrieval with gradients)."""
import os

import numpy as np

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _g9_setup(device="cpu"):
    import torch

    from oracle.ref_embedding import RefBERTEmbedding

    g = np.load(os.path.join(G, "g9_v18_train_grad.npz"))
    L, D, k = int(g["L"]), int(g["D"]), int(g["k"])
    emb = RefBERTEmbedding(int(g["vocab_size"]), D, dropout=0.0, use_af=True)
    state = {n[len("state/"):]: torch.from_numpy(g[n]) for n in g.files if n.startswith("state/")}
    emb.load_state_dict(state, strict=True)
    emb.to(device).train()
    W = 2
    ref_tokens = [g[f"ref_tokens_{w}"] for w in range(W)]
    ref_af = [g[f"ref_af_{w}"] for w in range(W)]
    masks = [g[f"mask_{w}"] for w in range(W)]
    batch = {"hap_1": torch.from_numpy(g["hap_1"]).to(device), "hap_2": torch.from_numpy(g["hap_2"]).to(device),
             "af": torch.from_numpy(g["af"]).to(device), "window_idx": g["window_idx"].tolist()}
    return g, emb, ref_tokens, ref_af, masks, batch, k


def _g9_check(g, emb, out1, out2, device="cpu", rtol=2e-5):
    import torch

    np.testing.assert_allclose(out1.detach().cpu().numpy(), g["rag_emb_h1"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out2.detach().cpu().numpy(), g["rag_emb_h2"], rtol=0, atol=2e-6)
    loss = (out1 * torch.from_numpy(g["W1"]).to(device)).sum() + (out2 * torch.from_numpy(g["W2"]).to(device)).sum()
    assert abs(float(loss) - float(g["loss"])) <= 1e-3
    emb.zero_grad()
    loss.backward()
    for name, prm in emb.named_parameters():
        want = g["grad/" + name]
        got = (prm.grad if prm.grad is not None else torch.zeros_like(prm)).cpu().numpy()
        scale = max(1e-6, float(np.abs(want).max()))
        assert np.abs(got - want).max() <= rtol * scale + 1e-6, f"gradient of {name}: max err {np.abs(got - want).max()} vs scale {scale}"
