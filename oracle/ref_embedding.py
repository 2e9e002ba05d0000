"""TEST INFRASTRUCTURE ONLY (imported by tests/ alone; never by the product path).

torch restatement of the reference's embedding layer, so that tests on the GPU box (which has no /root/reference)
can re-create the exact module whose parameters, outputs and gradients tests/golden/g9 records:

  BERTEmbedding          src/model/embedding/bert.py:11-75     token embedding (padding_idx 0) + sinusoidal position
                                                               + Fourier AF embedding, summed, then dropout
  PositionalEmbedding    src/model/embedding/position.py:9-39  sin / cos table, max_len 1030
  AFEmbedding            src/model/embedding/af_embedding.py:17-92  af * learnable basis -> sin, cos -> Linear, LayerNorm,
                                                               GELU, Linear

Parameter names match the reference's state_dict (tokenizer.weight, af_embedding.basis_freqs,
af_embedding.projection.{0,1,3}.*, position.pe), so a recorded state_dict loads with load_state_dict(strict=True).
Pinned by tests/test_oracle_golden.py against the embeddings inside g4 / g9.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

MAX_SEQ_LEN = 1030


class _Position(nn.Module):
    def __init__(self, dims: int, max_len: int = MAX_SEQ_LEN):
        super().__init__()
        pe = torch.zeros([max_len, dims]).float()
        position = torch.arange(0, max_len).float().unsqueeze(1)
        div_term = (torch.arange(0, dims, 2).float() * -(math.log(10000.0) / dims)).exp()
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))

    def forward(self, x):
        return self.pe[:, : x.size(dim=1)]


class _AF(nn.Module):
    def __init__(self, embed_size: int, num_basis: int = 32):
        super().__init__()
        self.basis_freqs = nn.Parameter(torch.logspace(0, math.log10(100), num_basis))
        self.projection = nn.Sequential(nn.Linear(num_basis * 2, embed_size), nn.LayerNorm(embed_size), nn.GELU(),
                                        nn.Linear(embed_size, embed_size))

    def forward(self, af):
        x = af.unsqueeze(-1) * self.basis_freqs
        feats = torch.cat([torch.sin(2 * math.pi * x), torch.cos(2 * math.pi * x)], dim=-1)
        return self.projection(feats)


class RefBERTEmbedding(nn.Module):
    def __init__(self, vocab_size: int, embed_size: int, dropout: float = 0.1, use_af: bool = True):
        super().__init__()
        self.tokenizer = nn.Embedding(vocab_size, embed_size, padding_idx=0)
        self.position = _Position(embed_size)
        self.use_af = use_af
        if use_af:
            self.af_embedding = _AF(embed_size, 32)
        self.embed_size = embed_size
        self.dropout = nn.Dropout(dropout)

    def forward(self, seq, af=None, pos: bool = False):
        out = self.tokenizer(seq)
        if pos:
            out = out + self.position(seq)
        if self.use_af and af is not None:
            out = out + self.af_embedding(af)
        return self.dropout(out)


def v18_process_batch_retrieval(ref_tokens, ref_af, masks, mask_index, batch, embedding_layer, k):
    """Plain restatement of EmbeddingRAGDataset.process_batch_retrieval (src/dataset/embedding_rag_dataset.py:285-444) on the
    tensors' own device: per window group in order of first appearance - masked reference embedded in eval mode without
    gradient, queries embedded, torch.cdist + topk(largest=False), unique ids, COMPLETE tokens re-encoded with gradient,
    scattered to [B, k, L, D].  (The reference rebuilds its one-window cache on every window switch; the result only
    depends on the weights, so the restatement embeds the window's panel whenever it meets the window.)"""
    h1, h2, af = batch["hap_1"], batch["hap_2"], batch["af"]
    groups = {}
    for i, w in enumerate(batch["window_idx"]):
        groups.setdefault(int(w), []).append(i)
    B, L = h1.shape
    D = embedding_layer.embed_size
    out1 = torch.zeros(B, k, L, D, dtype=torch.float32, device=h1.device)
    out2 = torch.zeros_like(out1)
    for w, idxs in groups.items():
        tok = torch.as_tensor(ref_tokens[w], device=h1.device)
        raf = torch.as_tensor(ref_af[w], device=h1.device)
        msk = torch.as_tensor(masks[w], device=h1.device)
        masked = tok.clone()
        masked[:, msk == 1] = mask_index
        was = embedding_layer.training
        embedding_layer.eval()
        with torch.no_grad():
            ref_emb = embedding_layer(masked, af=raf.unsqueeze(0).expand(tok.shape[0], -1), pos=True)
        embedding_layer.train(was)
        e1 = embedding_layer(h1[idxs], af=af[idxs], pos=True)
        e2 = embedding_layer(h2[idxs], af=af[idxs], pos=True)
        bw = len(idxs)
        flat = ref_emb.reshape(-1, L * D)
        I1 = torch.cdist(e1.reshape(bw, L * D), flat, p=2).topk(k, largest=False, dim=1)[1]
        I2 = torch.cdist(e2.reshape(bw, L * D), flat, p=2).topk(k, largest=False, dim=1)[1]
        uniq = torch.cat([I1.flatten(), I2.flatten()]).unique()
        emb = embedding_layer(tok[uniq], af=raf.unsqueeze(0).expand(uniq.numel(), -1), pos=True)
        pos = {int(o): n for n, o in enumerate(uniq)}
        for i, b in enumerate(idxs):
            for j in range(k):
                out1[b, j] = emb[pos[int(I1[i, j])]]
                out2[b, j] = emb[pos[int(I2[i, j])]]
    return out1, out2
