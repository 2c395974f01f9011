"""Compact-transformer building blocks (CCT / CVT / ViT-Lite and their text variants).

Capability parity with the vendored Compact-Transformers tree of the reference
(/root/reference/src/blades/models/cifar10/cctnets/: cct.py, cvt.py, vit.py, text/*,
utils/{transformers,tokenizer,embedder,stochastic_depth,helpers}.py).  Re-designed as
ONE parametrised implementation: a single attention / encoder-layer / classifier that
takes an optional key-padding mask, instead of parallel masked/unmasked copies, and a
spec-table driven factory registry instead of ~60 hand-written factory functions.
Module / parameter names are kept (``tokenizer.conv_layers.N.0.weight``,
``classifier.blocks.N.self_attn.qkv.weight`` ...) so state-dicts interchange.

Attention uses ``F.scaled_dot_product_attention`` (fused kernel on B200) when no
attention-dropout is active, and the explicit softmax path otherwise.
"""
from __future__ import annotations

import math
from typing import Callable, Dict

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import Dropout, Identity, LayerNorm, Linear, Module, ModuleList, Parameter, init

__all__ = [
    "drop_path", "DropPath", "Attention", "MaskedAttention", "TransformerEncoderLayer",
    "MaskedTransformerEncoderLayer", "TransformerClassifier", "MaskedTransformerClassifier",
    "Tokenizer", "TextTokenizer", "Embedder", "resize_pos_embed", "pe_check", "fc_check",
    "register_model", "MODEL_REGISTRY",
]

MODEL_REGISTRY: Dict[str, Callable] = {}


def register_model(fn):
    MODEL_REGISTRY[fn.__name__] = fn
    return fn


# ------------------------------------------------------------------------------ stochastic depth
def drop_path(x, drop_prob: float = 0., training: bool = False):
    """Per-sample stochastic depth: zero the whole residual branch of a sample w.p. ``drop_prob``."""
    if not training or drop_prob == 0.:
        return x
    keep = 1.0 - drop_prob
    mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
    return x * mask / keep


class DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training)


# ------------------------------------------------------------------------------ attention / encoder
class MaskedAttention(Module):
    """Multi-head self attention with an optional ``[B, N]`` boolean validity mask."""

    def __init__(self, dim, num_heads=8, attention_dropout=0.1, projection_dropout=0.1):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = Linear(dim, dim * 3, bias=False)
        self.attn_drop = Dropout(attention_dropout)
        self.proj = Linear(dim, dim)
        self.proj_drop = Dropout(projection_dropout)

    def forward(self, x, mask=None):
        B, N, C = x.shape
        q, k, v = self.qkv(x).view(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        pair = None
        if mask is not None:
            assert mask.shape[-1] == N, 'mask has incorrect dimensions'
            pair = (mask[:, None, :] & mask[:, :, None])[:, None]          # [B,1,N,N]
        p = self.attn_drop.p if self.training else 0.0
        if p == 0.0 and pair is None:
            y = F.scaled_dot_product_attention(q, k, v, scale=self.scale)
        else:
            att = (q @ k.transpose(-2, -1)) * self.scale
            if pair is not None:
                att = att.masked_fill(~pair, -torch.finfo(att.dtype).max)
            y = self.attn_drop(att.softmax(dim=-1)) @ v
        return self.proj_drop(self.proj(y.transpose(1, 2).reshape(B, N, C)))


class Attention(MaskedAttention):
    """Unmasked attention (same parameters; ``forward(x)``)."""

    def forward(self, x, mask=None):  # noqa: D401
        return super().forward(x, None)


class MaskedTransformerEncoderLayer(Module):
    """Pre-norm attention + post-norm MLP block of Compact Transformers."""
    _attn_cls = MaskedAttention

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1,
                 attention_dropout=0.1, drop_path_rate=0.1):
        super().__init__()
        self.pre_norm = LayerNorm(d_model)
        self.self_attn = self._attn_cls(dim=d_model, num_heads=nhead,
                                        attention_dropout=attention_dropout, projection_dropout=dropout)
        self.linear1 = Linear(d_model, dim_feedforward)
        self.dropout1 = Dropout(dropout)
        self.norm1 = LayerNorm(d_model)
        self.linear2 = Linear(dim_feedforward, d_model)
        self.dropout2 = Dropout(dropout)
        self.drop_path = DropPath(drop_path_rate) if drop_path_rate > 0 else Identity()
        self.activation = F.gelu

    def forward(self, src, mask=None, *args, **kwargs):
        src = src + self.drop_path(self.self_attn(self.pre_norm(src), mask))
        src = self.norm1(src)
        ff = self.linear2(self.dropout1(self.activation(self.linear1(src))))
        return src + self.drop_path(self.dropout2(ff))


class TransformerEncoderLayer(MaskedTransformerEncoderLayer):
    _attn_cls = Attention

    def forward(self, src, *args, **kwargs):
        return super().forward(src, None)


def _sinusoid(n_pos: int, dim: int, padding_idx: bool = False) -> torch.Tensor:
    pos = torch.arange(n_pos, dtype=torch.float32)[:, None]
    i = torch.arange(dim, dtype=torch.float32)[None, :]
    ang = pos / torch.pow(torch.tensor(10000.0), 2 * torch.div(i, 2, rounding_mode='floor') / dim)
    pe = torch.where((torch.arange(dim) % 2 == 0)[None, :], torch.sin(ang), torch.cos(ang))[None]
    if padding_idx:
        pe = torch.cat([torch.zeros(1, 1, dim), pe], dim=1)
    return pe


class _ClassifierBase(Module):
    _layer_cls = MaskedTransformerEncoderLayer
    _masked = True

    def _build(self, seq_pool, embedding_dim, num_layers, num_heads, mlp_ratio, num_classes, dropout,
               attention_dropout, stochastic_depth, positional_embedding, seq_len):
        if positional_embedding not in ('sine', 'learnable', 'none'):
            positional_embedding = 'sine'
        self.embedding_dim = embedding_dim
        self.seq_pool = seq_pool
        self.num_tokens = 0
        assert seq_len is not None or positional_embedding == 'none', \
            f"Positional embedding is set to {positional_embedding} and the sequence length was not specified."
        if not seq_pool:
            seq_len += 1
            self.class_emb = Parameter(torch.zeros(1, 1, embedding_dim), requires_grad=True)
            self.num_tokens = 1
        else:
            self.attention_pool = Linear(embedding_dim, 1)
        if positional_embedding == 'learnable':
            if self._masked:
                seq_len += 1                      # slot for the padding index
            self.positional_emb = Parameter(torch.zeros(1, seq_len, embedding_dim), requires_grad=True)
            init.trunc_normal_(self.positional_emb, std=0.2)
        elif positional_embedding == 'sine':
            self.positional_emb = Parameter(_sinusoid(seq_len, embedding_dim, padding_idx=self._masked),
                                            requires_grad=False)
        else:
            self.positional_emb = None
        self._total_len = seq_len
        self.dropout = Dropout(p=dropout)
        rates = torch.linspace(0, stochastic_depth, num_layers).tolist()
        self.blocks = ModuleList([
            self._layer_cls(d_model=embedding_dim, nhead=num_heads,
                            dim_feedforward=int(embedding_dim * mlp_ratio), dropout=dropout,
                            attention_dropout=attention_dropout, drop_path_rate=rates[i])
            for i in range(num_layers)])
        self.norm = LayerNorm(embedding_dim)
        self.fc = Linear(embedding_dim, num_classes)
        self.apply(self.init_weight)

    def _run(self, x, mask=None):
        if self.positional_emb is None and self._total_len is not None and x.size(1) < self._total_len:
            x = F.pad(x, (0, 0, 0, self._total_len - x.size(1)), mode='constant', value=0)
        if not self.seq_pool:
            x = torch.cat((self.class_emb.expand(x.shape[0], -1, -1), x), dim=1)
            if mask is not None:
                mask = torch.cat([mask.new_ones((mask.shape[0], 1), dtype=torch.bool), mask.bool()], dim=1)
        if self.positional_emb is not None:
            x = x + self.positional_emb[:, :x.size(1)] if self.positional_emb.size(1) != x.size(1) \
                else x + self.positional_emb
        x = self.dropout(x)
        for blk in self.blocks:
            x = blk(x, mask=mask) if self._masked else blk(x)
        x = self.norm(x)
        if self.seq_pool:
            x = (F.softmax(self.attention_pool(x), dim=1).transpose(-1, -2) @ x).squeeze(-2)
        else:
            x = x[:, 0]
        return self.fc(x)

    @staticmethod
    def init_weight(m):
        if isinstance(m, Linear):
            init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                init.constant_(m.bias, 0)
        elif isinstance(m, LayerNorm):
            init.constant_(m.bias, 0)
            init.constant_(m.weight, 1.0)

    @staticmethod
    def sinusoidal_embedding(n_channels, dim, padding_idx=False):
        return _sinusoid(n_channels, dim, padding_idx)


class TransformerClassifier(_ClassifierBase):
    _layer_cls = TransformerEncoderLayer
    _masked = False

    def __init__(self, seq_pool=True, embedding_dim=768, num_layers=12, num_heads=12, mlp_ratio=4.0,
                 num_classes=1000, dropout=0.1, attention_dropout=0.1, stochastic_depth=0.1,
                 positional_embedding='learnable', sequence_length=None):
        super().__init__()
        self.sequence_length = sequence_length
        self._build(seq_pool, embedding_dim, num_layers, num_heads, mlp_ratio, num_classes, dropout,
                    attention_dropout, stochastic_depth, positional_embedding, sequence_length)

    def forward(self, x):
        return self._run(x)


class MaskedTransformerClassifier(_ClassifierBase):
    def __init__(self, seq_pool=True, embedding_dim=768, num_layers=12, num_heads=12, mlp_ratio=4.0,
                 num_classes=1000, dropout=0.1, attention_dropout=0.1, stochastic_depth=0.1,
                 positional_embedding='sine', seq_len=None, *args, **kwargs):
        super().__init__()
        self.seq_len = seq_len
        self._build(seq_pool, embedding_dim, num_layers, num_heads, mlp_ratio, num_classes, dropout,
                    attention_dropout, stochastic_depth, positional_embedding, seq_len)

    def forward(self, x, mask=None):
        return self._run(x, mask)


# ------------------------------------------------------------------------------ tokenizers / embedder
class Tokenizer(nn.Module):
    """Conv(+act)(+maxpool) stack that turns an image into a token sequence ``[B, L, C]``."""

    def __init__(self, kernel_size, stride, padding, pooling_kernel_size=3, pooling_stride=2,
                 pooling_padding=1, n_conv_layers=1, n_input_channels=3, n_output_channels=64,
                 in_planes=64, activation=None, max_pool=True, conv_bias=False):
        super().__init__()
        chans = [n_input_channels] + [in_planes] * (n_conv_layers - 1) + [n_output_channels]
        stages = []
        for cin, cout in zip(chans[:-1], chans[1:]):
            stages.append(nn.Sequential(
                nn.Conv2d(cin, cout, kernel_size=(kernel_size, kernel_size), stride=(stride, stride),
                          padding=(padding, padding), bias=conv_bias),
                nn.Identity() if activation is None else activation(),
                nn.MaxPool2d(kernel_size=pooling_kernel_size, stride=pooling_stride,
                             padding=pooling_padding) if max_pool else nn.Identity()))
        self.conv_layers = nn.Sequential(*stages)
        self.flattener = nn.Flatten(2, 3)
        self.apply(self.init_weight)

    def sequence_length(self, n_channels=3, height=224, width=224):
        with torch.no_grad():
            return self.forward(torch.zeros((1, n_channels, height, width))).shape[1]

    def forward(self, x):
        return self.flattener(self.conv_layers(x)).transpose(-2, -1)

    @staticmethod
    def init_weight(m):
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight)


class TextTokenizer(nn.Module):
    """1-D (over sequence) convolutional tokenizer for word embeddings, mask-aware."""

    def __init__(self, kernel_size, stride, padding, pooling_kernel_size=3, pooling_stride=2,
                 pooling_padding=1, embedding_dim=300, n_output_channels=128, activation=None,
                 max_pool=True, *args, **kwargs):
        super().__init__()
        self.max_pool = max_pool
        self.conv_layers = nn.Sequential(
            nn.Conv2d(1, n_output_channels, kernel_size=(kernel_size, embedding_dim), stride=(stride, 1),
                      padding=(padding, 0), bias=False),
            nn.Identity() if activation is None else activation(),
            nn.MaxPool2d(kernel_size=(pooling_kernel_size, 1), stride=(pooling_stride, 1),
                         padding=(pooling_padding, 0)) if max_pool else nn.Identity())
        self.apply(self.init_weight)

    def seq_len(self, seq_len=32, embed_dim=300):
        with torch.no_grad():
            return self.forward(torch.zeros((1, seq_len, embed_dim)))[0].shape[1]

    def forward_mask(self, mask):
        conv, pool = self.conv_layers[0], self.conv_layers[2]
        m = mask.unsqueeze(1).float()
        ones = torch.ones((1, 1, conv.kernel_size[0]), device=mask.device)
        m = F.conv1d(m, ones, None, conv.stride[0], conv.padding[0], 1, 1)
        if self.max_pool:
            m = F.max_pool1d(m, pool.kernel_size[0], pool.stride[0], pool.padding[0], 1, False, False)
        return m.squeeze(1) > 0

    def forward(self, x, mask=None):
        x = self.conv_layers(x.unsqueeze(1)).transpose(1, 3).squeeze(1)
        if mask is not None:
            mask = self.forward_mask(mask).unsqueeze(-1).float()
            x = x * mask
        return x, mask

    @staticmethod
    def init_weight(m):
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight)


class Embedder(nn.Module):
    def __init__(self, word_embedding_dim=300, vocab_size=100000, padding_idx=1, pretrained_weight=None,
                 embed_freeze=False, *args, **kwargs):
        super().__init__()
        if pretrained_weight is not None:
            self.embeddings = nn.Embedding.from_pretrained(pretrained_weight, freeze=embed_freeze)
        else:
            self.embeddings = nn.Embedding(vocab_size, word_embedding_dim, padding_idx=padding_idx)
        self.embeddings.weight.requires_grad = not embed_freeze

    def forward_mask(self, mask):
        return mask.view(mask.shape[0], mask.shape[1], 1).sum(-1) > 0

    def forward(self, x, mask=None):
        e = self.embeddings(x)
        if mask is not None:
            e = e * self.forward_mask(mask).unsqueeze(-1).float()
        return e, mask

    @staticmethod
    def init_weight(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        else:
            nn.init.normal_(m.weight)


# ------------------------------------------------------------------------------ checkpoint helpers
def resize_pos_embed(posemb, posemb_new, num_tokens=1):
    """Bilinear-resize a learnable positional grid to a new token count (keeps class tokens)."""
    n_new = posemb_new.shape[1] - num_tokens
    tok, grid = posemb[:, :num_tokens], posemb[0, num_tokens:]
    g_old, g_new = int(math.sqrt(len(grid))), int(math.sqrt(n_new))
    grid = grid.reshape(1, g_old, g_old, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=(g_new, g_new), mode='bilinear')
    grid = grid.permute(0, 2, 3, 1).reshape(1, g_new * g_new, -1)
    return torch.cat([tok, grid], dim=1)


def pe_check(model, state_dict, pe_key='classifier.positional_emb'):
    own = model.state_dict()
    if pe_key in state_dict and pe_key in own and own[pe_key].shape != state_dict[pe_key].shape:
        state_dict[pe_key] = resize_pos_embed(state_dict[pe_key], own[pe_key],
                                              num_tokens=model.classifier.num_tokens)
    return state_dict


def fc_check(model, state_dict, fc_key='classifier.fc'):
    own = model.state_dict()
    for key in (f'{fc_key}.weight', f'{fc_key}.bias'):
        if key in state_dict and key in own and own[key].shape != state_dict[key].shape:
            state_dict[key] = own[key]          # class count changed: keep fresh head
    return state_dict
