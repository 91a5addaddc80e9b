"""Parameter containers that reproduce the reference's module tree.

These classes hold parameters/buffers only: they register the same child names in the
same order as the reference so that (a) ``state_dict()`` has the reference's exact 225
keys and (b) ``torch.manual_seed(s); VQAModel()`` consumes the RNG stream identically and
yields bit-identical random-init weights.  None of them has a
``forward``: all arithmetic runs in the CUDA engine (see ``model.py`` / ``program.py``).

Reference layout followed: models/cnn_backbone.py:101-418, models/attention_modules.py:27-243,391-433,
models/text_encoder.py:33-477, models/cross_attention.py:41-365, models/fusion.py:30-250,
models/vqa_model.py:30-92.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn


class _Holder(nn.Module):
    """A module that only stores parameters; calling it is a bug."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is a parameter container; "
                           "arithmetic runs in the vqa_b200 CUDA engine")


def _conv(cin, cout, k, stride=1, pad=0):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=pad, bias=False)


# ----------------------------------------------------------------------------- backbone
class SEAttention(_Holder):
    def __init__(self, channels: int, reduction: int = 16):
        super().__init__()
        r = max(channels // reduction, 1)
        self.fc1 = nn.Linear(channels, r, bias=False)
        self.fc2 = nn.Linear(r, channels, bias=False)
        self.channels, self.reduced_channels = channels, r


class SpatialAttention(_Holder):
    def __init__(self, kernel_size: int = 7):
        super().__init__()
        if kernel_size % 2 != 1:
            raise AssertionError("Kernel size must be odd")
        self.conv = nn.Conv2d(2, 1, kernel_size=kernel_size, padding=kernel_size // 2, bias=False)


class AttentionWrapper(_Holder):
    def __init__(self, channels, use_se=True, use_spatial=True, se_reduction=16, spatial_kernel=7):
        super().__init__()
        self.use_se, self.use_spatial = use_se, use_spatial
        if use_se:
            self.se = SEAttention(channels, se_reduction)
        if use_spatial:
            self.spatial = SpatialAttention(spatial_kernel)


class ResidualBlock(_Holder):
    def __init__(self, cin, cout, stride=1, downsample: Optional[nn.Module] = None):
        super().__init__()
        self.conv1 = _conv(cin, cout, 3, stride, 1)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = _conv(cout, cout, 3, 1, 1)
        self.bn2 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class ResidualStage(_Holder):
    def __init__(self, cin, cout, num_blocks=2, stride=1, use_se=True, use_spatial=True, se_reduction=16):
        super().__init__()
        shortcut = None
        if stride != 1 or cin != cout:  # shortcut is constructed before the block (RNG order)
            shortcut = nn.Sequential(_conv(cin, cout, 1, stride), nn.BatchNorm2d(cout))
        blocks = [ResidualBlock(cin, cout, stride, shortcut)]
        blocks += [ResidualBlock(cout, cout) for _ in range(1, num_blocks)]
        self.blocks = nn.Sequential(*blocks)
        self.attention = AttentionWrapper(cout, use_se=use_se, use_spatial=use_spatial,
                                          se_reduction=se_reduction)


class CustomResNet(_Holder):
    def __init__(self, in_channels=3, base_channels=64, num_blocks=(2, 2, 2, 2),
                 use_se=True, use_spatial=True, se_reduction=16):
        super().__init__()
        self.use_se, self.use_spatial = use_se, use_spatial
        ch = [base_channels * m for m in (1, 2, 4, 8)]
        self.stem = nn.Sequential(_conv(in_channels, ch[0], 7, 2, 3), nn.BatchNorm2d(ch[0]),
                                  nn.ReLU(inplace=True), nn.MaxPool2d(kernel_size=3, stride=2, padding=1))
        self.stage1 = ResidualStage(ch[0], ch[0], num_blocks[0], 1, use_se, False, se_reduction)
        self.stage2 = ResidualStage(ch[0], ch[1], num_blocks[1], 2, use_se, False, se_reduction)
        self.stage3 = ResidualStage(ch[1], ch[2], num_blocks[2], 2, use_se, use_spatial, se_reduction)
        self.stage4 = ResidualStage(ch[2], ch[3], num_blocks[3], 2, use_se, use_spatial, se_reduction)
        self.output_channels = ch[3]
        self.output_spatial_size = 7
        for m in self.modules():  # same walk order as the reference initialiser
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def get_feature_map_size(self):
        return (self.output_channels, self.output_spatial_size, self.output_spatial_size)


# ----------------------------------------------------------------------------- text encoder
class PositionalEncoding(_Holder):
    def __init__(self, embed_dim: int, max_length: int = 512, dropout: float = 0.1):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        table = torch.zeros(max_length, embed_dim)
        pos = torch.arange(0, max_length, dtype=torch.float).unsqueeze(1)
        freq = torch.exp(torch.arange(0, embed_dim, 2).float() * (-math.log(10000.0) / embed_dim))
        table[:, 0::2] = torch.sin(pos * freq)
        table[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pe", table.unsqueeze(0))


class MultiHeadSelfAttention(_Holder):
    def __init__(self, embed_dim, num_heads, dropout=0.1):
        super().__init__()
        if embed_dim % num_heads:
            raise AssertionError(f"embed_dim ({embed_dim}) must be divisible by num_heads ({num_heads})")
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.head_dim = embed_dim // num_heads
        self.scale = math.sqrt(self.head_dim)
        for name in ("W_q", "W_k", "W_v", "W_o"):
            setattr(self, name, nn.Linear(embed_dim, embed_dim, bias=False))
        self.dropout = nn.Dropout(dropout)


class FeedForwardNetwork(_Holder):
    def __init__(self, embed_dim, hidden_dim, dropout=0.1):
        super().__init__()
        self.fc1 = nn.Linear(embed_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, embed_dim)
        self.dropout = nn.Dropout(dropout)


class TransformerEncoderLayer(_Holder):
    def __init__(self, embed_dim, num_heads, ffn_hidden_dim, dropout=0.1):
        super().__init__()
        self.self_attention = MultiHeadSelfAttention(embed_dim, num_heads, dropout)
        self.norm1 = nn.LayerNorm(embed_dim)
        self.dropout1 = nn.Dropout(dropout)
        self.ffn = FeedForwardNetwork(embed_dim, ffn_hidden_dim, dropout)
        self.norm2 = nn.LayerNorm(embed_dim)
        self.dropout2 = nn.Dropout(dropout)


class TransformerTextEncoder(_Holder):
    def __init__(self, vocab_size, embed_dim=256, num_layers=4, num_heads=8, ffn_hidden_dim=1024,
                 max_length=50, dropout=0.1, pad_idx=0):
        super().__init__()
        self.embed_dim, self.pad_idx = embed_dim, pad_idx
        self.token_embedding = nn.Embedding(vocab_size, embed_dim, padding_idx=pad_idx)
        self.positional_encoding = PositionalEncoding(embed_dim, max_length, dropout)
        self.layers = nn.ModuleList(
            TransformerEncoderLayer(embed_dim, num_heads, ffn_hidden_dim, dropout) for _ in range(num_layers))
        self.final_norm = nn.LayerNorm(embed_dim)
        nn.init.normal_(self.token_embedding.weight, mean=0, std=embed_dim ** -0.5)
        if pad_idx is not None:
            self.token_embedding.weight.data[pad_idx].zero_()


# ----------------------------------------------------------------------------- fusion
class CrossAttention(_Holder):
    def __init__(self, embed_dim, num_heads=8, dropout=0.1, bias=False):
        super().__init__()
        if embed_dim % num_heads:
            raise AssertionError(f"embed_dim ({embed_dim}) must be divisible by num_heads ({num_heads})")
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.head_dim = embed_dim // num_heads
        self.scale = math.sqrt(self.head_dim)
        for name in ("W_q", "W_k", "W_v", "W_o"):
            setattr(self, name, nn.Linear(embed_dim, embed_dim, bias=bias))
        self.dropout = nn.Dropout(dropout)
        for name in ("W_q", "W_k", "W_v", "W_o"):
            lin = getattr(self, name)
            nn.init.xavier_uniform_(lin.weight)
            if lin.bias is not None:
                nn.init.zeros_(lin.bias)


class MultiHeadCrossAttention(_Holder):
    def __init__(self, embed_dim, num_heads=8, dropout=0.1, use_ffn=True, ffn_hidden_dim=None):
        super().__init__()
        self.norm_query = nn.LayerNorm(embed_dim)
        self.norm_kv = nn.LayerNorm(embed_dim)
        self.cross_attention = CrossAttention(embed_dim, num_heads, dropout)
        self.dropout1 = nn.Dropout(dropout)
        self.use_ffn = use_ffn
        if use_ffn:
            hidden = ffn_hidden_dim or 4 * embed_dim
            self.norm_ffn = nn.LayerNorm(embed_dim)
            self.ffn = nn.Sequential(nn.Linear(embed_dim, hidden), nn.ReLU(inplace=True), nn.Dropout(dropout),
                                     nn.Linear(hidden, embed_dim), nn.Dropout(dropout))


class StackedCrossAttention(_Holder):
    def __init__(self, embed_dim, num_heads=8, num_layers=2, dropout=0.1):
        super().__init__()
        self.layers = nn.ModuleList(MultiHeadCrossAttention(embed_dim, num_heads, dropout)
                                    for _ in range(num_layers))


class ImageFeatureProjector(_Holder):
    def __init__(self, in_channels, embed_dim, spatial_size=7, use_position_embed=True, dropout=0.1):
        super().__init__()
        self.in_channels, self.embed_dim, self.spatial_size = in_channels, embed_dim, spatial_size
        self.num_positions = spatial_size * spatial_size
        self.projection = nn.Sequential(nn.Linear(in_channels, embed_dim), nn.LayerNorm(embed_dim),
                                        nn.Dropout(dropout))
        self.use_position_embed = use_position_embed
        if use_position_embed:
            self.position_embedding = nn.Parameter(torch.randn(1, self.num_positions, embed_dim) * 0.02)


class GatingMechanism(_Holder):
    def __init__(self, embed_dim):
        super().__init__()
        self.gate = nn.Sequential(nn.Linear(embed_dim * 2, embed_dim), nn.Sigmoid())


class MultimodalFusion(_Holder):
    def __init__(self, image_channels=512, image_spatial_size=7, embed_dim=256, num_heads=8,
                 num_cross_layers=2, dropout=0.1, use_gating=True):
        super().__init__()
        self.embed_dim, self.use_gating = embed_dim, use_gating
        self.image_projector = ImageFeatureProjector(image_channels, embed_dim, image_spatial_size, True, dropout)
        self.cross_attention = StackedCrossAttention(embed_dim, num_heads, num_cross_layers, dropout)
        if use_gating:
            self.gate = GatingMechanism(embed_dim)
        self.output_norm = nn.LayerNorm(embed_dim)

    def get_attention_visualization(self, attention_weights: list, spatial_size: int = 7) -> torch.Tensor:
        """Layer- and head-averaged cross-attention as [B, L, S, S] (models/fusion.py:338-363)."""
        avg = torch.stack(attention_weights, dim=0).mean(dim=0).mean(dim=1)
        b, lq, _ = avg.shape
        return avg.view(b, lq, spatial_size, spatial_size)


# ----------------------------------------------------------------------------- head
class AnswerHead(_Holder):
    def __init__(self, input_dim, hidden_dim, num_answers, dropout=0.3):
        super().__init__()
        self.classifier = nn.Sequential(
            nn.Linear(input_dim, hidden_dim), nn.ReLU(inplace=True), nn.Dropout(dropout),
            nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(inplace=True), nn.Dropout(dropout),
            nn.Linear(hidden_dim // 2, num_answers))
        for m in self.classifier:
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)
