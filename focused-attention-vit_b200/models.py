"""Callers of the hot path with the reference's constructors, attribute names and state_dict keys:
`TransformerBlock`, `VisionTransformerMHLA` (/root/reference/models/vit_mhla.py:20-267) and `SPPPViTMHLA`
(/root/reference/models/sppp_mhla.py:113-333).  They are the end-to-end harness of the images/sec metric.

Differences from the reference, all on the SPPP front end and none numerical:
  * the per-image Python loop `map_patches` + `pool` + `torch.stack` (sppp_mhla.py:283-300) is one batched
    `assign_batch` + `pool_batch` call;
  * superpixel centroids (sppp_mhla.py:226-262, a B x K Python loop with a device sync per superpixel) are one
    batched segment-mean on the device;
  * `forward(x, segmentation_maps=None)`: label maps are an input of the path; when they are not given,
    `self.segmentation.segment(x)` is called exactly like the reference does (SLIC via scikit-image, if installed).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import fused_block
from .mhla import MultiHeadLatentAttention, compute_dtype, fold_latent
from .sppp import PatchToSuperpixelMapper, SuperpixelPooling


class PatchEmbedding(nn.Module):
    """models/vit.py:19-53.  `projection` = Sequential(rearrange, Linear) so that the weight keeps the reference's
    state_dict key `projection.1.weight`."""

    class _Patchify(nn.Module):
        def __init__(self, patch_size: int):
            super().__init__()
            self.patch_size = patch_size

        def forward(self, x):  # 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)'
            B, C, H, W = x.shape
            p = self.patch_size
            if H % p or W % p:
                raise RuntimeError(f"image size {H}x{W} is not divisible by patch_size {p}")
            return x.reshape(B, C, H // p, p, W // p, p).permute(0, 2, 4, 3, 5, 1).reshape(B, (H // p) * (W // p),
                                                                                         p * p * C)

    def __init__(self, img_size=224, patch_size=16, in_channels=3, embed_dim=768):
        super().__init__()
        self.img_size = img_size
        self.patch_size = patch_size
        self.num_patches = (img_size // patch_size) ** 2
        self.projection = nn.Sequential(self._Patchify(patch_size),
                                        nn.Linear(patch_size * patch_size * in_channels, embed_dim))

    def forward(self, x):
        p = self.patch_size
        if (x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and not x.requires_grad and x.shape[2] == x.shape[3]
                and x.shape[2] % p == 0 and p % 4 == 0):
            # rearrangement + operand cast in one favit pass, projection on the favit GEMM (fp32 output like nn.Linear
            # without autocast, the compute dtype under autocast: what the reference's module returns)
            from . import ops
            cd = compute_dtype(x)
            if cd in (torch.float32, torch.bfloat16):
                lin = self.projection[1]
                with torch.autocast("cuda", enabled=False):
                    return ops.linear(ops.patchify(x, p, cd), lin.weight, lin.bias)
        return self.projection(x)


def favit_linear(lin: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    """nn.Linear through the favit GEMM (the classification head): same dtypes as nn.Linear with / without autocast."""
    if x.is_cuda and x.dtype in (torch.float32, torch.bfloat16):
        from . import ops
        cd = compute_dtype(x)
        if cd in (torch.float32, torch.bfloat16) and lin.in_features % 8 == 0 and lin.out_features % 8 == 0:
            with torch.autocast("cuda", enabled=False):
                return ops.linear(x if x.dtype == cd else x.to(cd), lin.weight, lin.bias)
    return lin(x)


class MLP(nn.Module):
    """models/vit.py:107-139."""

    def __init__(self, in_features, hidden_features, out_features, dropout=0.0):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        return self.dropout(self.fc2(self.dropout(self.act(self.fc1(x)))))


class TransformerBlock(nn.Module):
    """models/vit_mhla.py:20-109 (identical copy at models/sppp_mhla.py:21-110)."""

    def __init__(self, embed_dim: int, num_heads: int, mlp_ratio: float = 4.0, dropout: float = 0.0,
                 attn_dropout: float = 0.0, window_size: int = 7, use_mhla: bool = False):
        super().__init__()
        self.norm1 = nn.LayerNorm(embed_dim)
        if use_mhla:
            self.attn = MultiHeadLatentAttention(embed_dim=embed_dim, num_heads=num_heads, window_size=window_size,
                                                 dropout=attn_dropout)
        else:
            self.attn = nn.MultiheadAttention(embed_dim=embed_dim, num_heads=num_heads, dropout=attn_dropout,
                                              batch_first=True)
        self.norm2 = nn.LayerNorm(embed_dim)
        self.mlp = MLP(in_features=embed_dim, hidden_features=int(embed_dim * mlp_ratio), out_features=embed_dim,
                       dropout=dropout)
        self.use_mhla = use_mhla

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.use_mhla:
            cd = compute_dtype(x)
            if fused_block.fusable(x, self.attn, self.mlp.dropout.p, self.training, attention_mask, cd,
                                   self.mlp.fc1.out_features):
                # whole block as two favit ops (LN, GEMMs with fused bias/GELU/residual, window attention)
                with torch.autocast("cuda", enabled=False):
                    return fused_block.fused_block(x, self.norm1, self.attn, self.norm2, self.mlp.fc1, self.mlp.fc2, cd,
                                                   self.mlp.dropout.p if self.training else 0.0)
        x_norm = self.norm1(x)
        if self.use_mhla:
            attn_output = self.attn(x_norm, attention_mask)
        else:
            attn_output, _ = self.attn(query=x_norm, key=x_norm, value=x_norm,
                                       key_padding_mask=None if attention_mask is None else ~attention_mask)
        x = x + attn_output
        x = x + self.mlp(self.norm2(x))
        return x


def run_blocks(blocks, x: torch.Tensor) -> torch.Tensor:
    """The block stack of both models.  When every block takes the fused path the latent fold of all of them is
    batched (fused_block.run_blocks); otherwise blocks run one by one."""
    blks = list(blocks)
    cd = compute_dtype(x)
    a0 = blks[0].attn if blks and getattr(blks[0], "use_mhla", False) else None
    if a0 is not None and len(blks) > 1 and all(
            isinstance(b, TransformerBlock) and b.use_mhla and b.attn.num_heads == a0.num_heads and
            b.attn.window_size == a0.window_size and b.attn.embed_dim == a0.embed_dim and
            b.mlp.dropout.p == blks[0].mlp.dropout.p and b.training == blks[0].training and
            fused_block.fusable(x, b.attn, b.mlp.dropout.p, b.training, None, cd, b.mlp.fc1.out_features)
            for b in blks):
        with torch.autocast("cuda", enabled=False):
            return fused_block.run_blocks(blks, x, cd, blks[0].mlp.dropout.p if blks[0].training else 0.0)
    for block in blks:
        x = block(x)
    return x


def _init_weights_recursive(m):
    if isinstance(m, nn.Linear):
        nn.init.normal_(m.weight, std=0.02)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, nn.LayerNorm):
        nn.init.ones_(m.weight)
        nn.init.zeros_(m.bias)


class VisionTransformerMHLA(nn.Module):
    """models/vit_mhla.py:112-267."""

    def __init__(self, img_size: int = 224, patch_size: int = 4, in_channels: int = 3, num_classes: int = 1000,
                 embed_dim: int = 768, depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0,
                 dropout: float = 0.0, attn_dropout: float = 0.0, embed_dropout: float = 0.0, window_size: int = 7,
                 use_mhla: bool = False):
        super().__init__()
        self.img_size = img_size
        self.patch_size = patch_size
        self.in_channels = in_channels
        self.num_classes = num_classes
        self.embed_dim = embed_dim
        self.depth = depth
        self.num_heads = num_heads
        self.use_mhla = use_mhla
        self.patch_embed = PatchEmbedding(img_size=img_size, patch_size=patch_size, in_channels=in_channels,
                                          embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(embed_dropout)
        self.blocks = nn.ModuleList([
            TransformerBlock(embed_dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, dropout=dropout,
                             attn_dropout=attn_dropout, window_size=window_size, use_mhla=use_mhla)
            for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)
        self._init_weights()

    def _init_weights(self):
        nn.init.normal_(self.cls_token, std=0.02)
        nn.init.normal_(self.pos_embed, std=0.02)
        self.apply(_init_weights_recursive)

    def forward_features(self, x: torch.Tensor) -> torch.Tensor:
        batch_size = x.shape[0]
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(batch_size, -1, -1), x), dim=1)
        x = run_blocks(self.blocks, self.pos_drop(x + self.pos_embed))
        # LayerNorm is per token and only the class token is used (vit_mhla.py:241-244): normalise that row alone
        return self.norm(x[:, 0])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return favit_linear(self.head, self.forward_features(x))

    def get_num_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)


class SuperpixelSegmentation:
    """models/sppp.py:26-74.  The reference runs SLIC through scikit-image on the CPU, image by image (device -> host
    copy, Python loop, host -> device copy).  CUDA images are segmented on the device by `favit::slic_segment` for the
    whole batch at once (same algorithm family; no Lab conversion / connectivity pass, see include/favit.h); CPU images
    go through scikit-image exactly like the reference, if it is installed."""

    def __init__(self, num_segments: int = 16, compactness: float = 0.1, sigma: float = 1.0, max_num_iter: int = 10):
        self.num_segments = num_segments
        self.compactness = compactness
        self.sigma = sigma
        self.max_num_iter = max_num_iter     # skimage's default

    def segment(self, image: torch.Tensor) -> torch.Tensor:
        if image.is_cuda:
            from . import ops
            batch_mode = image.dim() == 4
            imgs = image if batch_mode else image.unsqueeze(0)
            maps = ops.slic_segment(imgs.float(), self.num_segments, self.compactness, self.sigma, self.max_num_iter)
            return maps if batch_mode else maps[0]
        try:
            from skimage.segmentation import slic
        except ImportError as e:  # not installed in this image
            raise RuntimeError("SuperpixelSegmentation.segment needs scikit-image (SLIC); pass segmentation_maps to "
                               "forward() or replace model.segmentation.segment") from e
        batch_mode = image.dim() == 4
        imgs = image if batch_mode else image.unsqueeze(0)
        maps = [torch.from_numpy(slic(im.permute(1, 2, 0).cpu().numpy(), n_segments=self.num_segments,
                                      compactness=self.compactness, sigma=self.sigma, start_label=0)).to(image.device)
                for im in imgs]
        return torch.stack(maps) if batch_mode else maps[0]


class DynamicPositionalEncoding(nn.Module):
    """models/sppp.py:226-300."""

    def __init__(self, embed_dim: int, dropout: float = 0.0):
        super().__init__()
        self.embed_dim = embed_dim
        self.dropout = nn.Dropout(dropout)

    def forward(self, x: torch.Tensor, superpixel_centroids: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, N, _ = x.shape
        dev = x.device
        D = self.embed_dim
        if superpixel_centroids is None:
            position = torch.arange(N, dtype=torch.float, device=dev).unsqueeze(1)
            div_term = torch.exp(torch.arange(0, D, 2, dtype=torch.float, device=dev) * (-math.log(10000.0) / D))
            pe = torch.zeros(N, D, device=dev)
            pe[:, 0::2] = torch.sin(position * div_term)
            pe[:, 1::2] = torch.cos(position * div_term)
            pe = pe.unsqueeze(0).expand(B, -1, -1)
        else:
            c = superpixel_centroids
            if c.shape[1] < N:
                c = torch.cat([torch.full((B, 1, 2), 0.5, device=dev), c], dim=1)
            freq = torch.exp(torch.arange(0, D // 2, dtype=torch.float, device=dev) * (-math.log(10000.0) / (D // 2)))
            pe = torch.cat([torch.sin(c[:, :, 0:1] * freq), torch.cos(c[:, :, 1:2] * freq)], dim=-1)
        return self.dropout(x + pe)


class SPPPViTMHLA(nn.Module):
    """models/sppp_mhla.py:113-333."""

    def __init__(self, img_size: int = 224, patch_size: int = 4, in_channels: int = 3, num_classes: int = 1000,
                 embed_dim: int = 768, depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0,
                 dropout: float = 0.0, attn_dropout: float = 0.0, embed_dropout: float = 0.0,
                 num_superpixels: int = 16, compactness: float = 0.1, pooling_type: str = 'mean',
                 window_size: int = 7, use_mhla: bool = False):
        super().__init__()
        self.img_size = img_size
        self.patch_size = patch_size
        self.in_channels = in_channels
        self.num_classes = num_classes
        self.embed_dim = embed_dim
        self.depth = depth
        self.num_heads = num_heads
        self.num_superpixels = num_superpixels
        self.use_mhla = use_mhla
        self.segmentation = SuperpixelSegmentation(num_segments=num_superpixels, compactness=compactness)
        self.patch_mapper = PatchToSuperpixelMapper(patch_size=patch_size)
        self.pooling = SuperpixelPooling(pooling_type=pooling_type)
        self.patch_embed = PatchEmbedding(img_size=img_size, patch_size=patch_size, in_channels=in_channels,
                                          embed_dim=embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = DynamicPositionalEncoding(embed_dim, embed_dropout)
        self.blocks = nn.ModuleList([
            TransformerBlock(embed_dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, dropout=dropout,
                             attn_dropout=attn_dropout, window_size=window_size, use_mhla=use_mhla)
            for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)
        self.validate_slots = True   # one D2H read per forward; bench.py checks its synthetic maps once up front
        # 'mean' pooling commutes with the linear patch embedding: pool the raw pixel patches of every superpixel, then
        # project R rows per image instead of P.  False = the reference's two steps (embed, then pool).
        self.fuse_patch_pool = True
        self._init_weights()

    def _init_weights(self):
        nn.init.normal_(self.cls_token, std=0.02)
        self.apply(_init_weights_recursive)

    def _calculate_superpixel_centroids(self, segmentation_maps: torch.Tensor) -> torch.Tensor:
        """[B,H,W] -> [B,K,2] (x, y) centroids of labels 0..K-1 in normalised coordinates; (0.5, 0.5) for labels that
        do not occur (sppp_mhla.py:226-262), as one batched segment mean."""
        seg = segmentation_maps
        B, H, W = seg.shape
        K = self.num_superpixels
        dev = seg.device
        if seg.is_cuda and K <= 4096:
            from . import ops
            return ops.sppp_centroids(seg if seg.dtype == torch.int64 else seg.to(torch.int64), K)
        valid = (seg >= 0) & (seg < K)
        idx = (torch.arange(B, device=dev).view(B, 1, 1) * K + seg.clamp(0, K - 1)).reshape(-1)
        w = valid.reshape(-1).float()
        ys = (torch.arange(H, device=dev).float() / H).view(1, H, 1).expand(B, H, W).reshape(-1)
        xs = (torch.arange(W, device=dev).float() / W).view(1, 1, W).expand(B, H, W).reshape(-1)
        n = torch.zeros(B * K, device=dev).index_add_(0, idx, w)
        sx = torch.zeros(B * K, device=dev).index_add_(0, idx, xs * w)
        sy = torch.zeros(B * K, device=dev).index_add_(0, idx, ys * w)
        has = n > 0
        cx = torch.where(has, sx / n.clamp_min(1), torch.full_like(sx, 0.5))
        cy = torch.where(has, sy / n.clamp_min(1), torch.full_like(sy, 0.5))
        return torch.stack([cx, cy], dim=-1).view(B, K, 2)

    def _pooled_patch_embeddings(self, x: torch.Tensor, assignment) -> torch.Tensor:
        """patch_embed + pool('mean') of sppp_mhla.py:281-300 as segment-mean of the raw pixel patches (one pass over the
        image, favit_sppp_pool_pixels) followed by the projection of the R pooled rows (favit GEMM; its autograd gives
        the Linear's weight / bias gradients).  fp32 out, like the reference's pool (sppp.py:198)."""
        from . import ops
        K = self.num_superpixels
        if self.validate_slots:
            ns = assignment.num_slots
            lo, hi = int(ns.min()), int(ns.max())
            if lo != hi or lo != K:
                raise RuntimeError(f"stack expects each tensor to be equal size: images have between {lo} and {hi} "
                                   f"superpixel slots, expected {K} (reference sppp_mhla.py:300)")
        cd = compute_dtype(x)
        lin = self.patch_embed.projection[1]
        px = ops.sppp_pool_pixels(x, assignment.order, assignment.offsets, assignment.num_slots, self.patch_size, K, cd)
        with torch.autocast("cuda", enabled=False):
            return ops.linear(px, lin.weight, lin.bias).float()

    def forward(self, x: torch.Tensor, segmentation_maps: Optional[torch.Tensor] = None) -> torch.Tensor:
        batch_size = x.shape[0]
        if segmentation_maps is None:
            segmentation_maps = self.segmentation.segment(x)
        centroids = None
        if segmentation_maps.is_cuda and segmentation_maps.dim() == 3 and self.num_superpixels <= 4096:
            # dominant labels and centroids from one pass over the label maps
            assignment, centroids = self.patch_mapper.assign_batch_with_centroids(
                segmentation_maps, self.img_size, self.num_superpixels, self.num_superpixels)
        else:
            assignment = self.patch_mapper.assign_batch(segmentation_maps, self.img_size, r_cap=self.num_superpixels)
        if (self.fuse_patch_pool and self.pooling.pooling_type == 'mean' and x.is_cuda and x.dtype == torch.float32
                and not x.requires_grad and x.shape[-1] % self.patch_size == 0 and x.shape[-2] % self.patch_size == 0):
            pooled = self._pooled_patch_embeddings(x, assignment)
        else:
            patch_embeddings = self.patch_embed(x)
            pooled = self.pooling.pool_batch(patch_embeddings, assignment, self.num_superpixels,
                                             validate=self.validate_slots)
        if centroids is None:
            centroids = self._calculate_superpixel_centroids(segmentation_maps)
        if (pooled.is_cuda and pooled.dtype == torch.float32 and self.embed_dim % 2 == 0
                and centroids.shape[1] == pooled.shape[1]):
            # class-token concat + dynamic positional encoding (sppp_mhla.py:302-310) in one favit launch
            from . import ops
            x = self.pos_embed.dropout(ops.sppp_embed_tokens(pooled, self.cls_token, centroids))
        else:
            x = torch.cat((self.cls_token.expand(batch_size, -1, -1), pooled), dim=1)
            x = self.pos_embed(x, centroids)
        x = run_blocks(self.blocks, x)
        # LayerNorm is per token and only the class token is used (sppp_mhla.py:317-323): normalise that row alone
        return favit_linear(self.head, self.norm(x[:, 0]))

    def get_num_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)
