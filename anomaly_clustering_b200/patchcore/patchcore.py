"""Mirror of PatchMaker / AnomalyClusteringCore (models/patchcore/patchcore.py:276-465).

`_embed` keeps the reference's signature and return convention (list of per-patch numpy rows
when detach=True) but computes LayerNorm + patchify + resize + Preprocessing + Aggregator in ONE
fused CUDA kernel instead of materialising the 9x unfolded tensor; `embed_device` is the
addition that keeps Z on the device."""
from __future__ import annotations

import numpy as np
import torch
import tqdm

from .. import ops
from . import common


class PatchMaker:
    """patchcore.py:434-482."""

    def __init__(self, patchsize, stride=None):
        self.patchsize = patchsize
        self.stride = stride

    def patchify(self, features, return_spatial_info=False):
        """[B,C,H,W] -> [B, h*w, C, patchsize, patchsize] (+ [h, w]); patchcore.py:439-465."""
        unfolded, grid = ops.patchify(features, self.patchsize, self.stride if self.stride else 1)
        if return_spatial_info:
            return unfolded, grid
        return unfolded

    def unpatch_scores(self, x, batchsize):
        return x.reshape(batchsize, -1, *x.shape[1:])

    def score(self, x):
        was_numpy = isinstance(x, np.ndarray)
        if was_numpy:
            x = torch.from_numpy(x)
        while x.ndim > 1:
            x = torch.max(x, dim=-1).values
        return x.numpy() if was_numpy else x


class AnomalyClusteringCore(torch.nn.Module):
    """patchcore.py:276-431.  Detection-only members of the reference (`anomaly_scorer`,
    `anomaly_segmentor`, `featuresampler`) are accepted and ignored: the clustering path never
    calls them."""

    def __init__(self, device):
        super().__init__()
        self.device = device

    def load(self, backbone, layers_to_extract_from, device, input_shape, pretrain_embed_dimension,
             target_embed_dimension, patchsize=3, patchstride=1, anomaly_score_num_nn=1, featuresampler=None,
             nn_method=None, **kwargs):
        self.backbone = backbone.to(device)
        self.layers_to_extract_from = layers_to_extract_from
        self.input_shape = input_shape
        self.device = device
        self.patch_maker = PatchMaker(patchsize, stride=patchstride)
        self.forward_modules = torch.nn.ModuleDict({})
        feature_aggregator = common.NetworkFeatureAggregator(self.backbone, self.layers_to_extract_from, self.device)
        feature_dimensions = feature_aggregator.feature_dimensions(input_shape)
        self.forward_modules["feature_aggregator"] = feature_aggregator
        self.forward_modules["preprocessing"] = common.Preprocessing(feature_dimensions, pretrain_embed_dimension)
        self.pretrain_embed_dimension = pretrain_embed_dimension
        self.target_embed_dimension = target_embed_dimension
        self.forward_modules["preadapt_aggregator"] = common.Aggregator(target_dim=target_embed_dimension)
        self.featuresampler = featuresampler
        return self

    # -- the hooked backbone features (torch; not the accelerated part)
    def _features(self, images):
        _ = self.forward_modules["feature_aggregator"].eval()
        with torch.no_grad():
            feats = self.forward_modules["feature_aggregator"](images)
        return [feats[layer] for layer in self.layers_to_extract_from]

    def embed_device(self, images, operand=None, want_lo=False):
        """Addition: images [B,3,H,W] -> (Z [B*P, D] device fp32, hi, lo, patch grid); B > 1 allowed."""
        feats = [f.float() for f in self._features(images.to(torch.float).to(self.device))]
        return ops.embed(feats, self.patch_maker.patchsize, self.patch_maker.stride or 1, self.pretrain_embed_dimension,
                         self.target_embed_dimension, layernorm=True, operand=operand, want_lo=want_lo)

    def embed(self, data, supervised):
        """patchcore.py:337-353."""
        print("{:-^80}".format("embedding"))
        if isinstance(data, torch.utils.data.DataLoader):
            features, labels = [], []
            with tqdm.tqdm(total=len(data)) as progress:
                for image in data:
                    is_anomaly = None
                    if isinstance(image, dict):
                        is_anomaly = image["is_anomaly"]
                        image = image["image"]
                    with torch.no_grad():
                        input_image = image.to(torch.float).to(self.device)
                        features.append(self._embed(input_image, supervised))
                        labels.append(is_anomaly)
                    progress.update(1)
            return features, labels
        return self._embed(data, supervised)

    def _embed(self, images, supervised, detach=True, provide_patch_shapes=False):
        """patchcore.py:355-431: returns the [B*P, D] embedding (as a list of numpy rows when
        `detach`, like the reference's `_detach`)."""
        Z, _, _, grid = self.embed_device(images)
        patch_shapes = [list(grid)]
        for f in self._last_feature_shapes()[1:]:
            patch_shapes.append(f)
        out = [x for x in Z.detach().cpu().numpy()] if detach else Z
        if provide_patch_shapes:
            return out, patch_shapes
        return out

    def _last_feature_shapes(self):
        outs = self.forward_modules["feature_aggregator"].outputs
        shapes = []
        for layer in self.layers_to_extract_from:
            v = ops.feature_view(outs[layer])
            shapes.append(list(ops.patch_grid(v.shape[2], v.shape[3], self.patch_maker.patchsize, self.patch_maker.stride or 1)))
        return shapes
