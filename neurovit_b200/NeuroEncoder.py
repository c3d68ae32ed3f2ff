"""Drop-in for the reference's src/models/NeuroEncoder.py: NeuroEncoder (3D / 4D switch, frozen-ViT loading,
Grad-CAM hooks), ViT3DEncoder (layout adapter + hard-coded dims), TemporalTransformer, ProjectionHead —
same constructors (one flat config dict), forward signatures, attribute tree and state_dict keys, with the
arithmetic on the sm_100a C-ABI library.

Config keys read (same as the reference, NeuroEncoder.py:19-25,85-87,136-137,174-179): DEVICE, TRAINING_DIM,
GLOBAL_BASE_PATH, BEST_MODEL_PATH, TRAINING_DROPOUT, TRAINING_VIT_INPUT_SIZE, GRADCAM_CUBE_SIZE,
TRAINING_VIT_PATCH_SIZE, DATASET_NAME, GRADCAM_THRESHOLD, GRADCAM_SLICE_DIM, GRADCAM_SLICE_IDX.
Optional new key: PRECISION ("bf16" | "fp32"), default = reference-equivalent mixed precision ("bf16").
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn
from .vit_3d import ViT, _has_hooks


class NeuroEncoder(nn.Module):
    """3D (one volume -> 2 logits) or 4D (T volumes -> frozen ViT3D per timepoint -> temporal head) encoder."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.device = config['DEVICE']

        self.volume_encoder = ViT3DEncoder(config)

        if config['TRAINING_DIM'] == 4:
            # frozen ViT3D from a 3D checkpoint: keep only volume_encoder.vit3d.* keys (NeuroEncoder.py:25-36)
            ckpt = os.path.join(config['GLOBAL_BASE_PATH'], config['BEST_MODEL_PATH'])
            full = torch.load(ckpt, map_location='cpu')
            prefix = "volume_encoder.vit3d."
            vit_sd = {"vit3d." + k[len(prefix):]: v for k, v in full.items() if k.startswith(prefix)}
            self.volume_encoder.load_state_dict(vit_sd, strict=True)
            for p in self.volume_encoder.parameters():
                p.requires_grad = False
            self.volume_encoder.eval()

            self.temporal_transformer = TemporalTransformer(config)
            self.projection_head = ProjectionHead(config)

        self.to(self.device)
        if 'PRECISION' in config:
            self.volume_encoder.vit3d.set_precision(config['PRECISION'])

        # Grad-CAM capture slots, filled by the hooks below
        self.gradients = {}
        self.activations = {}
        self.register_hooks()

    def forward(self, fmri):
        if self.config['TRAINING_DIM'] == 3:
            return self.volume_encoder(fmri)                                   # [B, num_classes]
        if self.config['TRAINING_DIM'] == 4:
            # [B,H,W,D,T] -> permute(0, 4, 1, 2, 3) -> reshape(B*T, H, W, D) (NeuroEncoder.py:54-56) as one
            # de-interleave kernel; config['INPUT_ZSCORE'] (optional, not a reference key) folds the dataset's
            # per-sample z-score (DatasetADNI_4D.py:84-86) into the same pass for raw, un-normalised sequences
            B, T = fmri.shape[0], fmri.shape[4]
            volumes = Fn.fmri_to_volumes(fmri, zscore=self.config.get('INPUT_ZSCORE', False))
            enc = self.volume_encoder(volumes).reshape(B, T, -1)               # [B, T, 2]
            tt, ph = self.temporal_transformer, self.projection_head
            if _has_hooks(tt) or _has_hooks(ph) or _has_hooks(ph.projection_head):
                seq = tt(enc)
                return ph(seq.mean(dim=1))
            layer = tt.transformer.layers[0]
            drop = temporal_dropout(layer)
            seed = Fn.draw_seed() if any(p > 0 for p in drop) else 0
            return Fn.TemporalHeadFn.apply(enc, layer.norm1.eps, drop, seed,
                                           *temporal_param_list(layer, ph.projection_head))
        raise ValueError(f"TRAINING_DIM must be 3 or 4, got {self.config['TRAINING_DIM']!r}")

    def register_hooks(self):
        """Capture the output of the last block's attention LayerNorm and the gradient flowing into it
        (NeuroEncoder.py:70-82). config['GRADCAM_CAPTURE'] (optional, not a reference key):
          "host"   (default) move both to the host inside the hook, exactly as the reference does — one device sync
                   and a [B, N, 1024] fp32 D2H copy per forward and per backward;
          "device" keep the detached tensors on the GPU (get_attention_map moves the small result instead);
          "off"    register no hooks: the last block then runs the fully fused path."""
        mode = self.config.get('GRADCAM_CAPTURE', 'host')
        if mode not in ('host', 'device', 'off'):
            raise ValueError(f"GRADCAM_CAPTURE must be 'host', 'device' or 'off', got {mode!r}")
        self.forward_handle = self.backward_handle = None
        if mode == 'off':
            return
        target = self.volume_encoder.vit3d.transformer.layers[-1][0].norm
        keep = (lambda t: t.detach().cpu()) if mode == 'host' else (lambda t: t.detach())

        def forward_hook(module, inputs, output):
            self.activations = keep(output)

        def backward_hook(module, grad_input, grad_output):
            self.gradients = keep(grad_output[0])

        self.forward_handle = target.register_forward_hook(forward_hook)
        self.backward_handle = target.register_full_backward_hook(backward_hook)

    def get_attention_map(self, x):
        """Grad-CAM over the patch tokens of the last attention LayerNorm (NeuroEncoder.py:84-133)."""
        grid = self.config['TRAINING_VIT_INPUT_SIZE']
        patch = self.config['TRAINING_VIT_PATCH_SIZE']
        threshold = self.config['GRADCAM_THRESHOLD']

        output = self.forward(x)
        class_idx = output.argmax(dim=1)
        one_hot = torch.zeros_like(output)
        one_hot[torch.arange(output.size(0)), class_idx] = 1
        output.backward(gradient=one_hot, retain_graph=True)

        grads, acts = self.gradients, self.activations
        if not torch.is_tensor(grads) or not torch.is_tensor(acts):
            raise RuntimeError("Grad-CAM needs the capture hooks (config GRADCAM_CAPTURE 'host' or 'device')")
        weights = grads.mean(dim=2, keepdim=True)          # importance = mean gradient over the feature axis
        cam = (weights * acts).sum(dim=2)[:, 1:].cpu()     # weighted activation per token, cls dropped
        side = grid // patch
        cam = F.relu(cam.reshape(1, side, side, side))
        cam = (cam - cam.min()) / (cam.max() - cam.min() + 1e-8)
        cut = np.percentile(cam, 100 - threshold)
        kept = torch.from_numpy(np.where(cam >= cut, cam, 0)).unsqueeze(0)
        cam_3d = F.interpolate(kept, size=(grid, grid, grid), mode='trilinear', align_corners=False).squeeze()
        return cam_3d, class_idx

    def visualize_slice(self, cam_3d, original_volume):
        dim = self.config['GRADCAM_SLICE_DIM']
        idx = self.config['GRADCAM_SLICE_IDX']
        if cam_3d is None:
            print("Error: No CAM computed")
            return
        original = original_volume.squeeze().detach().cpu().numpy()
        if original.ndim != 3 or cam_3d.ndim != 3:
            print(f"Shape mismatch: original {original.shape}, CAM {cam_3d.shape}")
            return
        if dim not in (0, 1, 2):
            print(f"Invalid slice dimension: {dim}")
            return
        sel = [slice(None)] * 3
        sel[dim] = idx
        return original[tuple(sel)], cam_3d[tuple(sel)]


class ViT3DEncoder(nn.Module):
    """[B, H, W, D] volumes -> ViT3D logits; model dims hard-coded as in NeuroEncoder.py:181-195."""

    def __init__(self, config):
        super().__init__()
        self.device = config['DEVICE']
        self.dropout = config['TRAINING_DROPOUT']
        self.grid_size = config['TRAINING_VIT_INPUT_SIZE']
        self.cube_size = config['GRADCAM_CUBE_SIZE']
        self.patch_size = config['TRAINING_VIT_PATCH_SIZE']
        number_classes = (self.grid_size // self.cube_size) ** 3 if config['DATASET_NAME'] == 'gradcam' else 2

        self.vit3d = ViT(
            channels=1,
            image_size=self.grid_size,
            image_patch_size=self.patch_size,
            frames=self.grid_size,
            frame_patch_size=self.patch_size,
            num_classes=number_classes,
            dim=1024,
            depth=6,
            heads=8,
            mlp_dim=2048,
            dropout=self.dropout,
            emb_dropout=self.dropout,
            pool='cls',
        ).to(self.device)

    def forward(self, x):
        # [B, H, W, D] -> zero-copy view [B, 1, D, H, W]; the patch-gather kernel reads it through its strides
        volume = x.to(self.device).permute(0, 3, 1, 2).unsqueeze(1)
        return self.vit3d(volume)


def temporal_dropout(layer: nn.TransformerEncoderLayer):
    """(p_attn, p_dropout1, p_ffn, p_dropout2) of the layer's four nn.Dropout sites as active right now: torch's
    default p = 0.1 each in training mode (NeuroEncoder.py:211 passes no dropout argument), all 0 in eval."""
    if not layer.training:
        return (0.0, 0.0, 0.0, 0.0)
    on = lambda m: float(m.p) if m.training else 0.0
    return (float(layer.self_attn.dropout), on(layer.dropout1), on(layer.dropout), on(layer.dropout2))


def temporal_param_list(layer: nn.TransformerEncoderLayer, head: nn.Linear):
    """Parameter tensors in the order of functional.TEMPORAL_KEYS."""
    sa = layer.self_attn
    return [sa.in_proj_weight, sa.in_proj_bias, sa.out_proj.weight, sa.out_proj.bias, layer.linear1.weight,
            layer.linear1.bias, layer.linear2.weight, layer.linear2.bias, layer.norm1.weight, layer.norm1.bias,
            layer.norm2.weight, layer.norm2.bias, head.weight, head.bias]


class TemporalTransformer(nn.Module):
    """One post-norm nn.TransformerEncoderLayer(d_model=2, nhead=2) over the per-timepoint embeddings
    (NeuroEncoder.py:207-217). The nn modules hold the parameters (state_dict keys unchanged); forward runs
    the fused CUDA kernel."""

    def __init__(self, config):
        super().__init__()
        self.device = config['DEVICE']
        encoder_layer = nn.TransformerEncoderLayer(d_model=2, nhead=2, batch_first=True)
        self.transformer = nn.TransformerEncoder(encoder_layer, num_layers=1).to(self.device)

    def forward(self, x):
        layer = self.transformer.layers[0]
        drop = temporal_dropout(layer)
        seed = Fn.draw_seed() if any(p > 0 for p in drop) else 0
        return Fn.TemporalSeqFn.apply(x, layer.norm1.eps, drop, seed, *temporal_param_list(layer, None)[:12])


class ProjectionHead(nn.Module):
    """Linear(2, 2) on the time-averaged embedding (NeuroEncoder.py:219-230)."""

    def __init__(self, config):
        super().__init__()
        self.device = config['DEVICE']
        self.projection_head = nn.Linear(2, 2).to(self.device)

    def forward(self, x):
        return Fn.SmallLinearFn.apply(x, self.projection_head.weight, self.projection_head.bias)
