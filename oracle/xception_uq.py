"""TEST INFRASTRUCTURE ONLY -- CPU restatement of Slideflow's Xception-UQ MC-dropout inference.

PARITY UNPINNED (model half).  The arithmetic of this half of the hot path is NOT in
/root/reference: it lives in un-vendored, un-pinned third-party packages -- `slideflow>=1.1.0rc1`
and `tensorflow>=2.7` (reference requirements.txt:1,5; `tf.keras.applications.Xception`) -- neither
of which is installed here or on the GPU box, and the reference holds no test or golden vector for
it.  This file therefore restates the published algorithm and anchors on the reference's own call
sites:

  * architecture contract: reference biscuit/hp.py:3-24 (xception, 299 px, dropout 0.1, 2 hidden
    layers x 1024, pooling 'avg', include_top False, uq);
  * pre-processing order and return shape: reference results.py:249-258
    (`tf.image.per_image_standardization` -> batch -> `interface(batch)` -> (mean softmax, std));
  * which outputs BISCUIT consumes: reference biscuit/utils.py:19-28 (class-1 mean and std);
  * layer list, TF padding rules, BN eps, dropout scaling, population std: SURVEY.md Appendix B
    (Keras `applications/xception.py`, Slideflow `UncertaintyInterface` -- recalled, frozen here).

Frozen choices (documented in DESIGN.md): BatchNorm eps = 1e-3; Dropout(rate) AFTER each hidden
Dense layer (sites 1 and 2), none after pooling (site 0) unless enabled; keep = u >= rate with
u a 32-bit Philox4x32-10 draw; y = x * keep / (1 - rate); uncertainty = population std (ddof=0) of
the softmax over T samples; T full forward passes in the reference schedule
(`reference_schedule=True`), one backbone pass + T head passes otherwise (bit-identical result here
because the backbone is deterministic in inference mode).

Two numeric tiers:
  fp32        -- every tensor float32 (the reference's arithmetic type);
  bf16-emulated -- activations / GEMM weights rounded to bfloat16 at exactly the points where the
                 CUDA path stores them (fp32 accumulation), to separate quantisation error from
                 kernel bugs.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
TILE_PX = 299

# The layer list and the random-init generator are the ORACLE'S OWN statement (oracle/xception_arch.py), independent of
# biscuit_b200.weights; tests/test_model_oracle_cpu.py checks that the two statements agree and anchors this one on
# Keras' published parameter count.
from oracle.xception_arch import (ENTRY_BLOCKS, FEATURES, MIDDLE_BLOCKS, layer_table,  # noqa: E402,F401
                                  make_weights)


# ----------------------------------------------------------------------------------------
# Philox4x32-10 dropout masks (counter based; the CUDA head kernel draws the same stream)
# ----------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11).  uint32 arrays in, 4 uint32 arrays out."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK32 for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & _MASK32, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & _MASK32, lo0
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def dropout_threshold(rate: float) -> int:
    """keep iff draw >= threshold; threshold = floor(rate * 2^32)"""
    return int(np.floor(float(np.float32(rate)) * 4294967296.0))


def keep_masks(n_tiles, T, width, rate, seed, tile_index_base=0, sites=(1, 2)):
    """uint8 [n_tiles, T, len(sites), width] keep-masks.
    counter = (element // 4, t * 4 + site, tile_lo, tile_hi), key = (seed_lo, seed_hi);
    element e uses output word e % 4."""
    assert width % 4 == 0
    thr = np.uint32(dropout_threshold(rate))
    g = np.arange(n_tiles, dtype=np.uint64) + np.uint64(tile_index_base)
    out = np.empty((n_tiles, T, len(sites), width), dtype=np.uint8)
    j = np.arange(width // 4, dtype=np.uint64)
    for si, site in enumerate(sites):
        for t in range(T):
            c0 = np.broadcast_to(j[None, :], (n_tiles, width // 4))
            c1 = np.full((n_tiles, width // 4), t * 4 + site, dtype=np.uint64)
            c2 = np.broadcast_to((g & _MASK32)[:, None], (n_tiles, width // 4))
            c3 = np.broadcast_to((g >> np.uint64(32))[:, None], (n_tiles, width // 4))
            words = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
            m = np.stack([wd >= thr for wd in words], axis=-1).reshape(n_tiles, width)
            out[:, t, si, :] = m
    return out


# ----------------------------------------------------------------------------------------
# the network
# ----------------------------------------------------------------------------------------

def _bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def _same_pad_s2(h, k=3, s=2):
    """TF 'SAME' padding (before, after) for one spatial dim."""
    out = -(-h // s)
    total = max((out - 1) * s + k - h, 0)
    return total // 2, total - total // 2


class XceptionUQOracle:
    def __init__(self, weights, emulate_bf16=False, dropout=0.1, hidden_layers=2,
                 dropout_sites=(False, True, True), threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.q = _bf16 if emulate_bf16 else (lambda x: x)
        self.emulate = emulate_bf16
        self.rate = float(np.float32(dropout))
        self.hidden_layers = hidden_layers
        self.sites = tuple(dropout_sites)
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in weights.items()}
        self._prep()

    # BN folded to scale/shift in fp32: y = x * scale + shift
    def _bn(self, name):
        g, b = self.w[f"{name}/gamma"], self.w[f"{name}/beta"]
        m, v = self.w[f"{name}/moving_mean"], self.w[f"{name}/moving_variance"]
        scale = g / torch.sqrt(v + BN_EPS)
        return scale, b - m * scale

    def _prep(self):
        q = self.q
        self.conv = {}
        for kind, name, cin, cout in layer_table():
            if kind.startswith("conv3x3"):
                k = self.w[f"{name}/kernel"].permute(3, 2, 0, 1).contiguous()      # OIHW
                # conv1 runs in fp32 on CUDA cores (raw fp32 weights); conv2 is a bf16 GEMM
                self.conv[name] = (k if name == "block1_conv1" else q(k),) + self._bn(f"{name}_bn")
            elif kind == "res":
                k = self.w[f"{name}/kernel"].permute(3, 2, 0, 1).contiguous()
                self.conv[name] = (q(k),) + self._bn(f"{name}_bn")
            else:
                dw = self.w[f"{name}/depthwise_kernel"].permute(2, 3, 0, 1).contiguous()   # [C,1,3,3] fp32
                pw = self.w[f"{name}/pointwise_kernel"].permute(3, 2, 0, 1).contiguous()
                # depthwise weights are bf16 tensor-core operands on the CUDA path (fp32 in the fp32 tier)
                self.conv[name] = (q(dw), q(pw)) + self._bn(f"{name}_bn")

    # ---- pre-processing: tf.image.per_image_standardization (results.py:255)
    @staticmethod
    def tile_stats(tiles_u8: np.ndarray):
        x = torch.from_numpy(tiles_u8).to(torch.float64).reshape(tiles_u8.shape[0], -1)
        n = x.shape[1]
        mean = x.mean(1)
        var = torch.clamp((x * x).mean(1) - mean * mean, min=0.0)
        std = torch.maximum(torch.sqrt(var), torch.tensor(1.0 / np.sqrt(n), dtype=torch.float64))
        return mean.to(torch.float32), (1.0 / std).to(torch.float32)

    def standardize(self, tiles_u8: np.ndarray) -> torch.Tensor:
        """uint8 NHWC -> float32 NCHW, (x - mean) * (1/std)"""
        mean, inv = self.tile_stats(tiles_u8)
        x = torch.from_numpy(tiles_u8).to(torch.float32).permute(0, 3, 1, 2)
        return (x - mean[:, None, None, None]) * inv[:, None, None, None]

    def _affine(self, x, scale, shift):
        return x * scale[None, :, None, None] + shift[None, :, None, None]

    def _sep(self, x, name, relu_in=False):
        dw, pw, scale, shift = self.conv[name]
        if relu_in:
            x = F.relu(x)
        x = self.q(F.conv2d(x, dw, padding=1, groups=x.shape[1]))       # depthwise 3x3 same
        return self._affine(F.conv2d(x, pw), scale, shift)               # pointwise + BN (pre-rounding)

    def _res(self, x, name):
        k, scale, shift = self.conv[name]
        return self.q(self._affine(F.conv2d(x[:, :, ::2, ::2], k), scale, shift))  # 1x1 stride 2, no pad

    def _pool(self, x):
        pt, pb = _same_pad_s2(x.shape[2])
        pl, pr = _same_pad_s2(x.shape[3])
        x = F.pad(x, (pl, pr, pt, pb), value=float("-inf"))
        return F.max_pool2d(x, 3, 2)

    def backbone(self, tiles_u8: np.ndarray, stages=None):
        """-> float32 features [n, 2048]; `stages` (dict) collects NHWC copies of stage outputs."""
        q = self.q

        def keep(name, t):
            if stages is not None:
                stages[name] = t.permute(0, 2, 3, 1).contiguous().numpy()

        x = self.standardize(tiles_u8)
        k, s, b = self.conv["block1_conv1"]
        x = q(F.relu(self._affine(F.conv2d(x, k, stride=2), s, b)))
        keep("block1_conv1", x)
        k, s, b = self.conv["block1_conv2"]
        x = q(F.relu(self._affine(F.conv2d(x, k), s, b)))
        keep("block1_conv2", x)
        for bi, _, _ in ENTRY_BLOCKS:
            res = self._res(x, f"block{bi}_res")
            y = q(F.relu(self._sep(x, f"block{bi}_sepconv1", relu_in=(bi != 2))))
            y = q(self._sep(y, f"block{bi}_sepconv2"))
            x = q(self._pool(y) + res)
            keep(f"block{bi}", x)
        for bi in MIDDLE_BLOCKS:
            y = q(F.relu(self._sep(x, f"block{bi}_sepconv1", relu_in=True)))
            y = q(F.relu(self._sep(y, f"block{bi}_sepconv2")))
            x = q(self._sep(y, f"block{bi}_sepconv3") + x)
            keep(f"block{bi}", x)
        res = self._res(x, "block13_res")
        y = q(F.relu(self._sep(x, "block13_sepconv1", relu_in=True)))
        y = q(self._sep(y, "block13_sepconv2"))
        x = q(self._pool(y) + res)
        keep("block13", x)
        x = q(F.relu(self._sep(x, "block14_sepconv1")))
        x = q(F.relu(self._sep(x, "block14_sepconv2")))
        keep("block14", x)
        return x.mean(dim=(2, 3))

    def head(self, feats: torch.Tensor, masks: np.ndarray):
        """feats [n,2048] fp32; masks uint8 [n,T,n_sites_enabled,width] -> softmax [T,n,classes].
        Dense+ReLU, Dropout after each hidden layer (sites 1..), prelogits, softmax.

        Dropout is applied as x*keep with the 1/(1-rate) factor moved onto the next layer's fp32
        accumulator: (x*keep/(1-p)) @ W == ((x*keep) @ W) / (1-p).  That is where the CUDA head
        applies it (the masked bf16 operand stays an exact selection of the stored activation)."""
        q = self.q
        T = masks.shape[1]
        one = np.float32(1.0)
        inv_keep = one / (one - np.float32(self.rate))
        m = torch.from_numpy(masks).to(torch.float32)
        out = []
        for t in range(T):
            x, pending, si = feats, one, 0
            if self.sites[0]:
                x, pending, si = x * m[:, t, si, :FEATURES], inv_keep, si + 1
            x = q(x)
            for i in range(self.hidden_layers):
                W, b = q(self.w[f"hidden_{i}/kernel"]), self.w[f"hidden_{i}/bias"]
                h = q(F.relu((x @ W) * pending + b))
                pending = one
                if self.sites[i + 1]:
                    h, pending, si = h * m[:, t, si, :h.shape[1]], inv_keep, si + 1
                x = h
            W, b = self.w["prelogits/kernel"], self.w["prelogits/bias"]
            out.append(torch.softmax((x @ W) * pending + b, dim=1))
        return torch.stack(out)

    def predict_uq(self, tiles_u8, T=30, seed=0, tile_index_base=0, masks=None,
                   reference_schedule=False, return_features=False):
        """-> (mean [n,classes], std [n,classes]) float32, population std over T samples."""
        n = tiles_u8.shape[0]
        enabled = [i for i, e in enumerate(self.sites) if e]
        if masks is None:
            width = max(self.w["hidden_0/kernel"].shape[1], FEATURES if self.sites[0] else 0)
            masks = keep_masks(n, T, width, self.rate, seed, tile_index_base, sites=enabled)
        with torch.no_grad():
            if reference_schedule:
                # what Slideflow does: T full forward passes of backbone + head
                probs = torch.stack([self.head(self.backbone(tiles_u8), masks[:, t:t + 1])[0]
                                     for t in range(T)])
                feats = None
            else:
                feats = self.backbone(tiles_u8)
                probs = self.head(feats, masks)
            mean = probs.mean(0)
            std = torch.sqrt(((probs - mean[None]) ** 2).mean(0))
        if return_features:
            return mean.numpy(), std.numpy(), None if feats is None else feats.numpy()
        return mean.numpy(), std.numpy()
