"""oracle/field.py -- TEST INFRASTRUCTURE, not product code.

CPU (torch fp32) restatement of the BRDF field the reference builds from
tiny-cuda-nn (reference model/brdf.py:222-260):

    tcnn.NetworkWithInputEncoding(3, 5, HashGrid{32 levels x 2 features, 2^19, base 16,
    x1.3}, FullyFusedMLP{64 neurons, 2 hidden layers, ReLU, no output activation})

tiny-cuda-nn ("tested with 1.7", unpinned git HEAD, reference scripts/conda_env.sh:10)
is an un-vendored third-party dependency and is absent from /root/reference, so the
algorithm is restated from its published design (Mueller et al. 2022, multiresolution hash
encoding; tiny-cuda-nn encodings/grid.h and networks/fully_fused_mlp.cu as documented):
PARITY WITH tcnn ITSELF IS UNPINNED -- nothing in the reference holds a known-answer
vector for it.  The contract is self-consistency: the CUDA kernels in
iris_b200/csrc/field.cu must match THIS file, rounding points included.

Definitions (every item "tcnn-1.7-compatible by construction, unverified"):
  level l:  scale_l = fp32(16 * (double)1.3f ^ l - 1)   (evaluated in float64, rounded once)
            res_l   = ceil(scale_l) + 1
            size_l  = min(next_multiple(res_l^3, 8), 2^19)      entries of 2 features
  lookup:   pos = fmaf(scale_l, x, 0.5); cell = floor(pos); w = pos - cell;
            cell -> int32 -> uint32 (wraps for negative x: the reference feeds x*2-1 in [-1,1],
            model/brdf.py:255)
            index = dense x + y*res + z*res^2 (uint32 arithmetic) if res^3 fits the level,
                    else (x*1) ^ (y*2654435761) ^ (z*805459861);   index %= size_l
            feature = sum over 8 corners of w_corner * table[index]   (fp32 accumulate,
            corner order bit0=x, bit1=y, bit2=z), then rounded to fp16
  params:   one flat fp32 vector [ W1(64x64) | W2(64x64) | W3(16x64) | level 0 | ... | level 31 ],
            W stored row-major [out][in]; grid values and weights are used at fp16 precision
  MLP:      h1 = fp16(relu(W1 x)), h2 = fp16(relu(W2 h1)), y = fp16(W3 h2)  (fp32 accumulate)
  output:   fp16 tensor (N,16) sliced to 5 -- the reference then applies .sigmoid() IN fp16
            and .float() (model/brdf.py:255-259)
  backward: straight-through across every fp16 rounding, fp32 gradients.
"""
from __future__ import annotations

import math

import numpy as np
import torch

N_LEVELS = 32
N_FEAT = 2
LOG2_HASHMAP = 19
BASE_RES = 16
PER_LEVEL_SCALE = 1.3
WIDTH = 64
N_OUT_PAD = 16
N_MLP = WIDTH * WIDTH * 2 + N_OUT_PAD * WIDTH          # 9216
PRIME_Y = 2654435761
PRIME_Z = 805459861


def level_table():
    """(scale fp32, res, size, offset-in-entries) per level and the total entry count."""
    out, off = [], 0
    base = float(np.float32(PER_LEVEL_SCALE))
    for l in range(N_LEVELS):
        # tcnn evaluates exp2f(l*log2f(1.3f))*16-1 whose last bits depend on the libm at hand; the restatement pins the
        # value portably: the same expression in float64, rounded once to fp32
        scale = np.float32(float(BASE_RES) * math.pow(base, l) - 1.0)
        res = int(math.ceil(float(scale))) + 1
        dense = res ** 3
        size = min(dense, (2 ** 32 - 1) // 2)
        size = (size + 7) // 8 * 8
        size = min(size, 1 << LOG2_HASHMAP)
        out.append((scale, res, size, off))
        off += size
    return out, off


LEVELS, N_ENTRIES = level_table()
N_GRID = N_ENTRIES * N_FEAT
N_PARAMS = N_MLP + N_GRID


def round16(x):
    """fp16 rounding with a straight-through gradient."""
    return x + (x.half().float() - x).detach()


def init_params(seed=0):
    """tcnn-style init: grid U(-1e-4,1e-4); MLP Xavier-uniform."""
    g = torch.Generator().manual_seed(seed)
    p = torch.empty(N_PARAMS)
    o = 0
    for fan_out, fan_in in ((WIDTH, WIDTH), (WIDTH, WIDTH), (N_OUT_PAD, WIDTH)):
        b = math.sqrt(6.0 / (fan_in + fan_out))
        n = fan_out * fan_in
        p[o:o + n] = (torch.rand(n, generator=g) * 2 - 1) * b
        o += n
    p[o:] = (torch.rand(N_GRID, generator=g) * 2 - 1) * 1e-4
    return p


def grid_indices(x):
    """x (N,3) fp32 -> idx (N,32,8) int64 absolute entry index, w (N,32,8) fp32 corner weights."""
    N = x.shape[0]
    idx = torch.empty(N, N_LEVELS, 8, dtype=torch.int64)
    wts = torch.empty(N, N_LEVELS, 8, dtype=torch.float32)
    xn = x.detach().numpy().astype(np.float32)
    M = np.uint64(0xFFFFFFFF)
    for l, (scale, res, size, off) in enumerate(LEVELS):
        # pos = fmaf(scale, x, 0.5): emulate the single rounding with float64 (exact product, one rounding)
        pos = (xn.astype(np.float64) * np.float64(scale) + 0.5).astype(np.float32)
        fl = np.floor(pos)
        frac = (pos - fl).astype(np.float32)
        cell = fl.astype(np.int64).astype(np.int32).astype(np.int64) & 0xFFFFFFFF      # uint32 wrap
        dense = res ** 3 <= size
        for c in range(8):
            cx = (cell[:, 0] + (c & 1)) & 0xFFFFFFFF
            cy = (cell[:, 1] + ((c >> 1) & 1)) & 0xFFFFFFFF
            cz = (cell[:, 2] + ((c >> 2) & 1)) & 0xFFFFFFFF
            if dense:
                # stride loop of tcnn grid_index: strides 1, res, res^2 all <= size here
                i = (cx + cy * res + cz * res * res) & 0xFFFFFFFF
            else:
                i = (cx ^ ((cy * PRIME_Y) & 0xFFFFFFFF) ^ ((cz * PRIME_Z) & 0xFFFFFFFF)) & 0xFFFFFFFF
            i = i % size
            wx = frac[:, 0] if (c & 1) else np.float32(1.0) - frac[:, 0]
            wy = frac[:, 1] if (c & 2) else np.float32(1.0) - frac[:, 1]
            wz = frac[:, 2] if (c & 4) else np.float32(1.0) - frac[:, 2]
            idx[:, l, c] = torch.from_numpy((i + off).astype(np.int64))
            wts[:, l, c] = torch.from_numpy(((wx * wy).astype(np.float32) * wz).astype(np.float32))
    return idx, wts


def encode(x, params):
    """(N,3) -> (N,64) fp16-rounded features (fp32 storage), differentiable wrt params."""
    idx, w = grid_indices(x)
    table = round16(params[N_MLP:]).view(N_ENTRIES, N_FEAT)
    N = x.shape[0]
    feat = torch.zeros(N, N_LEVELS, N_FEAT)
    for c in range(8):                                    # fixed corner order, fp32 accumulate
        feat = feat + w[:, :, c, None] * table[idx[:, :, c]]
    return round16(feat.reshape(N, N_LEVELS * N_FEAT))


def mlp(feat, params):
    W1 = round16(params[0:4096]).view(64, 64)
    W2 = round16(params[4096:8192]).view(64, 64)
    W3 = round16(params[8192:9216]).view(16, 64)
    h = round16(torch.relu(feat @ W1.t()))
    h = round16(torch.relu(h @ W2.t()))
    return round16(h @ W3.t())


def network(x, params):
    """The tcnn module's forward: (N,3) in [-1,1]-ish -> (N,5) fp16-rounded values (fp32 storage)."""
    return mlp(encode(x, params), params)[:, :5]


def material(position, params, voxel_min, voxel_max):
    """NGPBRDF.forward (reference model/brdf.py:243-260) on top of `network`.

    fp32 op order: (p - vmin) / (vmax - vmin), *2, -1; sigmoid evaluated on the fp16 output and
    rounded to fp16 again (the reference calls .sigmoid() on tcnn's half tensor), then .float().
    """
    vmin = np.float32(voxel_min)
    rng = np.float32(voxel_max - voxel_min)
    x = (position - float(vmin)) / float(rng)
    x = x * 2.0 - 1.0
    y = network(x, params)
    s = round16(torch.sigmoid(y))
    return {"albedo": s[:, 0:3], "roughness": s[:, 3:4] * 0.98 + 0.02, "metallic": s[:, 4:5]}


class CpuHashGridMLP(torch.nn.Module):
    """Stand-in for tcnn.NetworkWithInputEncoding(3,5,enc,net) used when the REFERENCE code is run
    on CPU to generate golden vectors (oracle/refharness.py): same constructor arity, one flat fp32
    parameter named `params`, fp16 output tensor."""

    def __init__(self, n_input_dims=3, n_output_dims=5, encoding_config=None, network_config=None, seed=0):
        super().__init__()
        assert n_input_dims == 3 and n_output_dims == 5
        self.params = torch.nn.Parameter(init_params(seed))

    def forward(self, x):
        return _ToHalfSTE.apply(network(x.float(), self.params))


class _ToHalfSTE(torch.autograd.Function):
    """Return a genuine fp16 tensor (so the reference's .sigmoid() runs in fp16) while keeping fp32
    gradients flowing back to `params`."""

    @staticmethod
    def forward(ctx, y):
        return y.half()

    @staticmethod
    def backward(ctx, g):
        return g.float()
