"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Torch-CPU restatement of the multiresolution grid encoding of tiny-cuda-nn
(`tcnn.Encoding` with `otype: HashGrid | DenseGrid | TiledGrid`, `interpolation: Linear`),
the position encoding of the reference's `arch: ngp` field
(robust_e_nerf/external/ngp.py:166-170, configs/train/*.yaml:61-69) and of its direction
encoding (robust_e_nerf/external/sh_encoder.py:15-193, restated for degree <= 4).

PARITY UNPINNED: tiny-cuda-nn is an un-vendored, UNPINNED dependency
(`environment.yml:31`: git HEAD of NVlabs/tiny-cuda-nn), CUDA-only and absent from this image, and the
reference holds no golden vector for it.  This file restates the published algorithm
(include/tiny-cuda-nn/encodings/grid.h of tiny-cuda-nn v1.6: `grid_scale`, `grid_resolution`,
`pos_fract`, `grid_index`, `coherent_prime_hash`, the parameter offset table of
`GridEncodingTemplated`), with plain differentiable torch ops so that first- and second-order autograd
(the reference's d logI / dt double backward) work.  The per-level scale is evaluated in float64 and
rounded to fp32 once, on the host, by the oracle and by the CUDA path alike (tiny-cuda-nn evaluates
`exp2f` on the device; the two can differ in the last ulp of the scale).
"""
import math

import numpy as np
import torch

PRIMES = (1, 2654435761, 805459861)          # coherent_prime_hash, first three of seven
GRID_TYPES = ("Hash", "Dense", "Tiled")


def grid_meta(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16,
              per_level_scale=1.4472692012786865, grid_type="Hash"):
    """Per-level (scale f32, resolution, entries, entry offset) + total entries, following the constructor of
    `GridEncodingTemplated` and `grid_scale` / `grid_resolution`."""
    assert grid_type in GRID_TYPES
    log2_pls = math.log2(per_level_scale)
    levels, offset = [], 0
    for lvl in range(n_levels):
        scale = float(np.float32(math.exp2(lvl * log2_pls) * base_resolution - 1.0))
        res = int(math.ceil(scale)) + 1
        n = min(res ** 3, (2 ** 32 - 1) // 2)
        n = (n + 7) // 8 * 8                                   # next_multiple(params_in_level, 8)
        if grid_type == "Tiled":
            n = min(n, base_resolution ** 3)
        elif grid_type == "Hash":
            n = min(n, 1 << log2_hashmap_size)
        levels.append(dict(scale=scale, resolution=res, entries=n, offset=offset))
        offset += n
    return dict(levels=levels, n_entries=offset, n_features=n_features_per_level,
                n_params=offset * n_features_per_level, n_output_dims=n_levels * n_features_per_level,
                grid_type=grid_type)


def grid_index(meta_level, grid_type, ix, iy, iz):
    """`grid_index<3, HASH_TYPE>`: dense index while the stride fits the level's table, else the prime hash;
    always modulo the table size.  ix / iy / iz are int64 tensors of (non-negative) grid coordinates."""
    hs, res = meta_level["entries"], meta_level["resolution"]
    index = torch.zeros_like(ix)
    stride = 1
    for c in (ix, iy, iz):
        if stride > hs:
            break
        index = index + c * stride
        stride *= res
    if grid_type == "Hash" and hs < stride:
        m = 0xFFFFFFFF
        index = ((ix * PRIMES[0]) & m) ^ ((iy * PRIMES[1]) & m) ^ ((iz * PRIMES[2]) & m)
    # uint32 wrap of the dense index cannot happen for res^3 <= 2^31
    return index % hs


def encode(x, params, meta):
    """x (N,3) in the unit cube (any float dtype), params (n_params,) -> (N, n_levels * F).
    Trilinear interpolation of the 8 corner feature vectors of every level (`kernel_grid`)."""
    F = meta["n_features"]
    table = params.reshape(-1, F)
    outs = []
    for lv in meta["levels"]:
        scale = lv["scale"]
        # pos = fmaf(scale, x, 0.5f): single rounding (the double product of two floats is exact)
        if x.dtype == torch.float32:
            pos = (x.double() * scale + 0.5).float()
            pos = x * scale + 0.5 + (pos - (x * scale + 0.5)).detach()     # value of the fma, gradient = scale
        else:
            pos = x * scale + 0.5
        base = torch.floor(pos.detach())
        frac = pos - base
        gi = base.to(torch.int64)
        feat = 0
        for corner in range(8):
            w = 1
            idx = []
            for dim in range(3):
                if (corner >> dim) & 1:
                    w = w * frac[:, dim]
                    idx.append(gi[:, dim] + 1)
                else:
                    w = w * (1 - frac[:, dim])
                    idx.append(gi[:, dim])
            e = grid_index(lv, meta["grid_type"], idx[0], idx[1], idx[2]) + lv["offset"]
            feat = feat + w[:, None] * table[e]
        outs.append(feat)
    return torch.cat(outs, dim=-1)


def init_params(meta, generator=None, dtype=torch.float32):
    """tiny-cuda-nn initialises grid parameters U(-1e-4, 1e-4) (`GridEncoding::initialize_params`) from its own
    pcg32 stream, which cannot be reproduced here: drawn from torch's generator instead."""
    u = torch.rand(meta["n_params"], generator=generator, dtype=torch.float32)
    return ((u * 2 - 1) * 1e-4).to(dtype)


# ----------------------------------------------------------------------------- spherical harmonics
def sh_encode(d, degree=4):
    """`SHEncoder.forward` (robust_e_nerf/external/sh_encoder.py:27-193) for degree <= 4: real spherical
    harmonics of the raw (unit) direction, 16 coefficients."""
    assert 1 <= degree <= 4, "the restatement covers the degrees the shipped configs use (dir_encoding.degree: 4)"
    x, y, z = d.unbind(-1)
    xy, xz, yz = x * y, x * z, y * z
    x2, y2, z2 = x * x, y * y, z * z
    out = [torch.full_like(x, 0.28209479177387814)]
    if degree > 1:
        out += [-0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x]
    if degree > 2:
        out += [1.0925484305920792 * xy, -1.0925484305920792 * yz,
                0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
                0.54627421529603959 * x2 - 0.54627421529603959 * y2]
    if degree > 3:
        out += [0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
                0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0),
                0.45704579946446572 * x * (1.0 - 5.0 * z2), 1.4453057213202769 * z * (x2 - y2),
                0.59004358992664352 * x * (-x2 + 3.0 * y2)]
    return torch.stack(out, dim=-1)
