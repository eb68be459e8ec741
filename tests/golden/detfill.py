"""Deterministic, RNG-library-independent tensor filler shared by the golden
generator (run once in the dev container against the real reference) and the
parity tests (run anywhere, without the reference).

Both sides build a model, then overwrite every entry of ``state_dict()`` with
``det_tensor(key, shape, seed)``.  Because the new model keeps the reference's
``state_dict`` key/shape layout the two models end up
with bit-identical weights without shipping any weight blob.

The generator is a counter-based splitmix64 -> Box-Muller, written in numpy
uint64 arithmetic only, so it does not depend on any library's RNG stream.
"""
import zlib
import numpy as np

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
    return z ^ (z >> np.uint64(31))


def det_uniform(key, n, seed=0):
    """n doubles in (0,1), deterministic in (key, seed)."""
    with np.errstate(over='ignore'):
        base = np.uint64(zlib.crc32(key.encode()) & 0xFFFFFFFF) * np.uint64(0x100000001B3) + np.uint64(seed * 7919 + 1)
        ctr = (np.arange(n, dtype=np.uint64) * np.uint64(0x632BE59BD9B4E019) + base) & _MASK
        bits = _splitmix64(ctr)
    return ((bits >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)


def det_normal(key, shape, seed=0):
    n = int(np.prod(shape)) if len(shape) else 1
    m = (n + 1) // 2
    u1 = det_uniform(key + '/a', m, seed)
    u2 = det_uniform(key + '/b', m, seed)
    r = np.sqrt(-2.0 * np.log(u1))
    z = np.concatenate([r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)])[:n]
    return z.reshape(shape)


def det_tensor(key, shape, seed=0):
    """Value for state_dict entry `key`: magnitudes follow the reference's
    initialisers so activations stay O(1) through the network."""
    shape = tuple(shape)
    z = det_normal(key, shape, seed)
    leaf = key.split('.')[-1]
    parts = key.split('.')
    if 'rp' in parts:
        idx = parts[parts.index('rp') + 2]          # rp.net.<idx>.<leaf>
        if idx in ('1', '4'):                       # LayerNorm
            return (1.0 + 0.1 * z) if leaf == 'weight' else 0.1 * z
        if leaf == 'bias':
            return 0.1 * z
        fan_in = shape[1]
        return z / np.sqrt(fan_in)                  # Linear weight [out, in]
    if leaf == 'scale':                             # NormSE3 scale
        return 1.0 + 0.1 * z
    if leaf == 'w_gate':
        return 1e-3 * z
    if 'weights' in parts:                          # LinearSE3 [d_in, d_out]
        return z / np.sqrt(shape[0])
    if 'null_keys' in parts or 'null_values' in parts:
        return 0.5 * z
    return z                                        # embeddings etc.


def fill_state_dict(module, seed=0):
    """Overwrite every floating-point entry of module.state_dict() in place."""
    import torch
    sd = module.state_dict()
    with torch.no_grad():
        for k, v in sd.items():
            if not torch.is_floating_point(v):
                continue
            if k.endswith('inv_freq'):
                continue
            v.copy_(torch.from_numpy(det_tensor(k, v.shape, seed)).to(v.dtype))
    return module


def det_inputs(name, shape, seed=0, scale=1.0):
    return (scale * det_normal('input/' + name, shape, seed)).astype(np.float32)
