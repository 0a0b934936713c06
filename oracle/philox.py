"""ORACLE (test infrastructure): the counter-based random numbers of the device's train-time ops.

The product draws dropout / stochastic-depth keep decisions with Philox4x32-10 from
(seed, step, layer id, element index) (myconvnet_b200/csrc/dropout.cu); TensorFlow's own generator
(reference convnet.py:2506, tf.nn.dropout) is a different stream, so parity of the RANDOM ops is
checked with the masks regenerated here from the same definition — the published Philox4x32-10
algorithm (Salmon et al., SC'11) restated in numpy — and the reference's formulas applied to them:
dropout keeps u >= rate and scales by 1/(1-rate); stochastic depth keeps a
whole sample (convnet.py:2506-2509).  Known-answer vectors of the Random123 distribution pin it
(tests/test_oracle.py)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: uint32 array [..., 4]; key: (k0, k1).  Returns uint32 [..., 4]."""
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def _draw(idx, step, seed, layer):
    idx = np.asarray(idx, dtype=np.uint64)
    ctr = np.stack([idx & MASK, idx >> np.uint64(32), np.full_like(idx, step), np.zeros_like(idx)], axis=-1)
    return philox4x32_10(ctr.astype(np.uint32), (seed, layer))


def uniform(bits):
    return (bits >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def dropout_keep(n, rate, seed, step, layer):
    """Keep mask (bool [n]) of mcn_dropout: element i uses output i % 4 of counter i // 4."""
    q = np.arange((n + 3) // 4)
    u = uniform(_draw(q, step, seed, layer)).reshape(-1)[:n]
    return u >= np.float32(rate)


def survive(batch, rate, seed, step, layer):
    """Per-sample survival mask (bool [batch]) of mcn_sd_add_*: output 0 of counter = sample."""
    u = uniform(_draw(np.arange(batch), step, seed, layer))[:, 0]
    return u >= np.float32(rate)
