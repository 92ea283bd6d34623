"""Host-side key arithmetic of `jax.random` (a few threefry2x32 blocks in Python integers).

The kernels derive every per-trajectory / per-fragment key on the device; the host only needs
`PRNGKey(seed)` and - for the `fold_in` sharding mode of `distributed.solve_many_sharded` - one block per
rank.  threefry2x32-20 as published by Salmon et al. (Random123) and used by `jax/_src/prng.py`."""

_M = 0xFFFFFFFF
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def threefry2x32(key, x0, x1):
    k0, k1 = int(key[0]) & _M, int(key[1]) & _M
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0 = (x0 + ks[0]) & _M
    x1 = (x1 + ks[1]) & _M
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 = (x0 + x1) & _M
            x1 = ((x1 << r) | (x1 >> (32 - r))) & _M
            x1 ^= x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M
        x1 = (x1 + ks[(i + 2) % 3] + i + 1) & _M
    return x0, x1


def PRNGKey(seed, x64=None):
    """Raw key data of `jax.random.PRNGKey(seed)`.

    JAX converts an integer seed to the default integer type first: int32 unless `jax_enable_x64` is set.  The high
    key word is therefore 0 in 32-bit mode (the seed wraps: -1 -> [0, 0xFFFFFFFF], 2**33 + 5 -> [0, 5]) and
    `(seed >> 32) & 0xFFFFFFFF` in 64-bit mode (-1 -> [0xFFFFFFFF, 0xFFFFFFFF]).  `x64=None` reads the package
    switch `config "enable_x64"`; the solver passes True for double-precision runs."""
    if x64 is None:
        from . import config
        x64 = config.get("enable_x64")
    seed = int(seed)
    return ((seed >> 32) & _M) if x64 else 0, seed & _M


def fold_in(key, data):
    """`jax.random.fold_in(key, data)`: both words of the block with counter (0, data)."""
    return threefry2x32(key, 0, int(data) & _M)


def split(key, n, layout="original"):
    """`jax.random.split(key, n)` as a list of (a, b) pairs (host-side, small n)."""
    n = int(n)
    if layout == "partitionable":
        return [threefry2x32(key, i >> 32, i & _M) for i in range(n)]
    words = []
    for k in range(2 * n):
        words.append(threefry2x32(key, k, n + k)[0] if k < n else threefry2x32(key, k - n, k)[1])
    return [(words[2 * i], words[2 * i + 1]) for i in range(n)]
