"""Oracle: the slice of `jax.random` that decides the reference's random stream.

TEST INFRASTRUCTURE ONLY (see `oracle/__init__.py`).

The reference draws every Wiener increment through `jax.random.PRNGKey / split / normal`
(reference `pychastic/sde_solver.py:223,292,299`, `pychastic/vectorized_I_generation.py:267-349`).
`jax` is an un-vendored dependency (pin: `jax==0.4.35`, reference `docs/requirements.txt:16`),
so its published algorithm is restated here:

* threefry2x32, 20 rounds (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
  SC'11; `jax/_src/prng.py::threefry2x32`),
* `threefry_2x32(key, counts)` half/half counter layout, `split`, `fold_in`,
* `_threefry_random_bits_original` (the non-partitionable layout, default for jax<0.5) for
  32- and 64-bit draws, plus the partitionable layout (jax>=0.5 default) as an option,
* `uniform` mantissa trick and `normal = sqrt(2) * erf_inv(uniform(nextafter(-1,0), 1))`,
* XLA's `erf_inv` expansion (M. Giles, "Approximating the erfinv function", single and
  double precision polynomials; `xla/client/lib/math.cc::ErfInv32/ErfInv64`).

All arithmetic is numpy, vectorised over the counter array.
"""
import numpy as np

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = U32(0x1BD11BDA)


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """One threefry2x32-20 block per element. Inputs broadcast; returns (y0, y1) uint32."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=U32)
        k1 = np.asarray(k1, dtype=U32)
        x0 = np.array(x0, dtype=U32, copy=True)
        x1 = np.array(x1, dtype=U32, copy=True)
        ks = (k0, k1, k0 ^ k1 ^ _PARITY)
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + U32(i + 1)
    return x0, x1


def PRNGKey(seed, x64=False):
    """`jax.random.PRNGKey(seed)` raw key data (jax/_src/prng.py::random_seed + threefry_seed).

    An int seed becomes `jnp.asarray(np.int64(seed))`: int32 (wrapped) unless jax_enable_x64, then
    k1 = convert(shift_right_logical(seed, 32)) - which is 0 for a 32-bit seed - and k2 = seed & 0xFFFFFFFF.
    So 32-bit mode gives [0, lo32(seed)] (seed=-1 -> [0, 0xFFFFFFFF]) and 64-bit mode [hi32(seed), lo32(seed)]."""
    seed = int(seed)
    return np.array([((seed >> 32) & 0xFFFFFFFF) if x64 else 0, seed & 0xFFFFFFFF], dtype=U32)


def threefry_2x32(key, counts):
    """`jax._src.prng.threefry_2x32`: counters split first-half / second-half."""
    counts = np.asarray(counts, dtype=U32).ravel()
    n = counts.size
    odd = n % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, dtype=U32)])
    h = counts.size // 2
    y0, y1 = threefry2x32(key[0], key[1], counts[:h], counts[h:])
    out = np.concatenate([y0, y1])
    return out[:-1] if odd else out


def split(key, n, layout="original"):
    """`jax.random.split(key, n)` -> uint32[n, 2]."""
    n = int(n)
    if layout == "original":
        return threefry_2x32(key, np.arange(2 * n, dtype=np.uint64).astype(U32)).reshape(n, 2)
    if layout == "partitionable":
        idx = np.arange(n, dtype=np.uint64)
        y0, y1 = threefry2x32(key[0], key[1], (idx >> np.uint64(32)).astype(U32), idx.astype(U32))
        return np.stack([y0, y1], axis=1)
    raise ValueError(layout)


def fold_in(key, data):
    """`jax.random.fold_in(key, data)` = threefry_2x32(key, [0, data])."""
    return threefry_2x32(key, np.array([0, int(data) & 0xFFFFFFFF], dtype=U32))


def random_bits(key, bit_width, shape, layout="original"):
    """`jax._src.prng.threefry_random_bits` for bit_width in {32, 64}."""
    shape = tuple(int(s) for s in np.atleast_1d(shape)) if not isinstance(shape, tuple) else shape
    size = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    if layout == "original":
        if bit_width == 32:
            bits = threefry_2x32(key, np.arange(size, dtype=np.uint64).astype(U32))
            return bits.reshape(shape)
        if bit_width == 64:
            bits = threefry_2x32(key, np.arange(2 * size, dtype=np.uint64).astype(U32))
            hi = bits[:size].astype(np.uint64)
            lo = bits[size:].astype(np.uint64)
            return ((hi << np.uint64(32)) | lo).reshape(shape)
    elif layout == "partitionable":
        idx = np.arange(size, dtype=np.uint64)
        y0, y1 = threefry2x32(key[0], key[1], (idx >> np.uint64(32)).astype(U32), idx.astype(U32))
        if bit_width == 32:
            return (y0 ^ y1).reshape(shape)
        if bit_width == 64:
            return ((y0.astype(np.uint64) << np.uint64(32)) | y1.astype(np.uint64)).reshape(shape)
    raise ValueError((bit_width, layout))


# ---------------------------------------------------------------------------------------
# erf_inv (XLA expansion; coefficients highest order first)
# ---------------------------------------------------------------------------------------
ERFINV32_LT5 = np.array([2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
                         0.00021858087, -0.00125372503, -0.00417768164, 0.246640727,
                         1.50140941], dtype=np.float32)
ERFINV32_GE5 = np.array([-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
                         0.00573950773, -0.0076224613, 0.00943887047, 1.00167406,
                         2.83297682], dtype=np.float32)

ERFINV64_LT6_25 = np.array([
    -3.6444120640178196996e-21, -1.685059138182016589e-19, 1.2858480715256400167e-18,
    1.115787767802518096e-17, -1.333171662854620906e-16, 2.0972767875968561637e-17,
    6.6376381343583238325e-15, -4.0545662729752068639e-14, -8.1519341976054721522e-14,
    2.6335093153082322977e-12, -1.2975133253453532498e-11, -5.4154120542946279317e-11,
    1.051212273321532285e-09, -4.1126339803469836976e-09, -2.9070369957882005086e-08,
    4.2347877827932403518e-07, -1.3654692000834678645e-06, -1.3882523362786468719e-05,
    0.0001867342080340571352, -0.00074070253416626697512, -0.0060336708714301490533,
    0.24015818242558961693, 1.6536545626831027356], dtype=np.float64)
ERFINV64_LT16 = np.array([
    2.2137376921775787049e-09, 9.0756561938885390979e-08, -2.7517406297064545428e-07,
    1.8239629214389227755e-08, 1.5027403968909827627e-06, -4.013867526981545969e-06,
    2.9234449089955446044e-06, 1.2475304481671778723e-05, -4.7318229009055733981e-05,
    6.8284851459573175448e-05, 2.4031110387097893999e-05, -0.0003550375203628474796,
    0.00095328937973738049703, -0.0016882755560235047313, 0.0024914420961078508066,
    -0.0037512085075692412107, 0.005370914553590063617, 1.0052589676941592334,
    3.0838856104922207635], dtype=np.float64)
ERFINV64_GE16 = np.array([
    -2.7109920616438573243e-11, -2.5556418169965252055e-10, 1.5076572693500548083e-09,
    -3.7894654401267369937e-09, 7.6157012080783393804e-09, -1.4960026627149240478e-08,
    2.9147953450901080826e-08, -6.7711997758452339498e-08, 2.2900482228026654717e-07,
    -9.9298272942317002539e-07, 4.5260625972231537039e-06, -1.9681778105531670567e-05,
    7.5995277030017761139e-05, -0.00021503011930044477347, -0.00013871931833623122026,
    1.0103004648645343977, 4.8499064014085844221], dtype=np.float64)


def erf_inv32(x):
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -np.log1p(-x * x)
        lt = w < np.float32(5.0)
        ws = np.where(lt, w - np.float32(2.5), np.sqrt(w) - np.float32(3.0)).astype(np.float32)
        p = np.where(lt, ERFINV32_LT5[0], ERFINV32_GE5[0]).astype(np.float32)
        for i in range(1, 9):
            c = np.where(lt, ERFINV32_LT5[i], ERFINV32_GE5[i]).astype(np.float32)
            p = (c + p * ws).astype(np.float32)
        res = (p * x).astype(np.float32)
        return np.where(np.abs(x) == np.float32(1.0), x * np.float32(np.inf), res).astype(np.float32)


def erf_inv64(x):
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -np.log1p(-x * x)
        lt625 = w < 6.25
        lt16 = w < 16.0
        ws = np.where(lt625, w - 3.125, np.sqrt(w) - np.where(lt16, 3.25, 5.0))

        def coef(i):
            c1 = ERFINV64_LT6_25[i]
            c2 = ERFINV64_LT16[i] if i < 19 else 0.0
            c3 = ERFINV64_GE16[i] if i < 17 else 0.0
            return np.where(lt625, c1, np.where(lt16, c2, c3))

        p = coef(0)
        for i in range(1, 17):
            p = coef(i) + p * ws
        for i in range(17, 19):
            p = np.where(lt16, coef(i) + p * ws, p)
        for i in range(19, 23):
            p = np.where(lt625, coef(i) + p * ws, p)
        res = p * x
        return np.where(np.abs(x) == 1.0, x * np.inf, res)


def uniform_from_bits(bits, dtype, lo, hi):
    """`jax.random.uniform` tail: mantissa-fill, subtract 1, affine map, clamp at lo."""
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        b = (np.asarray(bits, dtype=U32) >> U32(9)) | U32(0x3F800000)
        f = b.view(np.float32) - np.float32(1.0)
        lo = np.float32(lo)
        hi = np.float32(hi)
        u = (f * np.float32(hi - lo) + lo).astype(np.float32)
        return np.maximum(lo, u)
    if dtype == np.float64:
        b = (np.asarray(bits, dtype=np.uint64) >> np.uint64(12)) | np.uint64(0x3FF0000000000000)
        f = b.view(np.float64) - 1.0
        lo = np.float64(lo)
        hi = np.float64(hi)
        u = f * (hi - lo) + lo
        return np.maximum(lo, u)
    raise ValueError(dtype)


def normal_from_bits(bits, dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
        u = uniform_from_bits(bits, dtype, lo, np.float32(1.0))
        return (np.float32(np.sqrt(2)) * erf_inv32(u)).astype(np.float32)
    lo = np.nextafter(np.float64(-1.0), np.float64(0.0))
    u = uniform_from_bits(bits, dtype, lo, np.float64(1.0))
    return np.float64(np.sqrt(2)) * erf_inv64(u)


def uniform(key, shape, dtype=np.float32, lo=0.0, hi=1.0, layout="original"):
    dtype = np.dtype(dtype)
    bits = random_bits(key, 32 if dtype == np.float32 else 64, tuple(shape), layout)
    return uniform_from_bits(bits, dtype, lo, hi)


def normal(key, shape, dtype=np.float32, layout="original"):
    """`jax.random.normal(key, shape, dtype)`."""
    dtype = np.dtype(dtype)
    bits = random_bits(key, 32 if dtype == np.float32 else 64, tuple(shape), layout)
    return normal_from_bits(bits, dtype)
