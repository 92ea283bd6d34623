"""ctypes binding of libsdeb200.so (C ABI in include/sdeb200.h) + compiled-program cache.

There is NO CPU fallback: every compute entry point raises if the library or a CUDA device
is missing.  NVRTC compilation needs no GPU, so `compile_source` also works on a CPU-only
box (used by `__graft_entry__.build()` to pre-populate the cubin cache that travels to the
GPU box).
"""
import ctypes
import hashlib
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdeb200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")
CACHE_DIR = os.environ.get("SDEB200_CACHE_DIR", os.path.join(_HERE, "_cubin_cache"))

EXPORTS = (
    "sdeb200_abi_version", "sdeb200_last_error", "sdeb200_set_device", "sdeb200_compile", "sdeb200_free",
    "sdeb200_program_load", "sdeb200_program_destroy", "sdeb200_program_info", "sdeb200_solve",
    "sdeb200_wiener_integrals", "sdeb200_keys_split", "sdeb200_normal", "sdeb200_fp64_peak", "sdeb200_bead_solve",
    "sdeb200_nvrtc_version",
)


class SolveArgs(ctypes.Structure):
    """Mirror of `sdeb200_solve_args_t`."""
    _fields_ = [
        ("key_a", ctypes.c_uint32), ("key_b", ctypes.c_uint32),
        ("n_fragments", ctypes.c_uint32), ("cpr", ctypes.c_uint32),
        ("chunk", ctypes.c_uint32), ("n_snapshots", ctypes.c_uint32),
        ("n_traj_total", ctypes.c_uint64), ("traj_begin", ctypes.c_uint64), ("traj_count", ctypes.c_uint64),
        ("x0", ctypes.c_void_p), ("x0_stride", ctypes.c_int64),
        ("dt", ctypes.c_double), ("dt32", ctypes.c_double),
        ("out_t", ctypes.c_void_p), ("out_x", ctypes.c_void_p), ("out_w", ctypes.c_void_p),
        ("moments", ctypes.c_void_p), ("out_hit", ctypes.c_void_p),
    ]


class WienerArgs(ctypes.Structure):
    """Mirror of `sdeb200_wiener_args_t`."""
    _fields_ = [
        ("key_a", ctypes.c_uint32), ("key_b", ctypes.c_uint32), ("steps", ctypes.c_uint32), ("_pad", ctypes.c_uint32),
        ("d_w", ctypes.c_void_p), ("d_ww", ctypes.c_void_p), ("d_wt", ctypes.c_void_p), ("d_tw", ctypes.c_void_p),
        ("d_www", ctypes.c_void_p),
    ]


class BeadArgs(ctypes.Structure):
    """Mirror of `sdeb200_bead_args_t`."""
    _fields_ = [
        ("key_a", ctypes.c_uint32), ("key_b", ctypes.c_uint32),
        ("n_fragments", ctypes.c_uint32), ("cpr", ctypes.c_uint32),
        ("chunk", ctypes.c_uint32), ("n_snapshots", ctypes.c_uint32),
        ("n_traj_total", ctypes.c_uint64), ("traj_begin", ctypes.c_uint64), ("traj_count", ctypes.c_uint64),
        ("x0", ctypes.c_void_p), ("x0_stride", ctypes.c_int64),
        ("dt", ctypes.c_double), ("radius", ctypes.c_double), ("ks", ctypes.c_double), ("d0", ctypes.c_double),
        ("kb_over_d04", ctypes.c_double), ("ko", ctypes.c_double), ("sigma2", ctypes.c_double),
        ("inv_delta2", ctypes.c_double), ("twist_k", ctypes.c_double), ("linking_number", ctypes.c_double),
        ("workspace", ctypes.c_void_p), ("out_t", ctypes.c_void_p), ("out_x", ctypes.c_void_p), ("out_w", ctypes.c_void_p),
    ]


assert ctypes.sizeof(SolveArgs) == 120 and ctypes.sizeof(WienerArgs) == 56 and ctypes.sizeof(BeadArgs) == 176

_lib = None
_lib_lock = threading.Lock()


class Sdeb200Error(RuntimeError):
    pass


def lib():
    """Load libsdeb200.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise Sdeb200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C stochastic_b200/csrc`). There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.sdeb200_last_error.restype = ctypes.c_char_p
        L.sdeb200_compile.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p,
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
                                      ctypes.POINTER(ctypes.c_void_p)]
        L.sdeb200_free.argtypes = [ctypes.c_void_p]
        L.sdeb200_free.restype = None
        L.sdeb200_program_load.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p,
                                           ctypes.POINTER(ctypes.c_void_p)]
        L.sdeb200_program_destroy.argtypes = [ctypes.c_void_p]
        L.sdeb200_program_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int),
                                           ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        L.sdeb200_solve.argtypes = [ctypes.c_void_p, ctypes.POINTER(SolveArgs), ctypes.c_int, ctypes.c_int,
                                    ctypes.c_size_t, ctypes.c_void_p]
        L.sdeb200_wiener_integrals.argtypes = [ctypes.c_void_p, ctypes.POINTER(WienerArgs), ctypes.c_int,
                                               ctypes.c_size_t, ctypes.c_void_p]
        L.sdeb200_bead_solve.argtypes = [ctypes.c_void_p, ctypes.POINTER(BeadArgs), ctypes.c_int, ctypes.c_size_t,
                                         ctypes.c_uint, ctypes.c_void_p]
        L.sdeb200_keys_split.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
                                         ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.sdeb200_normal.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_void_p]
        L.sdeb200_fp64_peak.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.sdeb200_set_device.argtypes = [ctypes.c_int]
        L.sdeb200_nvrtc_version.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                            ctypes.POINTER(ctypes.c_char_p)]
        _lib = L
    return _lib


def stage_count(m, p, scheme):
    """Normals per step staged through shared memory; must equal SDE_STAGE_COUNT in csrc/sde_core.cuh
    (the emitted sources carry a static_assert)."""
    if m > 1 and scheme == "wagner_platen":
        return 2 * m * p + 3 * m
    if m > 4 and scheme == "euler":
        return m
    return 0


def pick_block(m, p, scheme, f64, preferred=None):
    """CTA size.  Kernels with a staged normal generator (large step bodies) keep their warps in lock-step by a per-step
    barrier: 256-thread CTAs for Wagner-Platen with several noise terms, 128-thread CTAs for Euler with m > 4 (three
    of them per SM at <= 168 registers: two spheres + 6 %, four beads + 10 % over two 256-thread CTAs at 128 registers,
    profiles/README.md); the staging area must fit in 200 KB of shared memory."""
    stage = stage_count(m, p, scheme)
    block = preferred or (256 if stage and scheme != "euler" else 128)
    while block > 32 and stage * block * (8 if f64 else 4) > 200 * 1024:
        block //= 2
    return block


def pick_snap_tile(d, m, block, f64, stage_bytes, n_snapshots, budget=96 * 1024):
    """Snapshots staged per trajectory before a transposed, coalesced write-out (SDE_SNAP_TILE).
    1 = direct stores (final-only runs); otherwise the largest power of two <= 16 whose staging tiles
    (one per warp, pitch 33) fit next to the normals staging area in `budget` bytes per CTA."""
    if n_snapshots < 4:
        return 1
    tile = 16
    while tile > 1:
        if tile <= n_snapshots and stage_bytes + snap_tile_bytes(d, m, block, f64, tile) <= budget:
            return tile
        tile //= 2
    return 1


def snap_tile_bytes(d, m, block, f64, tile):
    return 0 if tile <= 1 else (block // 32) * (1 + d + m) * tile * 33 * (8 if f64 else 4)


def pick_snap_tma(d, m, warps, f64, other_bytes, n_snapshots, budget=96 * 1024):
    """Snapshot tiles stored by the TMA unit (csrc/sde_snap_tma.cuh): (snapshots per tile, CUtensorMapSwizzle), or
    (0, 0) when the output rows cannot be described by a tensor map (n_snapshots * sizeof(real) not a multiple of
    16 bytes) or no tile fits.  Tile rows are 128, 64, 32 or 16 bytes; the largest that is not longer than the run
    and whose boxes (`warps` x (1 + d + m) x 32 rows, plus 1 KB of alignment slack) fit in `budget` bytes per CTA
    next to `other_bytes` of other dynamic shared memory."""
    elem = 8 if f64 else 4
    if n_snapshots < 4 or (n_snapshots * elem) % 16:
        return 0, 0
    for swz, rowb in ((3, 128), (2, 64), (1, 32), (0, 16)):
        tile = rowb // elem
        if tile <= n_snapshots and other_bytes + snap_tma_bytes(d, m, warps, f64, tile) <= budget:
            return tile, swz
    return 0, 0


def snap_tma_bytes(d, m, warps, f64, tile):
    return 0 if tile < 1 else 1024 + warps * (1 + d + m) * 32 * tile * (8 if f64 else 4)


def check(status, what=""):
    if status == 0:
        return
    msg = lib().sdeb200_last_error().decode(errors="replace")
    if status < 0:
        raise ValueError(f"{what}: {msg}")
    raise Sdeb200Error(f"{what}: {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise Sdeb200Error("no CUDA device available: stochastic_b200 has no CPU fallback")
    return torch


def current_stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# ------------------------------------------------------------------------------------------
# compile + cache
# ------------------------------------------------------------------------------------------
def nvrtc_version():
    """(major, minor, path) of the NVRTC the library compiles with (the CUDA toolkit's, not the host process's)."""
    major, minor, path = ctypes.c_int(), ctypes.c_int(), ctypes.c_char_p()
    check(lib().sdeb200_nvrtc_version(ctypes.byref(major), ctypes.byref(minor), ctypes.byref(path)), "NVRTC")
    return major.value, minor.value, (path.value or b"").decode()


def _core_fingerprint():
    h = hashlib.sha256()
    h.update("nvrtc %d.%d\n".encode() % nvrtc_version()[:2])      # cubins of different compilers never mix
    for name in ("sde_core.cuh", "sde_hooks.cuh", "sde_snap_tma.cuh", "sde_kernel.cuh", "sde_kernel_ws.cuh", "sde_wiener_kernel.cuh",
                 "sde_bead_kernel.cuh"):
        with open(os.path.join(CSRC_DIR, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


_FINGERPRINT = None
_cubins = {}
_programs = {}
STATS = {"nvrtc_compiles": 0, "disk_hits": 0, "memory_hits": 0, "launches": 0}


def source_key(source):
    global _FINGERPRINT
    if _FINGERPRINT is None:
        _FINGERPRINT = _core_fingerprint()
    return hashlib.sha256((_FINGERPRINT + "\n" + source).encode()).hexdigest()[:32]


def compile_source(source, use_cache=True):
    """CUDA text -> cubin bytes (NVRTC, sm_100a).  Cached in memory and under CACHE_DIR."""
    key = source_key(source)
    if use_cache and key in _cubins:
        STATS["memory_hits"] += 1
        return key, _cubins[key]
    path = os.path.join(CACHE_DIR, key + ".cubin")
    if use_cache and os.path.exists(path):
        with open(path, "rb") as f:
            data = f.read()
        _cubins[key] = data
        STATS["disk_hits"] += 1
        return key, data
    L = lib()
    cubin = ctypes.c_void_p()
    size = ctypes.c_size_t()
    log = ctypes.c_void_p()
    st = L.sdeb200_compile(source.encode(), CSRC_DIR.encode(), b"sm_100a", ctypes.byref(cubin), ctypes.byref(size),
                           ctypes.byref(log))
    if log.value:
        L.sdeb200_free(log)
    check(st, "NVRTC compilation of the emitted problem source failed")
    data = ctypes.string_at(cubin.value, size.value)
    L.sdeb200_free(cubin)
    STATS["nvrtc_compiles"] += 1
    _cubins[key] = data
    if use_cache:
        try:
            os.makedirs(CACHE_DIR, exist_ok=True)
            tmp = path + f".tmp{os.getpid()}"
            with open(tmp, "wb") as f:
                f.write(data)
            os.replace(tmp, path)
            with open(os.path.join(CACHE_DIR, key + ".cu"), "w") as f:
                f.write(source)
        except OSError:
            pass
    return key, data


class Program:
    """A cubin loaded on the current device."""

    def __init__(self, key, cubin, entry):
        self.key, self.entry = key, entry
        self._cubin = cubin
        self.handle = ctypes.c_void_p()
        check(lib().sdeb200_program_load(cubin, len(cubin), entry.encode(), ctypes.byref(self.handle)),
              "loading compiled kernel")

    def info(self):
        r, s, t = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(lib().sdeb200_program_info(self.handle, ctypes.byref(r), ctypes.byref(s), ctypes.byref(t)), "program_info")
        return {"registers": r.value, "static_smem": s.value, "max_threads": t.value}


def load_program(source, entry):
    if os.environ.get("SDEB200_PRECOMPILE_ONLY"):
        # cache-warming mode used on the CPU-only build box: compile, then refuse to run
        compile_source(source)
        raise Sdeb200Error("SDEB200_PRECOMPILE_ONLY is set: kernel compiled into the cache, not launched")
    torch = require_cuda()
    dev = torch.cuda.current_device()
    key = source_key(source)
    pk = (key, entry, dev)
    prog = _programs.get(pk)
    if prog is None:
        _, cubin = compile_source(source)
        check(lib().sdeb200_set_device(dev), "set_device")
        prog = Program(key, cubin, entry)
        _programs[pk] = prog
    return prog
