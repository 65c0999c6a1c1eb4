"""ctypes binding of libosb200.so (the C ABI in include/osb200.h).

PyTorch is used only for device memory and streams: every op takes torch CUDA tensors, passes
raw device pointers + the current stream to the library and returns the output tensor.  There is
NO fallback: importing without the built library, or calling an op on a non-CUDA tensor, raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OSB200_LIB", os.path.join(_HERE, "libosb200.so"))  # override: instrumented debug builds

# every symbol include/osb200.h declares (tests check the library exports all of them)
EXPORTS = (
    "osb_init",
    "osb_version",
    "osb_last_error",
    "osb_launch_count",
    "osb_ln_modulate",
    "osb_gemm_bf16",
    "osb_attn_short",
    "osb_conv3d_ndhwc",
    "osb_group_stats",
    "osb_group_stats_workspace_bytes",
    "osb_vae_prep",
    "osb_cfg_euler",
    "osb_gemm_head_tiles",
    "osb_head_tiles_per_head",
    "osb_attn_tiles",
    "osb_tmap_cache_stats",
    "osb_ln_modulate_scatter",
    "osb_comm_barrier",
)

EPI_BIAS, EPI_BIAS_GELU_TANH, EPI_BIAS_GATE_RES = 0, 1, 2
ATTN_IMPL = int(os.environ.get("OSB_ATTN_IMPL", "0"))  # experiment switch for osb_attn_short's implementation


class OsbError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise OsbError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "osb200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise OsbError(f"libosb200.so does not export {name}")
    lib.osb_last_error.restype = C.c_char_p
    lib.osb_launch_count.restype = C.c_int64
    lib.osb_init.argtypes = [C.c_int]
    lib.osb_ln_modulate.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p,
        C.c_int64, C.c_float, C.c_void_p,
    ]
    lib.osb_gemm_bf16.argtypes = [C.c_void_p, C.c_void_p]
    lib.osb_attn_short.argtypes = [C.c_void_p, C.c_void_p]
    lib.osb_conv3d_ndhwc.argtypes = [C.c_void_p, C.c_void_p]
    lib.osb_vae_prep.argtypes = [C.c_void_p, C.c_void_p]
    lib.osb_cfg_euler.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float,
                                  C.c_float, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]
    lib.osb_group_stats.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_void_p,
                                    C.c_int64, C.c_void_p, C.c_void_p]
    lib.osb_group_stats_workspace_bytes.argtypes = [C.c_int64, C.c_int64, C.c_int32]
    lib.osb_group_stats_workspace_bytes.restype = C.c_int64
    lib.osb_gemm_head_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.osb_attn_tiles.argtypes = [C.c_void_p, C.c_void_p]
    lib.osb_head_tiles_per_head.argtypes = [C.c_void_p, C.c_int64]
    lib.osb_head_tiles_per_head.restype = C.c_int64
    lib.osb_tmap_cache_stats.argtypes = [C.c_void_p, C.c_void_p]
    lib.osb_tmap_cache_stats.restype = None
    lib.osb_ln_modulate_scatter.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_float,
        C.c_void_p, C.c_void_p,
    ]
    lib.osb_comm_barrier.argtypes = [C.c_void_p, C.c_void_p]
    return lib


_lib = _load()
_initialised_devices: set[int] = set()

# Optional per-launch device timing (bench.py's roofline leg): when a list is installed with
# start_profile(), every op brackets its launch with CUDA events on the launching stream and appends
# (kernel family, algorithmic work, start_event, end_event).  Off (None) in normal operation.
_profile: list | None = None


def start_profile() -> None:
    global _profile
    _profile = []


def stop_profile() -> list:
    """Returns [(family, work, milliseconds)] after synchronising."""
    global _profile
    import torch

    torch.cuda.synchronize()
    rec, _profile = _profile or [], None
    return [(n, w, s.elapsed_time(e)) for (n, w, s, e) in rec]


class _Timed:
    def __init__(self, family: str, work: float):
        self.family, self.work = family, work

    def __enter__(self):
        if _profile is not None:
            import torch

            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if _profile is not None:
            self.e.record()
            _profile.append((self.family, self.work, self.s, self.e))
        return False


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("W", C.c_void_p), ("bias", C.c_void_p), ("D", C.c_void_p),
        ("R", C.c_void_p), ("gate", C.c_void_p), ("mod_index", C.c_void_p),
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
        ("lda", C.c_int64), ("ldw", C.c_int64), ("ldd", C.c_int64), ("ldr", C.c_int64),
        ("group_rows", C.c_int64), ("gate_stride", C.c_int64),
        ("epilogue", C.c_int32), ("cta_group", C.c_int32), ("block_n", C.c_int32), ("reserved", C.c_int32),
    ]


class AttnShortArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("q_ld", C.c_int64), ("k_ld", C.c_int64), ("v_ld", C.c_int64), ("out_ld", C.c_int64),
        ("num_seqs", C.c_int64), ("seqs_per_batch", C.c_int64),
        ("q_batch_stride", C.c_int64), ("q_seq_stride", C.c_int64), ("q_tok_stride", C.c_int64),
        ("k_batch_stride", C.c_int64), ("k_seq_stride", C.c_int64), ("k_tok_stride", C.c_int64),
        ("Lq", C.c_int32), ("Lk", C.c_int32),
        ("kv_lens", C.c_void_p),
        ("num_heads", C.c_int32), ("head_dim", C.c_int32),
        ("q_norm_w", C.c_void_p), ("k_norm_w", C.c_void_p),
        ("norm_eps", C.c_float),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
        ("softmax_scale", C.c_float),
        ("q_norm_w2", C.c_void_p), ("k_norm_w2", C.c_void_p), ("norm_split", C.c_int32), ("reserved", C.c_int32),
        ("rope_half", C.c_int32), ("reserved2", C.c_int32),
    ]


class Conv3dArgs(C.Structure):
    _fields_ = [
        ("x_pad", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("y", C.c_void_p), ("residual", C.c_void_p),
        ("nb", C.c_int32), ("tp", C.c_int32), ("hp", C.c_int32), ("wp", C.c_int32), ("cp", C.c_int32),
        ("t_out", C.c_int32), ("h_out", C.c_int32), ("w_out", C.c_int32), ("cout", C.c_int32),
        ("st", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("kt", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("narrow", C.c_int32), ("block_n", C.c_int32),
    ]


class VaePrepArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("mean_rstd", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("nb", C.c_int32), ("t", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
        ("groups", C.c_int32), ("silu", C.c_int32),
        ("ft", C.c_int32), ("fh", C.c_int32), ("fw", C.c_int32),
        ("pad_t", C.c_int32), ("pad_h", C.c_int32), ("pad_w", C.c_int32),
        ("cp", C.c_int32),
    ]


def last_error() -> str:
    return _lib.osb_last_error().decode()


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise OsbError(f"{what} failed ({rc}): {last_error()}")


def version() -> int:
    return _lib.osb_version()


def launch_count() -> int:
    return int(_lib.osb_launch_count())


def init(device: int | None = None) -> None:
    import torch

    if not torch.cuda.is_available():
        raise OsbError("osb200 needs a CUDA device (sm_100a); there is no CPU path")
    if device is None:
        device = torch.cuda.current_device()
    if device in _initialised_devices:
        return
    torch.cuda.init()
    _check(_lib.osb_init(int(device)), "osb_init")
    _initialised_devices.add(device)


def _stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _need(t, dtype, name):
    import torch

    if t is None:
        return
    if not t.is_cuda:
        raise OsbError(f"{name} must be a CUDA tensor (osb200 has no CPU path)")
    if t.dtype != dtype:
        raise OsbError(f"{name} must be {dtype}, got {t.dtype}")
    if t.dim() >= 1 and t.stride(-1) != 1 and t.shape[-1] != 1:
        raise OsbError(f"{name} must have unit stride in the last dimension")
    init(t.device.index)


def require_cuda_bf16(t, what: str) -> None:
    """The product path has no CPU / eager fallback: the host-side models call this on entry so that a model left on the
    CPU or in another dtype fails loudly instead of silently running somewhere else."""
    import torch

    if not t.is_cuda or t.dtype != torch.bfloat16:
        raise OsbError(f"{what} (osb200) runs on CUDA in bfloat16 only: call .cuda().to(torch.bfloat16); "
                       "there is no CPU / eager fallback")


MAX_PEERS = 16


class Scatter(C.Structure):
    """include/osb200.h `osb_scatter`: output rows viewed as [B, I, J] are routed to peer buffers (sequence parallel)."""
    _fields_ = [("mode", C.c_int32), ("P", C.c_int32), ("rank", C.c_int32), ("I", C.c_int32), ("J", C.c_int32),
                ("reserved", C.c_int32 * 3), ("peer", C.c_void_p * MAX_PEERS)]


class CommBarrierArgs(C.Structure):
    _fields_ = [("P", C.c_int32), ("rank", C.c_int32), ("epoch", C.c_void_p), ("flags_local", C.c_void_p),
                ("flags_peer", C.c_void_p * MAX_PEERS)]


def make_scatter(mode: int, P: int, rank: int, I: int, J: int, peer_ptrs) -> Scatter:
    sc = Scatter()
    sc.mode, sc.P, sc.rank, sc.I, sc.J = mode, P, rank, I, J
    for i, ptr in enumerate(peer_ptrs):   # raw (peer-mapped) device pointers, or local tensors
        sc.peer[i] = int(ptr) if isinstance(ptr, int) else ptr.data_ptr()
    return sc


def comm_barrier(P: int, rank: int, epoch, flags_local_ptr: int, flags_peer_ptrs) -> None:
    """Cross-rank ordering point of a peer-memory exchange (osb_comm_barrier): enqueued on the current stream."""
    import torch

    _need(epoch, torch.int32, "epoch")
    a = CommBarrierArgs()
    a.P, a.rank, a.epoch, a.flags_local = P, rank, epoch.data_ptr(), int(flags_local_ptr)
    for i, ptr in enumerate(flags_peer_ptrs):
        a.flags_peer[i] = int(ptr)
    _check(_lib.osb_comm_barrier(C.byref(a), _stream()), "osb_comm_barrier")


def ln_modulate(x, shift, scale, *, group_rows: int, mod_index=None, eps: float = 1e-6, out=None, scatter: Scatter | None = None):
    """y = LN(x) * (1 + scale[g]) + shift[g];  x bf16 [rows, C]; shift/scale fp32 [G, C] views.  With `scatter` the rows
    are stored straight into peer buffers (osb_ln_modulate_scatter) and nothing is returned."""
    import torch

    _need(x, torch.bfloat16, "x"); _need(shift, torch.float32, "shift"); _need(scale, torch.float32, "scale")
    _need(mod_index, torch.int32, "mod_index")
    assert x.dim() == 2 and x.is_contiguous()
    assert shift.dim() == 2 and scale.dim() == 2 and shift.stride(0) == scale.stride(0)
    rows, Cdim = x.shape
    if scatter is not None:
        with _Timed("ln_modulate", 4.0 * rows * Cdim):
            _check(_lib.osb_ln_modulate_scatter(_ptr(x), _ptr(shift), _ptr(scale), rows, Cdim, group_rows, _ptr(mod_index),
                                                shift.stride(0), eps, C.byref(scatter), _stream()), "osb_ln_modulate_scatter")
        return None
    if out is None:
        out = torch.empty_like(x)
    with _Timed("ln_modulate", 4.0 * rows * Cdim):  # algorithmic bytes: read x + write y (bf16)
        _check(_lib.osb_ln_modulate(_ptr(x), _ptr(shift), _ptr(scale), _ptr(out), rows, Cdim, group_rows,
                                    _ptr(mod_index), shift.stride(0), eps, _stream()), "osb_ln_modulate")
    return out


def gemm(a, w, bias=None, *, epilogue: int = EPI_BIAS, residual=None, gate=None, group_rows: int = 0,
         mod_index=None, out=None, cta_group: int = 0, block_n: int = 0):
    """out = epilogue(a @ w.T + bias).  a bf16 [M,K] (row stride free), w bf16 [N,K], bias bf16 [N].
    GATE_RES: out = residual + gate[g] * (a @ w.T + bias), gate fp32 [G, N] view, may be None."""
    import torch

    _need(a, torch.bfloat16, "a"); _need(w, torch.bfloat16, "w"); _need(bias, torch.bfloat16, "bias")
    _need(residual, torch.bfloat16, "residual"); _need(gate, torch.float32, "gate")
    _need(mod_index, torch.int32, "mod_index")
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _need(out, torch.bfloat16, "out")
    args = GemmArgs()
    args.A, args.W, args.bias, args.D = a.data_ptr(), w.data_ptr(), (bias.data_ptr() if bias is not None else None), out.data_ptr()
    args.R = residual.data_ptr() if residual is not None else None
    args.gate = gate.data_ptr() if gate is not None else None
    args.mod_index = mod_index.data_ptr() if mod_index is not None else None
    args.M, args.N, args.K = M, N, K
    args.lda, args.ldw, args.ldd = a.stride(0), w.stride(0), out.stride(0)
    args.ldr = residual.stride(0) if residual is not None else 0
    args.group_rows = group_rows if group_rows > 0 else M
    args.gate_stride = gate.stride(0) if gate is not None else 0
    args.epilogue, args.cta_group, args.block_n = epilogue, cta_group, block_n
    with _Timed("gemm", 2.0 * M * N * K):  # algorithmic FLOPs
        _check(_lib.osb_gemm_bf16(C.byref(args), _stream()), "osb_gemm_bf16")
    return out


def attn_short(q, k, v, out, *, num_seqs: int, seqs_per_batch: int, q_strides, k_strides, Lq: int, Lk: int,
               num_heads: int, head_dim: int, kv_lens=None, q_norm_w=None, k_norm_w=None, norm_eps: float = 1e-6,
               rope_cos=None, rope_sin=None, softmax_scale: float | None = None, q_norm_w2=None, k_norm_w2=None,
               norm_split: int = 0, impl: int = 0, rope_half: bool = False):
    """softmax(q k^T * scale) v per (sequence, head) with optional fused QK-RMSNorm and RoPE.
    q/k/v/out are 2-D bf16 views [rows, ld]; *_strides = (batch, seq, token) strides in rows."""
    import torch

    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out"), (q_norm_w, "q_norm_w"), (k_norm_w, "k_norm_w")):
        _need(t, torch.bfloat16, n)
    _need(rope_cos, torch.float32, "rope_cos"); _need(rope_sin, torch.float32, "rope_sin")
    _need(kv_lens, torch.int32, "kv_lens")
    a = AttnShortArgs()
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.q_ld, a.k_ld, a.v_ld, a.out_ld = q.stride(0), k.stride(0), v.stride(0), out.stride(0)
    a.num_seqs, a.seqs_per_batch = num_seqs, seqs_per_batch
    a.q_batch_stride, a.q_seq_stride, a.q_tok_stride = q_strides
    a.k_batch_stride, a.k_seq_stride, a.k_tok_stride = k_strides
    a.Lq, a.Lk = Lq, Lk
    a.kv_lens = kv_lens.data_ptr() if kv_lens is not None else None
    a.num_heads, a.head_dim = num_heads, head_dim
    a.q_norm_w = q_norm_w.data_ptr() if q_norm_w is not None else None
    a.k_norm_w = k_norm_w.data_ptr() if k_norm_w is not None else None
    a.norm_eps = norm_eps
    a.rope_cos = rope_cos.data_ptr() if rope_cos is not None else None
    a.rope_sin = rope_sin.data_ptr() if rope_sin is not None else None
    a.softmax_scale = softmax_scale if softmax_scale is not None else head_dim ** -0.5
    _need(q_norm_w2, torch.bfloat16, "q_norm_w2"); _need(k_norm_w2, torch.bfloat16, "k_norm_w2")
    a.q_norm_w2 = q_norm_w2.data_ptr() if q_norm_w2 is not None else None
    a.k_norm_w2 = k_norm_w2.data_ptr() if k_norm_w2 is not None else None
    a.norm_split = norm_split
    a.rope_half = int(rope_half)
    a.reserved = impl if impl else ATTN_IMPL  # 0 = library default; 1 resident keys, 2 flash (P via smem), 3 flash (P in TMEM), 4 ping-pong
    with _Timed("attn_short", 4.0 * num_seqs * Lq * Lk * num_heads * head_dim):  # QK^T + PV FLOPs
        _check(_lib.osb_attn_short(C.byref(a), _stream()), "osb_attn_short")
    return out


# ---- head tiles: projection GEMM -> attention without a layout pass (include/osb200.h) ------------------------------
class TileMap(C.Structure):
    _fields_ = [("mode", C.c_int32), ("L", C.c_int32), ("S", C.c_int32), ("T", C.c_int32), ("G", C.c_int32),
                ("tps", C.c_int32), ("tile_rows", C.c_int32), ("reserved", C.c_int32)]

    def key(self):
        return (self.mode, self.L, self.S, self.T, self.G, self.tps, self.tile_rows)


class HeadTilesArgs(C.Structure):
    _fields_ = [
        ("tiles", C.c_void_p), ("kind_stride", C.c_int64), ("head_stride", C.c_int64), ("map", TileMap),
        ("num_heads", C.c_int32), ("head_dim", C.c_int32), ("nkinds", C.c_int32),
        ("norm_mask", C.c_uint32), ("rope_mask", C.c_uint32), ("reserved", C.c_int32),
        ("norm_w", C.c_void_p * 4), ("norm_eps", C.c_float), ("reserved2", C.c_int32),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
    ]


class AttnTilesArgs(C.Structure):
    _fields_ = [
        ("q_tiles", C.c_void_p), ("k_tiles", C.c_void_p), ("v_tiles", C.c_void_p),
        ("q_head_stride", C.c_int64), ("kv_head_stride", C.c_int64), ("q_map", TileMap),
        ("kv_tile_rows", C.c_int32), ("kv_tiles_per_set", C.c_int32), ("Lk", C.c_int32),
        ("num_heads", C.c_int32), ("head_dim", C.c_int32), ("reserved", C.c_int32),
        ("num_seqs", C.c_int64), ("kv_lens", C.c_void_p), ("out", C.c_void_p), ("out_ld", C.c_int64),
        ("softmax_scale", C.c_float), ("reserved2", C.c_int32), ("out_scatter", C.c_void_p),
    ]


def tile_map(mode: int, L: int, S: int = 0, T: int = 0, *, keys_only: bool = False, pack: bool = True) -> TileMap:
    """Token row -> (tile, row) map of include/osb200.h `osb_tile_map`.  mode 0: sequences are contiguous blocks of L
    rows; mode 1: sequences run along T of a frame-major [B, T, S] token stream (L == T).  Short sequences (L <= 64) are
    packed 128 // L per tile (self-attention: a tile is its own key set; `pack=False` for cross-attention queries,
    whose key set is per sequence); `keys_only` (text keys of cross-attention) never packs either."""
    m = TileMap()
    m.mode, m.L, m.S, m.T = mode, L, S, T
    if L <= 64 and pack and not keys_only:
        m.G, m.tps = 128 // L, 1
        m.tile_rows = -(-(m.G * L) // 16) * 16
    else:
        m.G = 1
        n = -(-L // 128)
        # (keys-only tiles used to be balanced, 300 -> 3 x 112; a 112-row tile ends in the middle of a 32-column softmax
        # chunk and sent a quarter of the chunks through the per-element masked path: full 128-row tiles + a short last one)
        m.tile_rows = 128 if L > 128 else -(-(-(-L // n)) // 16) * 16
        m.tps = -(-L // m.tile_rows)
    return m


class HeadTiles:
    """A buffer of head tiles: `kinds` column groups (q | k | v ...) x heads x tiles.  Zero-initialised: rows no token
    maps to (ragged last tile, head-dim tail) must stay finite, they are multiplied by P = 0 in the PV product."""

    def __init__(self, rows: int, tmap: TileMap, kinds: int, heads: int, head_dim: int, device):
        import torch

        self.rows, self.map, self.kinds, self.heads, self.head_dim = rows, tmap, kinds, heads, head_dim
        self.tiles_per_head = int(_lib.osb_head_tiles_per_head(C.byref(tmap), rows))
        if self.tiles_per_head <= 0:
            raise OsbError(f"tile map {tmap.key()} does not fit {rows} rows")
        self.tile_bytes = tmap.tile_rows * (-(-head_dim // 16) * 16) * 2
        self.head_stride = self.tiles_per_head * self.tile_bytes
        self.kind_stride = heads * self.head_stride
        self.buf = torch.zeros(kinds * self.kind_stride, dtype=torch.uint8, device=device)

    def kind_ptr(self, kind: int) -> int:
        return self.buf.data_ptr() + kind * self.kind_stride


def gemm_head_tiles(a, w, bias, tiles: HeadTiles, *, nkinds: int, norm_w=(), rope=None, rope_kinds: int = 0,
                    eps: float = 1e-6, kind0: int = 0, general: bool = False):
    """tiles[kind0 + n // C] = head_tiles(a @ w.T + bias): each output row is split into heads; kinds listed in
    `norm_w` (bf16 [D] or None per kind) get per-head RMSNorm, kinds in the `rope_kinds` bit mask get interleaved-pair
    RoPE by token position from `rope` = (cos, sin) fp32 [L, D/2]; one rounding to bf16 at the row's place in its tile."""
    import torch

    _need(a, torch.bfloat16, "a"); _need(w, torch.bfloat16, "w"); _need(bias, torch.bfloat16, "bias")
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    Cc = tiles.heads * tiles.head_dim
    assert M == tiles.rows and N % Cc == 0 and kind0 + N // Cc <= tiles.kinds
    g = GemmArgs()
    g.A, g.W, g.bias = a.data_ptr(), w.data_ptr(), (bias.data_ptr() if bias is not None else None)
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldw = a.stride(0), w.stride(0)
    t = HeadTilesArgs()
    t.tiles = tiles.kind_ptr(kind0)
    t.kind_stride, t.head_stride, t.map = tiles.kind_stride, tiles.head_stride, tiles.map
    t.num_heads, t.head_dim, t.nkinds = tiles.heads, tiles.head_dim, nkinds
    t.reserved = 1 if general else 0   # force the general per-row-store epilogue (tests / A-B)
    mask = 0
    for i, nw in enumerate(norm_w):
        if nw is not None:
            _need(nw, torch.bfloat16, "norm_w")
            t.norm_w[i] = nw.data_ptr()
            mask |= 1 << i
    t.norm_mask, t.rope_mask, t.norm_eps = mask, (rope_kinds if rope is not None else 0), eps
    if rope is not None:
        _need(rope[0], torch.float32, "rope cos"); _need(rope[1], torch.float32, "rope sin")
        t.rope_cos, t.rope_sin = rope[0].data_ptr(), rope[1].data_ptr()
    with _Timed("gemm", 2.0 * M * N * K):
        _check(_lib.osb_gemm_head_tiles(C.byref(g), C.byref(t), _stream()), "osb_gemm_head_tiles")
    return tiles


def attn_tiles(q: HeadTiles, kv: HeadTiles, out, *, q_kind: int = 0, k_kind: int = 1, v_kind: int = 2, Lk: int,
               num_seqs: int, kv_lens=None, softmax_scale: float | None = None, out_scatter: Scatter | None = None,
               out_ld: int | None = None, out_map: TileMap | None = None):
    """out = softmax(q k^T * scale) v per (sequence, head) over head tiles (osb_attn_tiles).  Self-attention: q and kv
    are the same buffer (kinds 0, 1, 2); cross-attention: kv holds the text keys / values (`keys_only` map).
    `out_map`: the token order of `out` when it differs from the order the q tiles were written from (temporal attention:
    tiles from the transposed [B, S, T] stream (mode 0), output rows frame-major (mode 1))."""
    import torch

    _need(out, torch.bfloat16, "out"); _need(kv_lens, torch.int32, "kv_lens")
    assert (out is None) != (out_scatter is None), "exactly one of out / out_scatter"
    a = AttnTilesArgs()
    a.q_tiles, a.k_tiles, a.v_tiles = q.kind_ptr(q_kind), kv.kind_ptr(k_kind), kv.kind_ptr(v_kind)
    if out_map is not None:   # output rows in another order than the rows the tiles were written from (same tiling)
        assert out_map.key()[4:] == q.map.key()[4:] and out_map.L == q.map.L
    a.q_head_stride, a.kv_head_stride, a.q_map = q.head_stride, kv.head_stride, (out_map if out_map is not None else q.map)
    a.kv_tile_rows = kv.map.tile_rows
    a.kv_tiles_per_set = kv.map.tps
    a.Lk, a.num_heads, a.head_dim = Lk, q.heads, q.head_dim
    a.num_seqs = num_seqs
    a.kv_lens = kv_lens.data_ptr() if kv_lens is not None else None
    if out_scatter is not None:   # rows go to peer buffers (row stride out_ld elements on every destination)
        a.out, a.out_ld = None, int(out_ld)
        a.out_scatter = C.addressof(out_scatter)
    else:
        a.out, a.out_ld = out.data_ptr(), out.stride(0)
    a.softmax_scale = softmax_scale if softmax_scale is not None else q.head_dim ** -0.5
    with _Timed("attn_tiles", 4.0 * num_seqs * q.map.L * Lk * q.heads * q.head_dim):
        _check(_lib.osb_attn_tiles(C.byref(a), _stream()), "osb_attn_tiles")
    return out


def tmap_cache_stats():
    h, m = C.c_int64(0), C.c_int64(0)
    _lib.osb_tmap_cache_stats(C.byref(h), C.byref(m))
    return int(h.value), int(m.value)


# ---- causal 3D VAE ops (NDHWC) -------------------------------------------------------------------------
def group_stats(x, groups: int, eps: float = 1e-6):
    """GroupNorm statistics of x bf16 [nb, T, H, W, C] (channels last) -> fp32 [nb, groups, 2] = (mean, rstd)."""
    import torch

    _need(x, torch.bfloat16, "x")
    assert x.dim() == 5 and x.is_contiguous()
    nb, Cc = x.shape[0], x.shape[-1]
    pos = x.shape[1] * x.shape[2] * x.shape[3]
    ws_bytes = int(_lib.osb_group_stats_workspace_bytes(nb, pos, groups))
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=x.device)
    out = torch.empty(nb, groups, 2, dtype=torch.float32, device=x.device)
    with _Timed("group_stats", 2.0 * x.numel()):  # algorithmic bytes: one read
        _check(_lib.osb_group_stats(_ptr(x), nb, pos, Cc, groups, eps, _ptr(ws), ws_bytes, _ptr(out), _stream()),
               "osb_group_stats")
    return out


def vae_prep(x, *, stats=None, gamma=None, beta=None, groups: int = 32, silu: bool = False, up=(1, 1, 1),
             pad=(0, 0, 0), cp: int | None = None, slack_bytes: int = 128):
    """[GroupNorm-apply + SiLU] + nearest upsample + replicate pad of x bf16 [nb,T,H,W,C] -> padded bf16
    [nb,Tp,Hp,Wp,cp] (one pass).  pad = (front frames, rows each side, cols each side)."""
    import torch

    _need(x, torch.bfloat16, "x"); _need(stats, torch.float32, "stats")
    _need(gamma, torch.bfloat16, "gamma"); _need(beta, torch.bfloat16, "beta")
    assert x.dim() == 5 and x.is_contiguous()
    nb, T, H, W, Cc = x.shape
    cp = cp or Cc
    tu = T if up[0] == 1 else 1 + up[0] * (T - 1)
    tp, hp, wp = tu + pad[0], H * up[1] + 2 * pad[1], W * up[2] + 2 * pad[2]
    n = nb * tp * hp * wp * cp
    buf = torch.empty(n + slack_bytes // 2, dtype=torch.bfloat16, device=x.device)  # slack: narrow-mode windows
    if slack_bytes:
        buf[n:].zero_()
    y = buf[:n].view(nb, tp, hp, wp, cp)
    a = VaePrepArgs()
    a.x, a.y = x.data_ptr(), y.data_ptr()
    a.mean_rstd = stats.data_ptr() if stats is not None else None
    a.gamma = gamma.data_ptr() if gamma is not None else None
    a.beta = beta.data_ptr() if beta is not None else None
    a.nb, a.t, a.h, a.w, a.c = nb, T, H, W, Cc
    a.groups, a.silu = groups, int(silu)
    a.ft, a.fh, a.fw = up
    a.pad_t, a.pad_h, a.pad_w = pad
    a.cp = cp
    with _Timed("vae_prep", 2.0 * (x.numel() + n)):
        _check(_lib.osb_vae_prep(C.byref(a), _stream()), "osb_vae_prep")
    return y


def conv3d(x_pad, w_packed, bias, *, out_thw, stride=(1, 1, 1), taps=(3, 3, 3), narrow: bool = False, residual=None,
           block_n: int = 0):
    """y = conv3d(x_pad) + bias (+ residual): x_pad bf16 [nb,Tp,Hp,Wp,Cp] (already padded), w_packed bf16 [Cout, K]
    (see pack_conv_weight), y bf16 [nb, T_out, H_out, W_out, Cout]."""
    import torch

    _need(x_pad, torch.bfloat16, "x_pad"); _need(w_packed, torch.bfloat16, "w_packed"); _need(bias, torch.bfloat16, "bias")
    _need(residual, torch.bfloat16, "residual")
    nb, tp, hp, wp, cp = x_pad.shape
    cout = w_packed.shape[0]
    y = torch.empty(nb, *out_thw, cout, dtype=torch.bfloat16, device=x_pad.device)
    a = Conv3dArgs()
    a.x_pad, a.w, a.y = x_pad.data_ptr(), w_packed.data_ptr(), y.data_ptr()
    a.bias = bias.data_ptr() if bias is not None else None
    a.residual = residual.data_ptr() if residual is not None else None
    a.nb, a.tp, a.hp, a.wp, a.cp = nb, tp, hp, wp, cp
    a.t_out, a.h_out, a.w_out = out_thw
    a.cout = cout
    a.st, a.sh, a.sw = stride
    a.kt, a.kh, a.kw = taps
    a.narrow, a.block_n = int(narrow), block_n
    with _Timed("conv3d", 2.0 * y.numel() * w_packed.shape[1]):  # MACs incl. K padding (algorithmic count is the caller's)
        _check(_lib.osb_conv3d_ndhwc(C.byref(a), _stream()), "osb_conv3d_ndhwc")
    return y


def pack_conv_weight(w, cp: int, narrow: bool, cout_pad: int | None = None):
    """torch Conv3d weight [Cout, Cin, kt, kh, kw] -> bf16 [Cout_p, K] K-major in the order the kernel walks K
    (include/osb200.h osb_conv3d_args.w).  Done once at load time."""
    import torch

    cout, cin, kt, kh, kw = w.shape
    co = cout_pad or cout
    if narrow:
        out = torch.zeros(co, kt * kh, 64, dtype=w.dtype, device=w.device)
        blk = torch.zeros(cout, kt * kh, kw, cp, dtype=w.dtype, device=w.device)
        blk[..., :cin] = w.permute(0, 2, 3, 4, 1).reshape(cout, kt * kh, kw, cin)
        out[:cout, :, : kw * cp] = blk.reshape(cout, kt * kh, kw * cp)
        return out.reshape(co, kt * kh * 64).to(torch.bfloat16).contiguous()
    out = torch.zeros(co, kt, kh, kw, cp, dtype=w.dtype, device=w.device)
    out[:cout, ..., :cin] = w.permute(0, 2, 3, 4, 1)
    return out.reshape(co, kt * kh * kw * cp).to(torch.bfloat16).contiguous()


def cfg_euler(cond, uncond, uncond2, x, *, g_txt: float, g_img: float = 1.0, g_img_map=None, dt: float, out=None):
    """out = x + dt * (uncond2 + g_img*(uncond - uncond2) + g_txt*(cond - uncond)); bf16 tensors of one shape."""
    import torch

    for t, n in ((cond, "cond"), (uncond, "uncond"), (uncond2, "uncond2"), (x, "x"), (g_img_map, "g_img_map")):
        _need(t, torch.bfloat16, n)
        assert t is None or t.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    n = x.numel()
    with _Timed("cfg_euler", 2.0 * n * (5 if uncond2 is not None else 4)):
        _check(_lib.osb_cfg_euler(_ptr(cond), _ptr(uncond), _ptr(uncond2), _ptr(x), _ptr(out), n, g_txt, g_img,
                                  _ptr(g_img_map), g_img_map.numel() if g_img_map is not None else 0, dt, _stream()),
               "osb_cfg_euler")
    return out
