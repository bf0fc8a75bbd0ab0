"""Thin call layer over the C ABI: torch tensors in, raw device pointers out.

PyTorch is used for device memory, streams and autograd bookkeeping only; all
arithmetic of the hot path happens in ``libcsg2im.so``.  Every helper insists
on CUDA tensors -- there is no CPU path.
"""
import torch

from . import _lib

A_ROW, A_COL, A_GATHER = 0, 1, 2
B_NK, B_KN, B_GATHER = 0, 1, 2


def lib():
    return _lib.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("csg2im kernels need CUDA tensors (got a %s tensor); there is no CPU fallback"
                               % t.device.type)


def f32c(t):
    """contiguous fp32 view/copy"""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class KernelTimer:
    """CUDA-event timing of one kernel class on the launching stream (bench.py's roofline leg)."""

    def __init__(self):
        self.records = []      # (work, start_event, end_event)
        self._pool = []

    def events(self):
        if self._pool:
            return self._pool.pop()
        return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def add(self, work, e0, e1):
        self.records.append((work, e0, e1))

    def summary(self):
        """(total work, total seconds, launches); call after a synchronize."""
        work = sum(r[0] for r in self.records)
        sec = sum(r[1].elapsed_time(r[2]) for r in self.records) * 1e-3
        n = len(self.records)
        for r in self.records:
            self._pool.append((r[1], r[2]))
        self.records = []
        return work, sec, n


TIMERS = {}     # kernel class name -> KernelTimer (set by bench.py; empty = no timing overhead)


def workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------------------- GEMM family (fp32)
class Gather:
    """Sources of the fused triple-input gather [obj[s] | pred | obj[o]] (graph.py:63-66)."""

    def __init__(self, obj, pred, s_idx, o_idx, p_idx=None):
        """``p_idx`` (int32 [rows]): ``pred`` is a table and the predicate segment is ``pred[p_idx[t]]`` (and ``obj``
        may then be the object embedding table with ``s_idx`` / ``o_idx`` class ids): model.py:108-109 fused."""
        assert obj.is_contiguous() and pred.stride(-1) == 1
        self.obj, self.pred, self.s_idx, self.o_idx, self.p_idx = obj, pred, s_idx, o_idx, p_idx
        self.din, self.dp, self.ldp = obj.shape[1], pred.shape[1], pred.stride(0)

    @property
    def width(self):
        return 2 * self.din + self.dp


# fp32 engine: "simt" = csg_gemm_f32 (fp32 FMA pipes, the default parity engine), "tc" = the same GEMMs on tcgen05 through
# an exact three-term bf16 split of both operands, K-concatenated (csg_split3_bf16 + csg_gemm_bf16 with fp32 accumulation
# and output).  tcgen05.mma truncates its accumulator at every instruction, so the six products are ordered smallest
# first and the hi*hi products of long reductions run as short chains: as accurate against float64
# as the FMA kernel (relative rms 1e-7 .. 1e-6), 1.8x faster on the cfg2 step.  Process-wide switch.
F32_ENGINE = __import__("os").environ.get("CSG_F32_ENGINE", "simt")


def set_f32_engine(name):
    global F32_ENGINE
    if name not in ("simt", "tc"):
        raise ValueError("fp32 engine must be 'simt' or 'tc'")
    F32_ENGINE = name


def _split3(X, rows, cols, transpose, k_is_cols, role):
    """K-concatenated three-term bf16 split of the fp32 matrix X [rows, cols] (or of its transpose); role 0 / 1: all six
    blocks of an A / B operand, 2 / 3: the hi block alone, 4 / 5: the five correction blocks."""
    R, C = (cols, rows) if transpose else (rows, cols)
    nb = (6, 1, 5)[role >> 1]
    out = torch.empty((R, nb * C) if k_is_cols else (nb * R, C), dtype=torch.bfloat16, device=X.device)
    _lib.check(lib().csg_split3_bf16(ptr(X), rows, cols, X.stride(0), int(transpose), int(k_is_cols), role, ptr(out),
                                     out.stride(0), _stream()), "csg_split3_bf16")
    return out


def _gemm_f32_tc(amode, bmode, M, N, K, A, B, out, bias, relu, rowscale, mask_aux, gather):
    """The fp32 GEMM modes of the parity engine on the tensor cores; returns None for shapes the tcgen05 kernels do not
    take (the caller then runs the SIMT kernel)."""
    if (N % 32 or (amode != A_COL and K % 8) or (amode == A_COL and M % 8) or M == 0 or N == 0 or K == 0
            or (out is not None and (out.stride(0) % 4 or not out.is_contiguous()))):
        return None
    dev = (A if A is not None else B).device
    if amode == A_GATHER or bmode == B_GATHER:
        g = gather
        X = torch.cat([g.obj[g.s_idx.long()], g.pred, g.obj[g.o_idx.long()]], dim=1)      # the reference's own concat
        if amode == A_GATHER:
            A, amode = X, A_ROW
        else:
            B, bmode = X, B_KN
    A, B = f32c(A), f32c(B)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=dev)
    L = lib()
    if amode == A_ROW:
        # C[M, N] = A[M, K] B^T with B as [N, K] (B_NK) or [K, N] (B_KN, transposed while it is split): K-major GEMM
        a6 = _split3(A, M, K, False, True, 0)
        b6 = _split3(B, N, K, False, True, 1) if bmode == B_NK else _split3(B, K, N, True, True, 1)
        rc = L.csg_gemm_bf16(0, 0, M, N, 6 * K, ptr(a6), a6.stride(0), ptr(b6), b6.stride(0), ptr(out), out.stride(0), 1,
                             ptr(bias), int(relu), ptr(rowscale), 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, _stream())
    else:
        # C[M, N] = A[K, M]^T B[K, N]: MN-major GEMM, parts stacked along the rows (= K)
        if bmode != B_KN or bias is not None or relu or rowscale is not None or mask_aux is not None:
            return None
        # The reduction is long here (K = triples): the hi*hi products must not sit in one accumulator for thousands of
        # truncating MMAs.  K is cut into chunks of KC rows; per chunk the hi*hi block is a GEMM of its own, whose split-K
        # leaves chains of <= ~64 MMAs, and the five correction blocks (2^-8 .. 2^-16 of the result) another; the chunk
        # results are added in fp32 (round to nearest).
        KC = 8192
        rc = 0
        first = True
        for k0 in range(0, K, KC):
            k1 = min(K, k0 + KC)
            kc = k1 - k0
            Ak, Bk = A[k0:k1], B[k0:k1]
            for ra, rb, nb in ((4, 5, 5), (2, 3, 1)):
                a6 = _split3(Ak, kc, M, False, False, ra)
                b6 = _split3(Bk, kc, N, False, False, rb)
                dst = out if first else torch.empty_like(out)
                ws = workspace(L.csg_gemm_bf16_workspace(M, N, nb * kc, 1), dev)
                rc = L.csg_gemm_bf16(1, 0, M, N, nb * kc, ptr(a6), a6.stride(0), ptr(b6), b6.stride(0), ptr(dst), dst.stride(0), 1,
                                     0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, ptr(ws), ws.numel(), _stream())
                if rc:
                    break
                if not first:
                    out += dst
                first = False
            if rc:
                break
    _lib.check(rc, "csg_gemm_bf16 (fp32 through three-term split)")
    if mask_aux is not None:
        # ReLU mask of the epilogue as its own exact pass (the tensor-core epilogue takes a 16-bit mask operand)
        m = f32c(mask_aux)
        _lib.check(L.csg_relu_mask_f32(ptr(out), ptr(m), ptr(out), out.numel(), _stream()), "csg_relu_mask_f32")
    return out


def gemm_f32(amode, bmode, M, N, K, A, B, out=None, bias=None, relu=False, rowscale=None, mask_aux=None,
             gather=None, lda=None, ldb=None):
    if F32_ENGINE == "tc" and lda in (None, 0) and ldb in (None, 0):
        r = _gemm_f32_tc(amode, bmode, M, N, K, A, B, out, bias, relu, rowscale, mask_aux, gather)
        if r is not None:
            return r
    dev = (A if A is not None else B).device
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=dev)
    need_cuda(A, B, out, bias, rowscale, mask_aux)
    L = lib()
    ws = None
    nws = L.csg_gemm_f32_workspace(M, N, K, amode) if (amode == A_COL and K > 256) else 0
    if nws:
        ws = workspace(nws, dev)
    g = gather
    timer = TIMERS.get("gemm")
    if timer is not None:
        e0, e1 = timer.events()
        e0.record()
    rc = L.csg_gemm_f32(
        amode, bmode, M, N, K,
        ptr(A), (lda if lda is not None else (A.stride(0) if A is not None else 0)),
        ptr(B), (ldb if ldb is not None else (B.stride(0) if B is not None else 0)),
        ptr(out), out.stride(0),
        ptr(bias), int(relu), ptr(rowscale), ptr(mask_aux), (mask_aux.stride(0) if mask_aux is not None else 0),
        ptr(g.obj) if g else 0, ptr(g.pred) if g else 0, ptr(g.s_idx) if g else 0, ptr(g.o_idx) if g else 0,
        g.din if g else 0, g.dp if g else 0, g.ldp if g else 0,
        ptr(ws), (ws.numel() if ws is not None else 0), _stream())
    _lib.check(rc, "csg_gemm_f32")
    if timer is not None:
        e1.record()
        timer.add(2.0 * M * N * K, e0, e1)
    return out


def colsum_f32(X):
    M, N = X.shape
    out = torch.empty(N, dtype=torch.float32, device=X.device)
    L = lib()
    ws = workspace(L.csg_colsum_f32_workspace(M, N), X.device)
    _lib.check(L.csg_colsum_f32(ptr(X), M, N, X.stride(0), ptr(out), ptr(ws), ws.numel(), _stream()), "csg_colsum_f32")
    return out


def relu_mask_f32(dy, y):
    dy, y = f32c(dy), f32c(y)
    out = torch.empty_like(dy)
    _lib.check(lib().csg_relu_mask_f32(ptr(dy), ptr(y), ptr(out), dy.numel(), _stream()), "csg_relu_mask_f32")
    return out


# ---------------------------------------------------------------------------- GEMM family (bf16 tensor cores)
F16, BF16 = torch.float16, torch.bfloat16


def gemm_bf16(M, N, K, A, B, mn_major=False, out=None, out_f32=False, bias=None, relu=False, rowscale=None,
              mask_aux=None, gather=None, gather_mode=0, out_dtype=None):
    """tcgen05 GEMM.  mn_major=False: C = epi(A[M,K] @ B[N,K]^T); mn_major=True: C = A[K,M]^T @ B[K,N] (fp32 out).
    A and B (or the gathered operand) may independently be fp16 or bf16; a 16-bit C is bf16 unless ``out_dtype`` /
    ``out`` says fp16."""
    dev = (A if A is not None else B).device
    if mn_major:
        out_f32 = True
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32 if out_f32 else (out_dtype or BF16), device=dev)
    need_cuda(A, B, out, bias, rowscale, mask_aux)
    for x in (A, B, mask_aux):
        assert x is None or (x.dtype in (BF16, F16) and x.stride(-1) == 1)
    g = gather
    a_t = g.obj.dtype if (g is not None and gather_mode == 1) else A.dtype
    b_t = g.obj.dtype if (g is not None and gather_mode == 2) else B.dtype
    if g is not None:
        assert g.obj.dtype == g.pred.dtype
    formats = int(a_t == F16) | (int(b_t == F16) << 1) | (int(out.dtype == F16) << 2)
    L = lib()
    ws = None
    if mn_major:
        ws = workspace(L.csg_gemm_bf16_workspace(M, N, K, 1), dev)
    timer = TIMERS.get("gemm")
    if timer is not None:
        e0, e1 = timer.events()
        e0.record()
    rc = L.csg_gemm_bf16(
        int(mn_major), int(gather_mode), M, N, K,
        ptr(A), A.stride(0) if A is not None else 0, ptr(B), B.stride(0) if B is not None else 0,
        ptr(out), out.stride(0), int(out_f32),
        ptr(bias), int(relu), ptr(rowscale), ptr(mask_aux), (mask_aux.stride(0) if mask_aux is not None else 0),
        ptr(g.obj) if g else 0, ptr(g.pred) if g else 0, ptr(g.s_idx) if g else 0, ptr(g.o_idx) if g else 0,
        g.din if g else 0, g.dp if g else 0, g.ldp if g else 0, g.obj.shape[0] if g else 0,
        ptr(g.p_idx) if g else 0, (g.pred.shape[0] if (g and g.p_idx is not None) else 0), formats,
        ptr(ws), (ws.numel() if ws is not None else 0), _stream())
    _lib.check(rc, "csg_gemm_bf16")
    if timer is not None:
        e1.record()
        timer.add(2.0 * M * N * K, e0, e1)
    return out


def embed_table_grad(dout, idx, V, E):
    """dtable [V, E] fp32 = sum over rows r with idx[r] = v of dout[r] (deterministic): the gradient of an embedding
    lookup (attribute_embed.py:38-48, model.py:109).  Large bf16 inputs (the predicate rows of the tensor-core engine)
    go through the one-hot tensor-core product onehot^T dout, everything else through csg_embed_bwd."""
    n = idx.numel()
    if dout.dtype not in (torch.float32, torch.bfloat16):
        dout = dout.float()
    if dout.stride(-1) != 1 or dout.stride(0) % 4 != 0 or dout.data_ptr() % 16 != 0:
        dout = dout.contiguous()
    L = lib()
    if (dout.dtype == torch.bfloat16 and n >= 1024 and E % 32 == 0 and dout.stride(0) % 8 == 0
            and dout.data_ptr() % 16 == 0):
        ld = (V + 63) // 64 * 64
        onehot = torch.empty((n, ld), dtype=torch.bfloat16, device=dout.device)
        _lib.check(L.csg_onehot_bf16(ptr(idx), idx.stride(0), n, V, ptr(onehot), ld, _stream()), "csg_onehot_bf16")
        return gemm_bf16(ld, E, n, onehot, dout, mn_major=True)[:V]
    dtable = torch.empty((V, E), dtype=torch.float32, device=dout.device)
    ws = workspace(L.csg_embed_bwd_workspace(n, V, E), dout.device)
    rc = L.csg_embed_bwd(ptr(dout), dout.stride(0) if n else E, int(dout.dtype == torch.bfloat16), ptr(idx),
                         idx.stride(0) if n else 1, n, V, E, ptr(dtable), ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "csg_embed_bwd")
    return dtable
