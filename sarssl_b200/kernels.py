"""One Python function per C-ABI kernel entry (see include/sarssl_b200.h).  `KernelSet` binds device, stream, activation
storage dtype and the shared reduction workspace so the engine's schedule reads like the math."""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_SWISH, GemmArgs, check, lib, ptr  # noqa: F401


def _addr(t, elem_off=0):
    return t.data_ptr() + elem_off * t.element_size()


class KernelSet:
    def __init__(self, device, dtype):
        self.dev = torch.device(device)
        self.dtype = dtype                       # torch.float32 | torch.bfloat16 (activation storage)
        self.dt = _lib.dtype_code(dtype)
        self.L = lib()
        self.launches = 0
        self.tc_launches = 0
        self.use_tc = True
        self.seed_dev = None                     # device uint64 added to every dropout seed inside the kernels (CUDA-graph replay: engine.py)
        nbytes = max(self.L.sarssl_reduce_workspace_bytes(4096), self.L.sarssl_stem_workspace_bytes(),
                     self.L.sarssl_dwconv_wgrad_workspace_bytes(512, 31), self.L.sarssl_conv3x3_wgrad_tc_workspace_bytes())
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)

    # ------------------------------------------------------------------ helpers
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    @property
    def _sd(self):
        return C.c_void_p(self.seed_dev.data_ptr()) if self.seed_dev is not None else None

    def empty(self, *shape, dtype=None):
        return torch.empty(shape, dtype=dtype or self.dtype, device=self.dev)

    def zeros_f32(self, *shape):
        t = torch.empty(shape, dtype=torch.float32, device=self.dev)
        self.fill(t, 0.0)
        return t

    def _ok(self, rc, name):
        self.launches += 1
        check(rc, name)

    # ------------------------------------------------------------------ GEMM
    def gemm(self, A, B, Cmat, M, N, K, sA, sB, ldc, *, a_off=0, b_off=0, c_off=0, batch=(1, 1), sAb=(0, 0), sBb=(0, 0), sCb=(0, 0),
             bias=None, act=ACT_NONE, pre=None, resid=None, ldr=0, alpha=1.0, beta=1.0, accumulate=False, drop=(0.0, 0), a_drop=(0.0, 0)):
        """C[z][m][n] = resid + beta*drop(act(alpha * sum_k A[z][m][k] B[z][n][k] + bias)); sA = (sAm, sAk), sB = (sBn, sBk)."""
        if A.dtype != B.dtype:
            raise _lib.SarsslError(f"gemm operands differ in dtype: {A.dtype} vs {B.dtype}")
        g = GemmArgs()
        g.A, g.B, g.C = _addr(A, a_off), _addr(B, b_off), _addr(Cmat, c_off)
        g.pre_out = _addr(pre, c_off) if pre is not None else None
        g.resid = _addr(resid) if resid is not None else None
        g.bias = bias.data_ptr() if bias is not None else None
        g.sAm, g.sAk, g.sAb1, g.sAb2 = sA[0], sA[1], sAb[0], sAb[1]
        g.sBn, g.sBk, g.sBb1, g.sBb2 = sB[0], sB[1], sBb[0], sBb[1]
        g.ldc, g.sCb1, g.sCb2, g.ldr = ldc, sCb[0], sCb[1], ldr
        g.M, g.N, g.K, g.nb1, g.nb2 = M, N, K, batch[0], batch[1]
        g.ab_dtype, g.c_dtype, g.act, g.accumulate = _lib.dtype_code(A), _lib.dtype_code(Cmat), act, int(accumulate)
        g.alpha, g.beta = alpha, beta
        g.drop_p, g.drop_seed = drop
        g.a_drop_p, g.a_drop_seed = a_drop
        g.seed_dev = self.seed_dev.data_ptr() if (self.seed_dev is not None and (drop[0] > 0 or a_drop[0] > 0)) else None
        if self.use_tc and A.dtype == torch.bfloat16 and a_drop[0] == 0.0:
            rc = self.L.sarssl_gemm_tc(C.byref(g), self.stream)      # tcgen05 path; -3 = shape/stride it does not take
            if rc == 0:
                self.launches += 1
                self.tc_launches += 1
                return
            if rc != -3:
                check(rc, "sarssl_gemm_tc")
        self._ok(self.L.sarssl_gemm(C.byref(g), self.stream), "sarssl_gemm")

    def linear(self, X, W, Y, M, N, K, **kw):
        """Y[M,N] = X[M,K] @ W[N,K]^T (+ epilogue): nn.Linear / pointwise Conv1d forward."""
        self.gemm(X, W, Y, M, N, K, (K, 1), (K, 1), N, **kw)

    def linear_dgrad(self, dY, W, dX, M, N, K, **kw):
        """dX[M,K] = dY[M,N] @ W[N,K]."""
        self.gemm(dY, W, dX, M, K, N, (N, 1), (1, K), K, **kw)

    def linear_wgrad(self, dY, X, dW, M, N, K, **kw):
        """dW[N,K] += dY[M,N]^T @ X[M,K]  (fp32 accumulation buffer)."""
        self.gemm(dY, X, dW, N, K, M, (1, N), (1, K), K, accumulate=True, **kw)

    # ------------------------------------------------------------------ norms / reductions
    def layernorm_fwd(self, x, ldx, gamma, beta, out, ldo, mean, rstd, rows, cols, out_off=0):
        self._ok(self.L.sarssl_layernorm_fwd(ptr(x), ldx, ptr(gamma), ptr(beta), _addr(out, out_off), ldo, ptr(mean), ptr(rstd), rows, cols, 1e-5,
                                             self.dt, self.stream), "layernorm_fwd")

    def layernorm_bwd(self, dy, lddy, x, ldx, mean, rstd, gamma, add, dx, dgamma, dbeta, rows, cols, dy_off=0):
        self._ok(self.L.sarssl_layernorm_bwd(_addr(dy, dy_off), lddy, ptr(x), ldx, ptr(mean), ptr(rstd), ptr(gamma), ptr(add), ptr(dx), ptr(dgamma),
                                             ptr(dbeta), rows, cols, self.dt, ptr(self.ws), self.ws.numel(), self.stream), "layernorm_bwd")
        self.launches += 1

    def colsum(self, x, ldx, out, rows, cols, accumulate=True, x_off=0):
        self._ok(self.L.sarssl_colsum(_addr(x, x_off), ldx, ptr(out), rows, cols, _lib.dtype_code(x), int(accumulate), ptr(self.ws), self.ws.numel(),
                                      self.stream), "colsum")
        self.launches += 1

    def bn_stats(self, y, rows, Cn, gamma, beta, rmean, rvar, nbt, training):
        stats = torch.empty(4 * Cn, dtype=torch.float32, device=self.dev)
        self._ok(self.L.sarssl_batchnorm_stats(ptr(y), rows, Cn, ptr(gamma), ptr(beta), 1e-5, 0.1, ptr(rmean), ptr(rvar), ptr(nbt), ptr(stats),
                                               int(training), self.dt, ptr(self.ws), self.ws.numel(), self.stream), "batchnorm_stats")
        self.launches += int(training)
        return stats

    def bn_act_fwd(self, y, stats, act, z, rows, Cn):
        self._ok(self.L.sarssl_batchnorm_act_fwd(ptr(y), ptr(stats), act, ptr(z), rows, Cn, self.dt, self.stream), "batchnorm_act_fwd")

    def bn_act_bwd(self, dz, y, stats, act, dy, dgamma, dbeta, rows, Cn):
        self._ok(self.L.sarssl_batchnorm_act_bwd(ptr(dz), ptr(y), ptr(stats), act, ptr(dy), ptr(dgamma), ptr(dbeta), rows, Cn, self.dt,
                                                 ptr(self.ws), self.ws.numel(), self.stream), "batchnorm_act_bwd")
        self.launches += 2

    # ------------------------------------------------------------------ pointwise
    def swish_bwd(self, ds, u, du, n, drop):
        self._ok(self.L.sarssl_swish_bwd(ptr(ds), ptr(u), ptr(du), n, drop[0], drop[1], self._sd, self.dt, self.stream), "swish_bwd")

    def scale_dropout(self, src, dst, n, alpha, drop):
        self._ok(self.L.sarssl_scale_dropout(ptr(src), ptr(dst), n, alpha, drop[0], drop[1], self._sd, self.dt, self.stream), "scale_dropout")

    def relu_bwd(self, dz, z, dy, n):
        self._ok(self.L.sarssl_relu_bwd(ptr(dz), ptr(z), ptr(dy), n, self.dt, self.stream), "relu_bwd")

    def glu_fwd(self, g, a, rows, D):
        self._ok(self.L.sarssl_glu_fwd(ptr(g), ptr(a), rows, D, self.dt, self.stream), "glu_fwd")

    def glu_bwd(self, da, g, dg, rows, D):
        self._ok(self.L.sarssl_glu_bwd(ptr(da), ptr(g), ptr(dg), rows, D, self.dt, self.stream), "glu_bwd")

    def add_head_bias(self, q, ld, u, v, qu, qv, rows, D):
        self._ok(self.L.sarssl_add_head_bias(ptr(q), ld, ptr(u), ptr(v), ptr(qu), ptr(qv), rows, D, self.dt, self.stream), "add_head_bias")

    def add2(self, a, lda, b, ldb, out, ldo, rows, cols, out_off=0):
        self._ok(self.L.sarssl_add2(ptr(a), lda, ptr(b), ldb, _addr(out, out_off), ldo, rows, cols, self.dt, self.stream), "add2")

    def attn_softmax_fwd(self, content, pos, prob, attn, B, H, T, scale, drop):
        self._ok(self.L.sarssl_attn_softmax_fwd(ptr(content), ptr(pos), ptr(prob), ptr(attn), B, H, T, scale, drop[0], drop[1], self._sd, self.dt,
                                                self.stream), "attn_softmax_fwd")

    def attn_softmax_bwd(self, dattn, prob, dpos, B, H, T, scale, drop):
        self._ok(self.L.sarssl_attn_softmax_bwd(ptr(dattn), ptr(prob), ptr(dpos), B, H, T, scale, drop[0], drop[1], self._sd, self.dt, self.stream),
                 "attn_softmax_bwd")
        self.launches += 1

    def dwconv(self, x, w, out, B, T, D, K, flip):
        self._ok(self.L.sarssl_dwconv(ptr(x), ptr(w), ptr(out), B, T, D, K, int(flip), self.dt, self.stream), "dwconv")

    def dwconv_wgrad(self, a, dc, dw, B, T, D, K):
        self._ok(self.L.sarssl_dwconv_wgrad(ptr(a), ptr(dc), ptr(dw), B, T, D, K, self.dt, ptr(self.ws), self.ws.numel(), self.stream), "dwconv_wgrad")
        self.launches += 1

    def mean_pool_fwd(self, x, ldx, pooled, B, T, D, x_off=0):
        self._ok(self.L.sarssl_mean_pool_fwd(_addr(x, x_off), ldx, ptr(pooled), B, T, D, _lib.dtype_code(x), self.stream), "mean_pool_fwd")

    def mean_pool_bwd(self, dpooled, dx, ldx, B, T, D, dx_off=0):
        self._ok(self.L.sarssl_mean_pool_bwd(ptr(dpooled), _addr(dx, dx_off), ldx, B, T, D, _lib.dtype_code(dx), self.stream), "mean_pool_bwd")

    def cast(self, src, dst, n):
        self._ok(self.L.sarssl_cast(ptr(src), _lib.dtype_code(src), ptr(dst), _lib.dtype_code(dst), n, self.stream), "cast")

    def permute4(self, src, dst, dims, strides, accumulate=False, src_off=0):
        d = (C.c_int * 4)(*dims)
        s = (C.c_longlong * 4)(*strides)
        self._ok(self.L.sarssl_permute4(_addr(src, src_off), _lib.dtype_code(src), ptr(dst), _lib.dtype_code(dst), d, s, int(accumulate), self.stream),
                 "permute4")

    def fill(self, t, value):
        assert t.dtype == torch.float32 and t.is_contiguous()
        self._ok(self.L.sarssl_fill_f32(ptr(t), float(value), t.numel(), self.stream), "fill_f32")

    # ------------------------------------------------------------------ stem
    def stem_expand(self, x, mode, flag, ch, w64x4, out, P, W, H):
        self._ok(self.L.sarssl_stem_expand(ptr(x), mode, ptr(flag), ptr(ch), ptr(w64x4), ptr(out), P, W, H, self.dt, self.stream), "stem_expand")

    def stem_expand_bn_relu(self, x, mode, flag, ch, w64x4, stats, out, P, W, H):
        self._ok(self.L.sarssl_stem_expand_bn_relu(ptr(x), mode, ptr(flag), ptr(ch), ptr(w64x4), _addr(stats, 2 * 64), _addr(stats, 3 * 64), ptr(out), P, W, H,
                                                   self.dt, self.stream), "stem_expand_bn_relu")

    def stem_input_bn_stats(self, x, mode, flag, ch, w64x4, P, W, H, bn):
        """Batch statistics (and running-stat update) of BatchNorm(conv1x1(x)) from the moments of x; bn = (gamma, beta, rmean, rvar, nbt)."""
        sums = torch.empty(2 * 64, dtype=torch.float32, device=self.dev)
        self._ok(self.L.sarssl_stem_input_stats(ptr(x), mode, ptr(flag), ptr(ch), ptr(w64x4), ptr(sums), P, W, H, self.dt, ptr(self.ws), self.ws.numel(),
                                                self.stream), "stem_input_stats")
        stats = torch.empty(4 * 64, dtype=torch.float32, device=self.dev)
        self._ok(self.L.sarssl_batchnorm_finalize(ptr(sums), 1, P, 64, ptr(bn[0]), ptr(bn[1]), 1e-5, 0.1, ptr(bn[2]), ptr(bn[3]), ptr(bn[4]),
                                                  ptr(stats), self.stream), "batchnorm_finalize")
        self.launches += 2
        return stats

    def stem_reduce(self, x, stats, w4x64, out, P):
        sc = _addr(stats, 2 * 64) if stats is not None else None
        sh = _addr(stats, 3 * 64) if stats is not None else None
        self._ok(self.L.sarssl_stem_reduce(ptr(x), sc, sh, ptr(w4x64), ptr(out), P, self.dt, self.stream), "stem_reduce")

    def stem_pw_wgrad(self, wide, wide_stats, narrow, mode, flag, ch, dw64x4, accumulate, P, W, H):
        sc = _addr(wide_stats, 2 * 64) if wide_stats is not None else None
        sh = _addr(wide_stats, 3 * 64) if wide_stats is not None else None
        self._ok(self.L.sarssl_stem_pw_wgrad(ptr(wide), sc, sh, ptr(narrow), mode, ptr(flag), ptr(ch), ptr(dw64x4), int(accumulate), P, W, H, self.dt,
                                             ptr(self.ws), self.ws.numel(), self.stream), "stem_pw_wgrad")
        self.launches += 1

    def stem_head_bwd(self, dz, z, stats, narrow, mode, flag, ch, w64x4, dgamma, dbeta, dw64x4, P, W, H):
        """Backward of conv1x1(4->64)+BN+ReLU in one pass over (dz, z); accumulates into dgamma, dbeta, dw64x4."""
        self._ok(self.L.sarssl_stem_head_bwd(ptr(dz), ptr(z), ptr(stats), ptr(narrow), mode, ptr(flag), ptr(ch), ptr(w64x4), ptr(dgamma), ptr(dbeta),
                                             ptr(dw64x4), P, W, H, self.dt, ptr(self.ws), self.ws.numel(), self.stream), "stem_head_bwd")
        self.launches += 3
        self.tc_launches += 1

    def stem_tail_bwd(self, y, stats, dq, w4x64, dgamma, dbeta, dw4x64, dy, P):
        """Backward of BN+ReLU+conv1x1(64->4): accumulates dgamma, dbeta, dw4x64 and writes dy (gradient w.r.t. the pre-BN activation)."""
        self._ok(self.L.sarssl_stem_tail_bwd(ptr(y), ptr(stats), ptr(dq), ptr(w4x64), ptr(dgamma), ptr(dbeta), ptr(dw4x64), ptr(dy), P, self.dt,
                                             ptr(self.ws), self.ws.numel(), self.stream), "stem_tail_bwd")
        self.launches += 4
        self.tc_launches += 1

    @property
    def conv_tc(self):
        return self.use_tc and self.dtype == torch.bfloat16

    def conv3x3_tc(self, x, wpacked, out, B, H, W, bn=None, in_stats=None):
        """bn = (gamma, beta, running_mean, running_var, num_batches_tracked): also returns the BatchNorm batch statistics of `out`,
        reduced inside the conv epilogue (no extra pass over the tensor).  in_stats: BatchNorm statistics of `x`; the operand becomes
        relu(bn(x)), applied to the landed tiles in shared memory (the activated tensor is never stored)."""
        sc = _addr(in_stats, 2 * 64) if in_stats is not None else None
        sh = _addr(in_stats, 3 * 64) if in_stats is not None else None
        partials = None
        if bn is not None:
            nparts = self.L.sarssl_conv3x3_tc_grid(B, H, W)
            partials = torch.empty(nparts * 128, dtype=torch.float32, device=self.dev)
        self._ok(self.L.sarssl_conv3x3_tc(ptr(x), ptr(wpacked), ptr(out), ptr(partials), sc, sh, B, H, W, self.stream), "conv3x3_tc")
        self.tc_launches += 1
        if bn is None:
            return None
        stats = torch.empty(4 * 64, dtype=torch.float32, device=self.dev)
        self._ok(self.L.sarssl_batchnorm_finalize(ptr(partials), nparts, B * H * W, 64, ptr(bn[0]), ptr(bn[1]), 1e-5, 0.1, ptr(bn[2]), ptr(bn[3]), ptr(bn[4]),
                                                  ptr(stats), self.stream), "batchnorm_finalize")
        return stats

    def conv3x3_wgrad_tc(self, dy, x, dwpacked, B, H, W, in_stats=None):
        sc = _addr(in_stats, 2 * 64) if in_stats is not None else None
        sh = _addr(in_stats, 3 * 64) if in_stats is not None else None
        self._ok(self.L.sarssl_conv3x3_wgrad_tc(ptr(dy), ptr(x), sc, sh, ptr(dwpacked), 0, B, H, W, ptr(self.ws), self.ws.numel(), self.stream),
                 "conv3x3_wgrad_tc")
        self.launches += 1
        self.tc_launches += 1

    def conv3x3(self, x, stats, wpacked, out, B, H, W):
        sc = _addr(stats, 2 * 64) if stats is not None else None
        sh = _addr(stats, 3 * 64) if stats is not None else None
        self._ok(self.L.sarssl_conv3x3(ptr(x), sc, sh, ptr(wpacked), ptr(out), B, H, W, self.dt, self.stream), "conv3x3")

    def conv3x3_wgrad(self, dy, x, stats, dwpacked, B, H, W):
        sc = _addr(stats, 2 * 64) if stats is not None else None
        sh = _addr(stats, 3 * 64) if stats is not None else None
        self._ok(self.L.sarssl_conv3x3_wgrad(ptr(dy), ptr(x), sc, sh, ptr(dwpacked), 0, B, H, W, self.dt, ptr(self.ws), self.ws.numel(), self.stream),
                 "conv3x3_wgrad")
        self.launches += 1

    def adam(self, p, g, m, v, p_bf16, n, step, lr, grad_scale=1.0, zero_grad=True, hyper_dev=None):
        self._ok(self.L.sarssl_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), ptr(p_bf16), n, step, lr, 0.9, 0.999, 1e-8, grad_scale, int(zero_grad),
                                         ptr(hyper_dev), self.stream), "adam_step")
