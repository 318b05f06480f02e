/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, scalar, CPU restatement of the Zuko flow hot path
 * (NormalizingFlow.log_prob / transform.inv for MAF / NSF / coupling flows).
 * It exists to CHECK the CUDA product; nothing under zuko_b200/ may import,
 * link or call it.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py use it.
 *
 * Parity of this restatement is PINNED against the unmodified Python reference:
 * the .npz files under tests/golden hold inputs/outputs produced by importing /root/reference
 * (script: tests/golden/make_golden.py); tests/test_oracle.py checks every
 * function below against them in fp32 and fp64.
 *
 * This header is included twice by zuko_oracle.c, once with REAL=float
 * (suffix _f32, mimics the reference's fp32 op order where that is cheap) and
 * once with REAL=double (suffix _f64, the arbiter used for tolerance decisions).
 *
 * Citations are file:line in /root/reference (probabilists/zuko @ 1063ae4) or
 * torch/... for the installed torch 2.11.
 */

#ifndef ZO_RB
#define ZO_RB 8 /* rows per block of zo_linear */
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* zuko/transforms.py:299-316  SoftclipTransform: y = x / (1 + |x/B|) */
static inline REAL FN(softclip_)(REAL x, REAL bound) { return x / (1 + R_ABS(x / bound)); }

/* ------------------------------------------------------------------------- */
/* zuko/nn.py:217-218  MaskedLinear.forward: F.linear(x, mask * W, b), followed
 * by the ReLU that MaskedMLP / MLP insert between layers (nn.py:311-313,
 * nn.py:172-186).  x may come from two row-major sources (x | c) as the
 * torch.cat in flows/autoregressive.py:209 does; ldc == 0 broadcasts one
 * context row (utils.py:236-244).  mask may be NULL (dense MLP, nn.py:122). */
void FN(zo_linear)(const REAL* x, int64_t ldx, int dx, const REAL* c, int64_t ldc, int dc,
                   int64_t B, const REAL* W, const uint8_t* mask, const REAL* bias, int out,
                   int relu, REAL* y, int64_t ldy) {
    const int in = dx + dc;
    /* mask * W (nn.py:218), stored transposed (in, out) so that the inner loop runs over the
     * contiguous outputs and vectorises; the sum over the inputs stays in index order. */
    REAL* Wt = (REAL*)malloc((size_t)in * out * sizeof(REAL));
    for (int o = 0; o < out; ++o)
        for (int i = 0; i < in; ++i) {
            const int64_t k = (int64_t)o * in + i;
            Wt[(int64_t)i * out + o] = (mask && !mask[k]) ? (REAL)0 : W[k];
        }
    /* ZO_RB rows share every weight row while it sits in L1 (the per-row loop streamed the whole
     * matrix from L2 once per sample); each output element still accumulates over the inputs in
     * index order, so the results are bit-identical to the row-at-a-time loop. */
#pragma omp parallel for schedule(static)
    for (int64_t r0 = 0; r0 < B; r0 += ZO_RB) {
        const int nr = (int)((B - r0 < ZO_RB) ? (B - r0) : ZO_RB);
        for (int j = 0; j < nr; ++j) {
            REAL* __restrict__ yr = y + (r0 + j) * ldy;
            for (int o = 0; o < out; ++o) yr[o] = 0;
        }
        for (int i = 0; i < in; ++i) {
            const REAL* __restrict__ w = Wt + (int64_t)i * out;
            for (int j = 0; j < nr; ++j) {
                const int64_t r = r0 + j;
                const REAL xv = (i < dx) ? x[r * ldx + i] : c[r * ldc + (i - dx)];
                REAL* __restrict__ yr = y + r * ldy;
                for (int o = 0; o < out; ++o) yr[o] += xv * w[o];
            }
        }
        for (int j = 0; j < nr; ++j) {
            REAL* __restrict__ yr = y + (r0 + j) * ldy;
            for (int o = 0; o < out; ++o) {
                REAL v = yr[o] + (bias ? bias[o] : (REAL)0);
                yr[o] = (relu && v < 0) ? (REAL)0 : v;
            }
        }
    }
    free(Wt);
}

/* ------------------------------------------------------------------------- */
/* zuko/transforms.py:469-490  MonotonicRQSTransform.__init__ for ONE
 * (sample, dim): raw (w[K], h[K], d[K-1]) -> knots X[K+1], Y[K+1], Dv[K+1]. */
static void FN(rqs_knots_)(const REAL* w, const REAL* h, const REAL* d, int K, REAL bound,
                           REAL slope, REAL* X, REAL* Y, REAL* Dv) {
    const REAL L = R_LOG(slope); /* math.log(slope) < 0; only |.| is used */
    REAL ws[ZO_MAX_BINS], hs[ZO_MAX_BINS];
    REAL mw = -INFINITY, mh = -INFINITY;
    for (int k = 0; k < K; ++k) {
        ws[k] = w[k] / (1 + R_ABS(2 * w[k] / L)); /* :480 */
        hs[k] = h[k] / (1 + R_ABS(2 * h[k] / L)); /* :481 */
        if (ws[k] > mw) mw = ws[k];
        if (hs[k] > mh) mh = hs[k];
    }
    /* :484-485 F.softmax (max-subtracted, torch/aten SoftMax) then left pad 0 */
    REAL sw = 0, sh = 0;
    for (int k = 0; k < K; ++k) {
        ws[k] = R_EXP(ws[k] - mw);
        hs[k] = R_EXP(hs[k] - mh);
        sw += ws[k];
        sh += hs[k];
    }
    /* :488-489 cumsum of the padded softmax; knots = bound * (2 cum - 1) */
    REAL cw = 0, ch = 0;
    X[0] = bound * (2 * cw - 1);
    Y[0] = bound * (2 * ch - 1);
    for (int k = 0; k < K; ++k) {
        cw += ws[k] / sw;
        ch += hs[k] / sh;
        X[k + 1] = bound * (2 * cw - 1);
        Y[k + 1] = bound * (2 * ch - 1);
    }
    /* :482,486,490 derivatives: softclip, pad (1,1) with 0, exp */
    Dv[0] = 1;
    Dv[K] = 1;
    for (int k = 0; k < K - 1; ++k) Dv[k + 1] = R_EXP(d[k] / (1 + R_ABS(d[k] / L)));
}

/* zuko/transforms.py:521-523 searchsorted (strict <) and :499-519 bin():
 * k = #{j : seq_j < v} - 1, mask = 0 <= k < K, k %= K. */
static inline int FN(rqs_search_)(const REAL* seq, int K, REAL v, int* inside) {
    int cnt = 0;
    for (int j = 0; j <= K; ++j) cnt += (seq[j] < v) ? 1 : 0;
    int k = cnt - 1;
    *inside = (0 <= k && k < K);
    k = ((k % K) + K) % K; /* python modulo */
    return k;
}

/* zuko/transforms.py:554-567 call_and_ladj for one element. Non-finite x
 * follow the reference arithmetic (mask * (x - x0) etc.) as closely as a
 * branching restatement can: inside the domain the formula is evaluated
 * verbatim; outside it is identity with ladj = 0 * log(jac(z=0)) = 0. */
static inline void FN(rqs_eval_fwd_)(const REAL* X, const REAL* Y, const REAL* Dv, int K, REAL x,
                                     REAL* y, REAL* ladj) {
    int inside;
    int k = FN(rqs_search_)(X, K, x, &inside);
    REAL x0 = X[k], x1 = X[k + 1], y0 = Y[k], y1 = Y[k + 1], d0 = Dv[k], d1 = Dv[k + 1];
    REAL s = (y1 - y0) / (x1 - x0);
    REAL z = inside ? (x - x0) / (x1 - x0) : (REAL)0 * (x - x0) / (x1 - x0);
    REAL z1 = z * (1 - z);
    REAL den = s + (d0 + d1 - 2 * s) * z1;
    REAL yy = y0 + (y1 - y0) * (s * z * z + d0 * z1) / den;
    REAL jac = s * s * (2 * s * z1 + d0 * (1 - z) * (1 - z) + d1 * z * z) / (den * den);
    *y = inside ? yy : x;
    *ladj = inside ? R_LOG(jac) : (REAL)0 * R_LOG(jac);
}

/* zuko/transforms.py:534-548 _inverse for one element. */
static inline REAL FN(rqs_eval_inv_)(const REAL* X, const REAL* Y, const REAL* Dv, int K, REAL y) {
    int inside;
    int k = FN(rqs_search_)(Y, K, y, &inside);
    REAL x0 = X[k], x1 = X[k + 1], y0 = Y[k], y1 = Y[k + 1], d0 = Dv[k], d1 = Dv[k + 1];
    REAL s = (y1 - y0) / (x1 - x0);
    REAL y_ = inside ? (y - y0) : (REAL)0 * (y - y0);
    REAL a = (y1 - y0) * (s - d0) + y_ * (d0 + d1 - 2 * s);
    REAL b = (y1 - y0) * d0 - y_ * (d0 + d1 - 2 * s);
    REAL c = -s * y_;
    REAL z = 2 * c / (-b - R_SQRT(b * b - 4 * a * c));
    REAL x = x0 + z * (x1 - x0);
    return inside ? x : y;
}

/* Batched element-wise RQS over (B, D) with per-(sample,dim) parameters
 * phi[(r * phi_ld) + d * (3K-1) + {0..K-1 | K..2K-1 | 2K..3K-2}] — the layout
 * produced by flows/autoregressive.py:149,212-213 (row index d*total + p).
 * phi_ld == 0 shares one (D, 3K-1) table across the batch
 * (flows/gaussianization.py:74-77,86-94).  ladj is per element (B, D). */
void FN(zo_rqs_forward)(const REAL* x, int64_t ldx, const REAL* phi, int64_t phi_ld, int64_t B,
                        int D, int K, REAL bound, REAL slope, REAL* y, int64_t ldy, REAL* ladj,
                        int64_t ldl) {
    const int P = 3 * K - 1;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < B; ++r) {
        REAL X[ZO_MAX_BINS + 1], Y[ZO_MAX_BINS + 1], Dv[ZO_MAX_BINS + 1];
        for (int d = 0; d < D; ++d) {
            const REAL* p = phi + r * phi_ld + (int64_t)d * P;
            FN(rqs_knots_)(p, p + K, p + 2 * K, K, bound, slope, X, Y, Dv);
            FN(rqs_eval_fwd_)(X, Y, Dv, K, x[r * ldx + d], &y[r * ldy + d], &ladj[r * ldl + d]);
        }
    }
}

void FN(zo_rqs_inverse)(const REAL* y, int64_t ldy, const REAL* phi, int64_t phi_ld, int64_t B,
                        int D, int K, REAL bound, REAL slope, REAL* x, int64_t ldx) {
    const int P = 3 * K - 1;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < B; ++r) {
        REAL X[ZO_MAX_BINS + 1], Y[ZO_MAX_BINS + 1], Dv[ZO_MAX_BINS + 1];
        for (int d = 0; d < D; ++d) {
            const REAL* p = phi + r * phi_ld + (int64_t)d * P;
            FN(rqs_knots_)(p, p + K, p + 2 * K, K, bound, slope, X, Y, Dv);
            x[r * ldx + d] = FN(rqs_eval_inv_)(X, Y, Dv, K, y[r * ldy + d]);
        }
    }
}

/* Exposes the knot construction itself (for knot-level golden checks). */
void FN(zo_rqs_knots)(const REAL* phi, int64_t n, int K, REAL bound, REAL slope, REAL* X, REAL* Y,
                      REAL* Dv) {
    const int P = 3 * K - 1;
    for (int64_t i = 0; i < n; ++i) {
        const REAL* p = phi + i * P;
        FN(rqs_knots_)(p, p + K, p + 2 * K, K, bound, slope, X + i * (K + 1), Y + i * (K + 1),
                       Dv + i * (K + 1));
    }
}

/* ------------------------------------------------------------------------- */
/* zuko/transforms.py:426-446 MonotonicAffineTransform with
 * phi[..., d, 0] = shift, phi[..., d, 1] = unconstrained scale
 * (flows/autoregressive.py:96,213: shapes=((),()) -> shift first). */
void FN(zo_affine_forward)(const REAL* x, int64_t ldx, const REAL* phi, int64_t phi_ld, int64_t B,
                           int D, REAL slope, REAL* y, int64_t ldy, REAL* ladj, int64_t ldl) {
    const REAL L = R_LOG(slope);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < B; ++r)
        for (int d = 0; d < D; ++d) {
            const REAL* p = phi + r * phi_ld + (int64_t)d * 2;
            REAL ls = p[1] / (1 + R_ABS(p[1] / L)); /* :437 */
            y[r * ldy + d] = x[r * ldx + d] * R_EXP(ls) + p[0]; /* :438,440-441 */
            ladj[r * ldl + d] = ls; /* :445-446 */
        }
}

void FN(zo_affine_inverse)(const REAL* y, int64_t ldy, const REAL* phi, int64_t phi_ld, int64_t B,
                           int D, REAL slope, REAL* x, int64_t ldx) {
    const REAL L = R_LOG(slope);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < B; ++r)
        for (int d = 0; d < D; ++d) {
            const REAL* p = phi + r * phi_ld + (int64_t)d * 2;
            REAL ls = p[1] / (1 + R_ABS(p[1] / L));
            x[r * ldx + d] = (y[r * ldy + d] - p[0]) / R_EXP(ls); /* :443-444 */
        }
}

/* zuko/transforms.py:299-316 SoftclipTransform, element-wise (B*D flat). */
void FN(zo_softclip_forward)(const REAL* x, int64_t n, REAL bound, REAL* y, REAL* ladj) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        y[i] = x[i] / (1 + R_ABS(x[i] / bound));           /* :309-310 */
        ladj[i] = -2 * R_LOG1P(R_ABS(x[i] / bound));       /* :315-316 */
    }
}
void FN(zo_softclip_inverse)(const REAL* y, int64_t n, REAL bound, REAL* x) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) x[i] = y[i] / (1 - R_ABS(y[i] / bound)); /* :312-313 */
}

/* zuko/transforms.py:1235-1244 RotationTransform: y = R x (transpose=0),
 * x = R^T y (transpose=1).  R = matrix_exp(A - A^T) is built by the caller. */
void FN(zo_rotate)(const REAL* x, int64_t ldx, const REAL* R, int64_t B, int D, int transpose,
                   REAL* y, int64_t ldy) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < B; ++r)
        for (int i = 0; i < D; ++i) {
            REAL acc = 0;
            for (int j = 0; j < D; ++j)
                acc += (transpose ? R[(int64_t)j * D + i] : R[(int64_t)i * D + j]) * x[r * ldx + j];
            y[r * ldy + i] = acc;
        }
}

/* torch/distributions/normal.py:87-102 Normal.log_prob summed over the event
 * dim by Independent (torch/distributions/independent.py:120-122), i.e.
 * zuko/distributions.py:337-363 DiagNormal.log_prob; plus the flow ladj
 * (zuko/distributions.py:115-119). out[r] = sum_d N(z_rd; loc_d, scale_d) + ladj[r]. */
void FN(zo_diag_normal_log_prob)(const REAL* z, int64_t ldz, const REAL* loc, const REAL* scale,
                                 int64_t B, int D, const REAL* ladj, REAL* out) {
    const REAL half_log_2pi = (REAL)0.91893853320467274178; /* log(sqrt(2 pi)) */
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < B; ++r) {
        REAL acc = 0;
        for (int d = 0; d < D; ++d) {
            REAL var = scale[d] * scale[d];
            REAL diff = z[r * ldz + d] - loc[d];
            acc += -(diff * diff) / (2 * var) - R_LOG(scale[d]) - half_log_2pi;
        }
        out[r] = acc + (ladj ? ladj[r] : (REAL)0);
    }
}

/* zuko/transforms.py:210-214 DependentTransform: ladj.sum(-1) over D, then
 * zuko/transforms.py:147 acc = acc + ladj.  acc[r] += sum_d ladj[r, d]. */
void FN(zo_sum_ladj)(const REAL* ladj, int64_t ldl, int64_t B, int D, REAL* acc) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < B; ++r) {
        REAL s = 0;
        for (int d = 0; d < D; ++d) s += ladj[r * ldl + d];
        acc[r] += s;
    }
}

#undef FN
#undef CAT
#undef CAT_
