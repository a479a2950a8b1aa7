/* lws_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, fp64).
 *
 * A plain-C restatement of the reference's Local-Weighted-Sums update rules, used ONLY by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * checker.  Nothing under lws_b200/ imports, links or executes this file.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every entry point below
 * bit-for-bit (max |diff| == 0.0) against oracle/_ref/liblws_ref.so, which oracle/Makefile
 * compiles from the reference's own sources where they lie, and tests/golden/ holds
 * vectors generated from the compiled reference module (tools/make_golden.py).
 *
 * The reference spells each (mode x Q) combination out as a separate 100-line function.
 * Here there is ONE row-update routine parameterised by
 *     fold   : how the +k / -k neighbours share a weight (reference variants Q2 / Q4 / anyQ)
 *     rframe : neighbour frames r < rframe are used on both sides, r >= rframe on the left only
 *     cframe : whether the centre frame's own +-k neighbours are used
 * and the batch / no-future / online entry points are expressed through it:
 *     batch sweep      = rframe Q, cframe 1   (lwslib.cpp:72-150, 153-280, 283-373)
 *     no-future sweep  = rframe 1, cframe 0   (lwslib.cpp:473-535, 620-690)
 *     online row update= Asym_UpdatePhase*    (lwslib.cpp:776-899, 902-1126, 1129-1273)
 * The accumulation ORDER of each reference variant is preserved so results are bit-identical.
 * NoFuture_LWSQ4 (lwslib.cpp:538-617) has an indexing slip (the row offset already contains
 * the bin index, then the bin index is added again); it is the reference behaviour for every
 * hop = fsize/4 configuration, so it is restated literally in nofuture_q4_row().
 */
#include <math.h>
#include <stddef.h>

/* FOLD_FRAC: the reference's *fractionalQ variants (lwslib.cpp:376-467, 693-764, 1276-1421): the anyQ formulas with one
 * weight row PER FREQUENCY BIN -- row (n-L) for the "mod" terms, row N-(n-L), N = 2(Nreal-1), for the "modneg" terms
 * (lwslib.cpp:393, 408).  At the DC bin the reference reads row N of a table that has rows 0..N-1 (create_weights builds
 * T = N rows, lws.pyx:166-178): an out-of-bounds read whose outcome depends on the heap (measured: the compiled reference
 * returns different results for repeated calls on the same input).  The oracle -- and the CUDA path -- define it: the
 * caller hands a table with N+1 rows whose row N is zero (mask clear), i.e. the modneg terms of the DC bin are skipped,
 * which is what the reference does whenever the memory behind its mask array happens to be zero (the common case for the
 * small tables of the tests).  tests/test_oracle_vs_ref.py pins this against the reference's own C functions called
 * on such padded tables. */
enum { FOLD_ANY = 0, FOLD_FRAC = 1, FOLD_Q2 = 2, FOLD_Q4 = 4 };

typedef struct {
    double *Sr, *Si;          /* extended spectrogram, (M+2(Q-1)) x Np, updated in place   */
    const double *wr, *wi;    /* weights (Q, Q, L+1), row-major (lws.pyx:180)              */
    const int *wf;            /* |W| > 1e-12 mask (lws.pyx:231-232)                        */
    const double *amp;        /* |extended spectrogram|, never changes                     */
    int Nreal, L, Q, Np;
} lws_t;

typedef struct { double r, i; } cacc;

/* acc += w*b + conj(w)*c  -- the reference's 2-line idiom (lwslib.cpp:98-99) */
static inline void add_pair(cacc *t, double ar, double ai, double br, double bi, double cr, double ci)
{
    t->r += ar * (br + cr) - ai * (bi - ci);
    t->i += ar * (bi + ci) + ai * (br - cr);
}
/* acc += w*b (lwslib.cpp:500-501) */
static inline void add_one(cacc *t, double ar, double ai, double br, double bi)
{
    t->r += ar * (br) - ai * (bi);
    t->i += ar * (bi) + ai * (br);
}
/* acc += conj(w)*c (lwslib.cpp:667-668) */
static inline void add_one_conj(cacc *t, double ar, double ai, double cr, double ci)
{
    t->r += ar * (cr) + ai * (ci);
    t->i += ar * (ci) - ai * (cr);
}

/* normalise to the stored magnitude and refresh the mirrored (negative / above-Nyquist)
 * copies immediately (lwslib.cpp:356-368) */
static inline void commit_bin(const lws_t *c, int m, int n, cacc t, double a)
{
    const int Np = c->Np, L = c->L, Naux = c->Nreal + c->L - 1;
    double mag = sqrt(t.r * t.r + t.i * t.i);
    if (mag > 0) {
        double *Sr = c->Sr, *Si = c->Si;
        Sr[m * Np + n] = t.r * a / mag;
        Si[m * Np + n] = t.i * a / mag;
        if (n >= L + 1 && n < 2 * L + 1) {
            Sr[m * Np + 2 * L - n] = Sr[m * Np + n];
            Si[m * Np + 2 * L - n] = -Si[m * Np + n];
        } else if (n >= c->Nreal - 1 && n < Naux) {
            Sr[m * Np + 2 * Naux - n] = Sr[m * Np + n];
            Si[m * Np + 2 * Naux - n] = -Si[m * Np + n];
        }
    }
}

/* two-sided contribution of frames m-r and m+r */
static inline void both_sides(const lws_t *c, cacc *t, int m, int n, int r, int wp, int wpn, int fold, int minus)
{
    const int L = c->L, Np = c->Np;
    const double *Sr = c->Sr, *Si = c->Si, *wr = c->wr, *wi = c->wi;
    const int *wf = c->wf;
    const int u = r * (L + 1), im = (m - r) * Np + n, ip = (m + r) * Np + n;
    if (wf[wp + u])
        add_pair(t, wr[wp + u], wi[wp + u], Sr[im], Si[im], Sr[ip], Si[ip]);
    for (int k = 1; k <= L; k++) {
        if (fold == FOLD_ANY) {
            if (wf[wp + u + k])
                add_pair(t, wr[wp + u + k], wi[wp + u + k], Sr[im - k], Si[im - k], Sr[ip - k], Si[ip - k]);
            if (wf[wpn + u + k])
                add_pair(t, wr[wpn + u + k], wi[wpn + u + k], Sr[ip + k], Si[ip + k], Sr[im + k], Si[im + k]);
        } else if (wf[wp + u + k]) {
            double br, bi, cr, ci;
            if (minus) { /* lwslib.cpp:204-207 */
                br = Sr[im - k] - Sr[ip + k]; bi = Si[im - k] - Si[ip + k];
                cr = Sr[ip - k] - Sr[im + k]; ci = Si[ip - k] - Si[im + k];
            } else {     /* lwslib.cpp:123-126 */
                br = Sr[im - k] + Sr[ip + k]; bi = Si[im - k] + Si[ip + k];
                cr = Sr[ip - k] + Sr[im + k]; ci = Si[ip - k] + Si[im + k];
            }
            add_pair(t, wr[wp + u + k], wi[wp + u + k], br, bi, cr, ci);
        }
    }
}

/* contribution of frame m-r only */
static inline void left_side(const lws_t *c, cacc *t, int m, int n, int r, int wp, int wpn, int fold, int minus)
{
    const int L = c->L, Np = c->Np;
    const double *Sr = c->Sr, *Si = c->Si, *wr = c->wr, *wi = c->wi;
    const int *wf = c->wf;
    const int u = r * (L + 1), im = (m - r) * Np + n;
    if (wf[wp + u])
        add_one(t, wr[wp + u], wi[wp + u], Sr[im], Si[im]);
    for (int k = 1; k <= L; k++) {
        if (fold == FOLD_ANY) {
            if (wf[wp + u + k])
                add_one(t, wr[wp + u + k], wi[wp + u + k], Sr[im - k], Si[im - k]);
            if (wf[wpn + u + k])
                add_one_conj(t, wr[wpn + u + k], wi[wpn + u + k], Sr[im + k], Si[im + k]);
        } else if (wf[wp + u + k]) {
            if (minus) /* lwslib.cpp:994-997 */
                add_pair(t, wr[wp + u + k], wi[wp + u + k], Sr[im - k], Si[im - k], -Sr[im + k], -Si[im + k]);
            else       /* lwslib.cpp:872-875 */
                add_pair(t, wr[wp + u + k], wi[wp + u + k], Sr[im - k], Si[im - k], Sr[im + k], Si[im + k]);
        }
    }
}

/* One in-place left-to-right update of extended row m. */
static void update_row(const lws_t *c, int m, int fold, int rframe, int cframe, double thr, int update, double qdiv)
{
    const int L = c->L, Q = c->Q, Np = c->Np, Naux = c->Nreal + c->L - 1;
    for (int n = L; n <= Naux; n++) {
        const double a = c->amp[m * Np + n];
        if (!(a > thr)) continue;
        cacc t = {0., 0.};
        int wp, wpn;
        if (fold == FOLD_FRAC) { /* lwslib.cpp:393, 408: rows n-L and N-(n-L) of the per-frequency table */
            wp = (n - L) * Q * (L + 1);
            wpn = (2 * (c->Nreal - 1) - (n - L)) * Q * (L + 1);
        } else {
            const int p = (n - L) % Q;
            wp = p * Q * (L + 1);
            wpn = ((Q - p) % Q) * Q * (L + 1);
        }
        if (cframe) {
            if (update == 1) { /* never taken through the Python binding (lws.pyx:363 passes 2) */
                t.r += c->Sr[m * Np + n] / qdiv;
                t.i += c->Si[m * Np + n] / qdiv;
            }
            for (int k = 1; k <= L; k++)
                if (c->wf[wp + k])
                    add_pair(&t, c->wr[wp + k], c->wi[wp + k], c->Sr[m * Np + n - k], c->Si[m * Np + n - k],
                             c->Sr[m * Np + n + k], c->Si[m * Np + n + k]);
        }
        if (fold == FOLD_Q4 && (n - L) % 2 == 1) {
            /* odd bins: odd frames first with the sign-flipped folding, then r = 2
             * (lwslib.cpp:186-235, 953-1052) */
            for (int r = 1; r < Q; r += 2) {
                if (r < rframe) both_sides(c, &t, m, n, r, wp, wpn, fold, 1);
                else            left_side(c, &t, m, n, r, wp, wpn, fold, 1);
            }
            if (2 < rframe) both_sides(c, &t, m, n, 2, wp, wpn, fold, 0);
            else            left_side(c, &t, m, n, 2, wp, wpn, fold, 0);
        } else {
            const int f = fold == FOLD_FRAC ? FOLD_ANY : fold; /* same formulas as anyQ, other weight rows */
            for (int r = 1; r < rframe; r++) both_sides(c, &t, m, n, r, wp, wpn, f, 0);
            for (int r = rframe; r < Q; r++) left_side(c, &t, m, n, r, wp, wpn, f, 0);
        }
        commit_bin(c, m, n, t, a);
    }
}

/* NoFuture_LWSQ4 exactly as the reference computes it (lwslib.cpp:547-616): frames in
 * descending order, +-k terms before the k = 0 term, and every read at flat offset
 * (m-r)*Np + 2n -+ k because `im` already holds n (lwslib.cpp:559, 567-570, 580-583, 593-594). */
static void nofuture_q4_row(const lws_t *c, int m, double thr)
{
    const int L = c->L, Q = 4, Np = c->Np, Naux = c->Nreal + c->L - 1;
    const double *Sr = c->Sr, *Si = c->Si;
    for (int n = L; n <= Naux; n++) {
        const double a = c->amp[m * Np + n];
        if (!(a > thr)) continue;
        cacc t = {0., 0.};
        const int wp = ((n - L) % Q) * Q * (L + 1);
        for (int r = Q - 1; r > 0; r--) {
            const int u = r * (L + 1);
            const ptrdiff_t f = (ptrdiff_t)(m - r) * Np + 2 * n;
            const int minus = ((n - L) % 2 == 1) && (r % 2 == 1);
            for (int k = 1; k <= L; k++) {
                if (!c->wf[wp + u + k]) continue;
                if (minus)
                    add_pair(&t, c->wr[wp + u + k], c->wi[wp + u + k], Sr[f - k], Si[f - k], -Sr[f + k], -Si[f + k]);
                else
                    add_pair(&t, c->wr[wp + u + k], c->wi[wp + u + k], Sr[f - k], Si[f - k], Sr[f + k], Si[f + k]);
            }
            if (c->wf[wp + u])
                add_one(&t, c->wr[wp + u], c->wi[wp + u], Sr[f], Si[f]);
        }
        commit_bin(c, m, n, t, a);
    }
}

static lws_t make_ctx(double *Sr, double *Si, const double *wr, const double *wi, const int *wf,
                      const double *amp, int Nreal, int L, int Q)
{
    lws_t c = {Sr, Si, wr, wi, wf, amp, Nreal, L, Q, Nreal + 2 * L};
    return c;
}

/* ------------------------------------------------------------------ public entry points */

/* lwslib.cpp:15-44 */
void orc_extend_spec(double *ESr, double *ESi, const double *Sr, const double *Si, int Nreal, int M, int L, int Q)
{
    const int Np = Nreal + 2 * L;
    for (int m = 0; m < M + 2 * (Q - 1); m++) {
        int p = m - (Q - 1);
        p = p < 0 ? 0 : (p > M - 1 ? M - 1 : p);
        double *er = ESr + (size_t)m * Np, *ei = ESi + (size_t)m * Np;
        const double *sr = Sr + (size_t)p * Nreal, *si = Si + (size_t)p * Nreal;
        for (int n = 0; n < Nreal; n++) { er[n + L] = sr[n]; ei[n + L] = si[n]; }
        for (int n = 0; n < L; n++) { er[n] = sr[L - n]; ei[n] = -si[L - n]; }
        for (int n = Nreal + L; n < Np; n++) {
            er[n] = er[2 * (Nreal + L - 1) - n];
            ei[n] = -ei[2 * (Nreal + L - 1) - n];
        }
    }
}

/* lwslib.cpp:47-57 */
void orc_copy_spec(const double *ESr, const double *ESi, double *Sr, double *Si, int Nreal, int M, int L, int Q)
{
    const int Np = Nreal + 2 * L;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < Nreal; n++) {
            Sr[(size_t)m * Nreal + n] = ESr[(size_t)(m + Q - 1) * Np + n + L];
            Si[(size_t)m * Nreal + n] = ESi[(size_t)(m + Q - 1) * Np + n + L];
        }
}

/* lwslib.cpp:59-65 */
void orc_amp_spec(const double *Sr, const double *Si, double *amp, int size)
{
    for (int n = 0; n < size; n++) amp[n] = sqrt(Sr[n] * Sr[n] + Si[n] * Si[n]);
}

/* One batch sweep.  fold = 2 / 4 / 0 / 1 selects LWSQ2 / LWSQ4 / LWSanyQ / LWSfractionalQ. */
void orc_batch_sweep(int fold, double *Sr, double *Si, const double *wr, const double *wi, const int *wf,
                     const double *amp, int Nreal, int M, int L, int Q, double thr)
{
    lws_t c = make_ctx(Sr, Si, wr, wi, wf, amp, Nreal, L, Q);
    for (int m = Q - 1; m < M + Q - 1; m++) update_row(&c, m, fold, Q, 1, thr, 2, (double)Q);
}

/* One no-future sweep.  fold = 2 / 4 / 0 selects NoFuture_LWSQ2 / Q4 / anyQ. */
void orc_nofuture_sweep(int fold, double *Sr, double *Si, const double *wr, const double *wi, const int *wf,
                        const double *amp, int Nreal, int M, int L, int Q, double thr)
{
    lws_t c = make_ctx(Sr, Si, wr, wi, wf, amp, Nreal, L, Q);
    for (int m = Q - 1; m < M + Q - 1; m++) {
        if (fold == FOLD_Q4) nofuture_q4_row(&c, m, thr);
        else update_row(&c, m, fold, 1, 0, thr, 2, (double)Q);
    }
}

/* Asym_UpdatePhase{Q2,Q4,anyQ}: update rows Q-1 .. M+Q-2 of the buffer handed in, using at
 * most M0 frames to the right (lwslib.cpp:788-798). */
void orc_asym_update(int fold, double *Sr, double *Si, const double *wr, const double *wi, const int *wf,
                     const double *amp, int Nreal, int M, int M0, int L, int Q, double thr, int update)
{
    lws_t c = make_ctx(Sr, Si, wr, wi, wf, amp, Nreal, L, Q);
    for (int m = Q - 1; m < M + Q - 1; m++) {
        int rframe = M0 + Q - m - 1, cframe = 1;
        if (rframe > Q) rframe = Q;
        if (rframe < 1) { cframe = 0; rframe = 1; }
        update_row(&c, m, fold, rframe, cframe, thr, update, (double)Q);
    }
}

/* TF_RTISI_LA (lwslib.cpp:1424-1492); fold = FOLD_FRAC is its !use_summarized_weights branch. */
void orc_rtisi_la(int fold, double *Sr, double *Si, const double *wr, const double *wi, const double *wr_ai,
                  const double *wi_ai, const double *wr_af, const double *wi_af, const int *wf, const int *wf_ai,
                  const int *wf_af, const double *amp, int iter, int LA, int Nreal, int M, int L, int Q,
                  const double *thresholds, int update)
{
    const int Np = Nreal + 2 * L;
    for (int m = 0; m < M; m++) {
        int lframe = m - LA, nframe = LA;
        if (lframe < 0) { lframe = 0; nframe = m; }
        const size_t om = (size_t)m * Np, ol = (size_t)lframe * Np;
        orc_asym_update(fold, Sr + om, Si + om, wr_ai, wi_ai, wf_ai, amp + om, Nreal, 1, 0, L, Q, 0., update);
        for (int h = 0; h < iter; h++) {
            const double thr = thresholds[h];
            if (LA > 0)
                orc_asym_update(fold, Sr + ol, Si + ol, wr, wi, wf, amp + ol, Nreal, nframe, nframe + 1, L, Q, thr, update);
            orc_asym_update(fold, Sr + om, Si + om, wr_af, wi_af, wf_af, amp + om, Nreal, 1, 1, L, Q, thr, update);
        }
    }
}
