// strip_pair.cuh -- pair-split block update of the cluster strip kernel (included by kernels_batch.cu).
//
// Two adjacent lanes share one task (one 8-bin block of one frame of one sweep): lane 2p holds the REAL
// component of everything, lane 2p+1 the IMAGINARY one.  The reference's term
//        vr = ar*(br+cr) - ai*(bi-ci)          vi = ar*(bi+ci) + ai*(br-cr)            (lwslib.cpp:98-99)
// splits cleanly: each lane forms s = b+c and d = b-c of its own component, the two lanes swap d (one
// 64-bit shuffle, sign flipped on the real lane), and each computes  ar*s + ai*d'  -- the same IEEE
// operations, rounded one by one, as the scalar code (x - y == x + (-y) and (-a)*b == -(a*b) exactly).
// The ordered sums of the real and of the imaginary part are independent chains, |t|^2 = tr*tr + ti*ti needs one
// more swap (addition commutes bit for bit), and each lane does ONE of the two divisions.
//
// What this buys on B200:
//   * twice the warps for the same shared-memory ring (the ring, not the register file, caps the number of
//     tasks per SM), each with half the fp64 instructions per bin: a lone warp cannot issue fp64 faster
//     than one instruction per ~2.6 clk, and 4-6 such warps left the fp64 pipe three quarters idle;
//   * half the registers per value, which makes room for SLIDING WINDOWS: the neighbour frames m-+r are read
//     from shared memory once per block (18 columns per row) into registers instead of once per use
//     (59 loads per bin): shared-memory bandwidth, 7.4 clk/bin/SM of traffic before, was the tighter floor.


// ring accessor of one lane: component h of cell (frame + dr, block column + dcol)
template <int Q>
struct PairCell {
    const unsigned char *base;      // ring + 8 * h
    unsigned rowoff[2 * Q - 1];     // byte offset of ring row (frame + dr) at [dr + Q - 1]
    int col0;                       // ring column of bin 0 of the block
    __device__ __forceinline__ double ld(int dr, int dcol) const
    {
        return *reinterpret_cast<const double *>(base + rowoff[dr + Q - 1] + (unsigned)(col0 + dcol) * 16u);
    }
};

// Which neighbour rows are held in register windows.  WM = 0: none (every use is a shared-memory load),
// 1: odd |dr| (for Q = 4 the dense frame pairs r = 1 and 3; r = 2 has 4 terms and is read directly), 2: all.
template <int Q, int WM>
__device__ __forceinline__ constexpr int pair_win_slot(int dr)
{
    const int a = dr < 0 ? -dr : dr;
    if (WM == 0 || a == 0) return -1;
    if (WM == 1) return (a & 1) ? (a - 1) + (dr > 0 ? 1 : 0) : -1;
    return 2 * (a - 1) + (dr > 0 ? 1 : 0);
}
template <int Q, int WM>
struct PairWin {
    static constexpr int NW = WM == 0 ? 0 : (WM == 1 ? 2 * (Q / 2) : 2 * (Q - 1));
    double v[NW > 0 ? NW : 1][SBK + 2 * SL]; // [slot][column + SL], columns -SL .. SBK + SL - 1 of the block
};

// columns CI of all windowed rows
template <int Q, int WM, int CI>
__device__ __forceinline__ void pair_win_load(const PairCell<Q> &cell, PairWin<Q, WM> &win)
{
#pragma unroll
    for (int r = 1; r < Q; ++r) {
        if (pair_win_slot<Q, WM>(-r) >= 0) {
            win.v[pair_win_slot<Q, WM>(-r) >= 0 ? pair_win_slot<Q, WM>(-r) : 0][CI + SL] = cell.ld(-r, CI);
            win.v[pair_win_slot<Q, WM>(+r) >= 0 ? pair_win_slot<Q, WM>(+r) : 0][CI + SL] = cell.ld(+r, CI);
        }
    }
}
template <int Q, int WM, int C0, int C1>
__device__ __forceinline__ void pair_win_load_range(const PairCell<Q> &cell, PairWin<Q, WM> &win)
{
    if constexpr (C0 <= C1) {
        pair_win_load<Q, WM, C0>(cell, win);
        pair_win_load_range<Q, WM, C0 + 1, C1>(cell, win);
    }
}

// the other lane's value; `sm` = 0x80000000 on the real lane (which needs the negated difference), 0 on the other
__device__ __forceinline__ double pair_swap(double d, unsigned sm)
{
    const int hi = __shfl_xor_sync(0xffffffffu, __double2hiint(d), 1) ^ (int)sm;
    const int lo = __shfl_xor_sync(0xffffffffu, __double2loint(d), 1);
    return __hiloint2double(hi, lo);
}

// this lane's component of  w*b + conj(w)*c  given its components of b and c
__device__ __forceinline__ double pair_value_h(double ar, double ai, double bh, double ch, unsigned sm)
{
    const double s = __dadd_rn(bh, ch), d = __dsub_rn(bh, ch);
    const double dx = pair_swap(d, sm);
    return __dadd_rn(__dmul_rn(ar, s), __dmul_rn(ai, dx));
}

template <int Q, int FOLD>
struct PairTerms { // this lane's component of the inter-frame term values of one bin, in the reference's order of addition
    static constexpr int N = (Q - 1) * TermCount<FOLD>::per_r;
    double v[N > 0 ? N : 1];
};

// values of frame pair (m - R_, m + R_) for bin I of the block into slots [BASE, BASE + per_r)
template <int Q, int P, int FOLD, int PAT, int WM, int I, int R_, bool MINUS, int BASE>
__device__ __forceinline__ void pair_terms_r(const PairCell<Q> &cell, const PairWin<Q, WM> &win, const StripW<Q> &w, unsigned sm,
                                             PairTerms<Q, FOLD> &tv)
{
    constexpr int PN = (Q - P) % Q;
    auto E = [&](int dr, int dk) -> double {
        const int s = pair_win_slot<Q, WM>(dr);
        return s >= 0 ? win.v[s >= 0 ? s : 0][I + dk + SL] : cell.ld(dr, I + dk);
    };
    if (pat_has<Q, PAT>(R_, 0)) tv.v[BASE] = pair_value_h(w.wr[P][R_][0], w.wi[P][R_][0], E(-R_, 0), E(+R_, 0), sm);
#pragma unroll
    for (int k = 1; k <= SL; ++k) {
        if (FOLD == LWSB_FOLD_ANY) {
            if (pat_has<Q, PAT>(R_, k)) {
                tv.v[BASE + 2 * k - 1] = pair_value_h(w.wr[P][R_][k], w.wi[P][R_][k], E(-R_, -k), E(+R_, -k), sm);
                tv.v[BASE + 2 * k] = pair_value_h(w.wr[PN][R_][k], w.wi[PN][R_][k], E(+R_, +k), E(-R_, +k), sm);
            }
        } else if (pat_has<Q, PAT>(R_, k)) {
            const double e1 = E(-R_, -k), e2 = E(+R_, +k), e3 = E(+R_, -k), e4 = E(-R_, +k);
            const double bh = MINUS ? __dsub_rn(e1, e2) : __dadd_rn(e1, e2); // lwslib.cpp:204-207 / 123-126
            const double ch = MINUS ? __dsub_rn(e3, e4) : __dadd_rn(e3, e4);
            tv.v[BASE + k] = pair_value_h(w.wr[P][R_][k], w.wi[P][R_][k], bh, ch, sm);
        }
    }
}

template <int Q, int P, int FOLD, int PAT, int R_, int BASE>
__device__ __forceinline__ void pair_accumulate_r(const StripW<Q> &w, const PairTerms<Q, FOLD> &tv, double &t)
{
    constexpr int PN = (Q - P) % Q;
    auto add = [&](int slot, unsigned flagword, int k) {
        const double nt = __dadd_rn(t, tv.v[slot]);
        if (PAT == 1) t = nt;
        else t = ((flagword >> k) & 1u) ? nt : t;
    };
    if (pat_has<Q, PAT>(R_, 0)) add(BASE, w.flag[P][R_], 0);
#pragma unroll
    for (int k = 1; k <= SL; ++k) {
        if (!pat_has<Q, PAT>(R_, k)) continue;
        if (FOLD == LWSB_FOLD_ANY) {
            add(BASE + 2 * k - 1, w.flag[P][R_], k);
            add(BASE + 2 * k, w.flag[PN][R_], k);
        } else add(BASE + k, w.flag[P][R_], k);
    }
}

template <int Q, int P, int FOLD, int PAT, int WM, int I>
__device__ __forceinline__ void pair_bin_terms(const PairCell<Q> &cell, const PairWin<Q, WM> &win, const StripW<Q> &w, unsigned sm,
                                               PairTerms<Q, FOLD> &tv)
{
    constexpr int TPR = TermCount<FOLD>::per_r;
    if constexpr (FOLD == LWSB_FOLD_Q4 && (P & 1)) { // odd bins: r = 1, 3 sign-flipped, then r = 2 (lwslib.cpp:186-235)
        pair_terms_r<Q, P, FOLD, PAT, WM, I, 1, true, 0>(cell, win, w, sm, tv);
        pair_terms_r<Q, P, FOLD, PAT, WM, I, 3, true, TPR>(cell, win, w, sm, tv);
        pair_terms_r<Q, P, FOLD, PAT, WM, I, 2, false, 2 * TPR>(cell, win, w, sm, tv);
    } else {
        if constexpr (Q > 1) pair_terms_r<Q, P, FOLD, PAT, WM, I, 1, false, 0>(cell, win, w, sm, tv);
        if constexpr (Q > 2) pair_terms_r<Q, P, FOLD, PAT, WM, I, 2, false, TPR>(cell, win, w, sm, tv);
        if constexpr (Q > 3) pair_terms_r<Q, P, FOLD, PAT, WM, I, 3, false, 2 * TPR>(cell, win, w, sm, tv);
    }
}

template <int Q, int P, int FOLD, int PAT>
__device__ __forceinline__ void pair_bin_accumulate(const StripW<Q> &w, const PairTerms<Q, FOLD> &tv, double &t)
{
    constexpr int TPR = TermCount<FOLD>::per_r;
    if constexpr (FOLD == LWSB_FOLD_Q4 && (P & 1)) {
        pair_accumulate_r<Q, P, FOLD, PAT, 1, 0>(w, tv, t);
        pair_accumulate_r<Q, P, FOLD, PAT, 3, TPR>(w, tv, t);
        pair_accumulate_r<Q, P, FOLD, PAT, 2, 2 * TPR>(w, tv, t);
    } else {
        if constexpr (Q > 1) pair_accumulate_r<Q, P, FOLD, PAT, 1, 0>(w, tv, t);
        if constexpr (Q > 2) pair_accumulate_r<Q, P, FOLD, PAT, 2, TPR>(w, tv, t);
        if constexpr (Q > 3) pair_accumulate_r<Q, P, FOLD, PAT, 3, 2 * TPR>(w, tv, t);
    }
}

// own row of the block as this lane sees it while the block is updated (PAT = 1 only: the centre-frame term of
// the default windows reaches one bin): cur[i + 1] = current component of column i, i = -1 .. SBK
struct PairOwn { double cur[SBK + 2]; };

// bins I .. 7 of a block.  PIPE = 1: `tv` holds the inter-frame term values of bin I on entry and the values of
// bin I + 1 are formed, explicitly, ahead of bin I's order-bound chain (more registers); PIPE = 0: each bin forms
// its own values and the instruction scheduler overlaps what the register budget allows.
template <int Q, int FOLD, int PAT, int WM, int PIPE, int I>
__device__ __forceinline__ void pair_block(const PairCell<Q> &cell, PairWin<Q, WM> &win, const StripW<Q> &w, const BlockCtx &bc,
                                           const double *amp, unsigned active, unsigned sm, int h, PairTerms<Q, FOLD> &tv,
                                           PairOwn &own)
{
    if constexpr (I < SBK) {
        constexpr int P = I % Q;
        // (1) inter-frame terms: independent of everything below
        PairTerms<Q, FOLD> tvn;
        if constexpr (I + 1 < SBK) pair_win_load<Q, WM, I + 1 + SL>(cell, win);
        if constexpr (PIPE == 1) {
            if constexpr (I + 1 < SBK) pair_bin_terms<Q, (I + 1) % Q, FOLD, PAT, WM, I + 1>(cell, win, w, sm, tvn);
        } else {
            pair_bin_terms<Q, P, FOLD, PAT, WM, I>(cell, win, w, sm, tv);
        }
        // (2) this bin: centre-frame terms (they see the bins just updated), then the ordered sum
        const int n = bc.n0 + I;
        double t = 0.0;
        if constexpr (PAT == 1) {
            // k = 1 only.  E(0, +1) is the value the block started with, except at the Nyquist bin, whose right
            // neighbour is the mirror image of the bin just updated (lwslib.cpp:362-368)
            const double b = own.cur[I];
            double c = own.cur[I + 2];
            if (I > 0 && n == bc.Nreal - 1) c = h ? -b : b;
            t = __dadd_rn(t, pair_value_h(w.wr[P][0][1], w.wi[P][0][1], b, c, sm));
        } else {
#pragma unroll
            for (int k = 1; k <= SL; ++k) {
                const double v = pair_value_h(w.wr[P][0][k], w.wi[P][0][k], cell.ld(0, I - k), cell.ld(0, I + k), sm);
                const double nt = __dadd_rn(t, v);
                t = ((w.flag[P][0] >> k) & 1u) ? nt : t;
            }
        }
        pair_bin_accumulate<Q, P, FOLD, PAT>(w, tv, t);
        // (3) project: |t| = sqrt(tr*tr + ti*ti), new value (t * a) / |t|  (lwslib.cpp:355-360); x + y == y + x bit for bit
        const double q = __dmul_rn(t, t);
        const double x = __dadd_rn(q, pair_swap(q, 0u));
        const double num = __dmul_rn(t, amp[I]);
        const bool act = (active >> I) & 1u;
        bool sok, rok, dok;
        double mag = fm_sqrt(x, sok);
        double val = fm_div(num, mag, fm_rcp(mag, rok), dok);
        if (__any_sync(0xffffffffu, act && !(x == 0.0) && !(sok && rok && dok))) { // rare: outside the fast ranges
            mag = __dsqrt_rn(x);
            val = __ddiv_rn(num, mag);
        }
        const bool ok = act && x > 0.0; // |t| > 0 (lwslib.cpp:356); sqrt(x) > 0 iff x > 0
        // (4) commit: own cell, its mirrored copy (lwslib.cpp:362-368), halo copies in the neighbouring strips
        double *row = reinterpret_cast<double *>(bc.ring + bc.ownoff) + h;
        const int col = cell.col0 + I;
        int mcol = col;
        if (bc.first_strip && n >= 1 && n <= SL) mcol = SL - n;
        else if (n >= bc.Nreal - 1 - SL && n <= bc.Nreal - 2) mcol = SL + 2 * (bc.Nreal - 1) - n - bc.b0;
        if (ok) {
            row[2 * col] = val;
            row[2 * mcol] = (mcol != col && h) ? -val : val;
            const int q = SBK * bc.xb + I; // bin inside the strip: the first / last L bins also live in a neighbour's halo
            if (q < SL && bc.ring_left)
                (reinterpret_cast<double *>(bc.ring_left + bc.ownoff) + h)[2 * (SL + SBK * bc.NBr + q)] = val;
            if (q >= SBK * bc.NBr - SL && bc.ring_right)
                (reinterpret_cast<double *>(bc.ring_right + bc.ownoff) + h)[2 * (q - (SBK * bc.NBr - SL))] = val;
        }
        if constexpr (PAT == 1) own.cur[I + 1] = ok ? val : own.cur[I + 1];
        if constexpr (I + 1 < SBK) {
            if constexpr (PIPE == 1) pair_block<Q, FOLD, PAT, WM, PIPE, I + 1>(cell, win, w, bc, amp, active, sm, h, tvn, own);
            else pair_block<Q, FOLD, PAT, WM, PIPE, I + 1>(cell, win, w, bc, amp, active, sm, h, tv, own);
        }
    }
}

template <int Q, int FOLD, int PAT, int WM, int PIPE>
__device__ __forceinline__ void pair_update_block(const PairCell<Q> &cell, const StripW<Q> &w, const BlockCtx &bc, const double *amp,
                                                  unsigned active, int h)
{
    const unsigned sm = h ? 0u : 0x80000000u;
    PairWin<Q, WM> win;
    pair_win_load_range<Q, WM, -SL, SL>(cell, win);
    PairOwn own;
    if constexpr (PAT == 1) {
#pragma unroll
        for (int i = 0; i < SBK + 2; ++i) own.cur[i] = cell.ld(0, i - 1);
    }
    PairTerms<Q, FOLD> tv;
    if constexpr (PIPE == 1) pair_bin_terms<Q, 0, FOLD, PAT, WM, 0>(cell, win, w, sm, tv);
    pair_block<Q, FOLD, PAT, WM, PIPE, 0>(cell, win, w, bc, amp, active, sm, h, tv, own);
}

