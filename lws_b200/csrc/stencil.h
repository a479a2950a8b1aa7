// stencil.h -- host-side construction of the linear update stencils.
//
// Every update rule of the reference C core is a linear combination of neighbouring cells of
// the extended spectrogram with coefficients W or conj(W):
//     term(w; b, c) = w*b + conj(w)*c            (lwslib.cpp:98-99)
// so for a given (weight set, formula family, rframe, cframe, residue p = bin mod Q) the sum
// can be tabulated once, at weight-upload time, as a list of (dr, dk, coef) terms.  Building
// the list per formula family (Q2 / Q4 / anyQ) -- and not from the algebraically "intended"
// stencil -- keeps the result identical to the reference for ANY weight array a caller hands
// in, including ones that do not have the symmetries the Q2/Q4 shortcuts assume.
// Terms whose weight fails the reference's |W| > 1e-12 mask (lws.pyx:231-232) are dropped,
// exactly like the `if (w_flag[...])` guards do.
#pragma once
#include <cmath>
#include <vector>
#include "lwsb_common.h"

namespace lwsb {

struct WeightSet {
    int Q = 0, L = 0;
    int Qp = 0; // rows: Q (summarised weights) or one per FFT bin (the reference's *fractionalQ variants)
    std::vector<double> wr, wi; // (Q, Q, L+1)
    bool valid() const { return Q > 0; }
    size_t idx(int p, int r, int k) const { return ((size_t)p * Q + r) * (L + 1) + k; }
    bool flag(int p, int r, int k) const { return std::hypot(wr[idx(p, r, k)], wi[idx(p, r, k)]) > 1.0e-12; }
};

// Appends the terms for residue p.  rframe / cframe as in Asym_UpdatePhase* (lwslib.cpp:1141-1151):
// batch sweep = (Q, 1) (lwslib.cpp:283-373); no-future sweep = (1, 0) (lwslib.cpp:620-690).
inline void build_terms(const WeightSet &w, int fold, int rframe, int cframe, int p, std::vector<LwsbTerm> &out)
{
    const int Q = w.Q, L = w.L;
    const int pn = (Q - p) % Q;
    auto push = [&](int dr, int dk, double cr, double ci) { out.push_back(LwsbTerm{dr, dk, cr, ci}); };

    if (fold == LWSB_FOLD_NF4) {
        // NoFuture_LWSQ4 as written (lwslib.cpp:557-598): offsets are relative to (m-r)*Np + 2n.
        for (int r = Q - 1; r > 0; --r) {
            const double s = ((p % 2 == 1) && (r % 2 == 1)) ? -1.0 : 1.0;
            for (int k = 1; k <= L; ++k)
                if (w.flag(p, r, k)) {
                    const double a = w.wr[w.idx(p, r, k)], b = w.wi[w.idx(p, r, k)];
                    push(-r, -k, a, b);
                    push(-r, +k, s * a, -s * b);
                }
            if (w.flag(p, r, 0)) push(-r, 0, w.wr[w.idx(p, r, 0)], w.wi[w.idx(p, r, 0)]);
        }
        return;
    }

    if (cframe)
        for (int k = 1; k <= L; ++k)
            if (w.flag(p, 0, k)) {
                const double a = w.wr[w.idx(p, 0, k)], b = w.wi[w.idx(p, 0, k)];
                push(0, -k, a, b);
                push(0, +k, a, -b);
            }
    for (int r = 1; r < Q; ++r) {
        const bool both = r < rframe;
        if (w.flag(p, r, 0)) {
            const double a = w.wr[w.idx(p, r, 0)], b = w.wi[w.idx(p, r, 0)];
            push(-r, 0, a, b);
            if (both) push(+r, 0, a, -b);
        }
        for (int k = 1; k <= L; ++k) {
            if (fold == LWSB_FOLD_ANY) {
                if (w.flag(p, r, k)) { // lwslib.cpp:333-342 / 1236-1243
                    const double a = w.wr[w.idx(p, r, k)], b = w.wi[w.idx(p, r, k)];
                    push(-r, -k, a, b);
                    if (both) push(+r, -k, a, -b);
                }
                if (w.flag(pn, r, k)) { // lwslib.cpp:343-352 / 1244-1251
                    const double a = w.wr[w.idx(pn, r, k)], b = w.wi[w.idx(pn, r, k)];
                    if (both) push(+r, +k, a, b);
                    push(-r, +k, a, -b);
                }
            } else if (w.flag(p, r, k)) { // Q2: lwslib.cpp:119-130; Q4: 200-211, 224-235, 248-259, 990-1001
                const double s = (fold == LWSB_FOLD_Q4 && (p % 2 == 1) && (r % 2 == 1)) ? -1.0 : 1.0;
                const double a = w.wr[w.idx(p, r, k)], b = w.wi[w.idx(p, r, k)];
                push(-r, -k, a, b);
                push(-r, +k, s * a, -s * b);
                if (both) {
                    push(+r, -k, a, -b);
                    push(+r, +k, s * a, s * b);
                }
            }
        }
    }
}

} // namespace lwsb
