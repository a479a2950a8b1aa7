// TEST INFRASTRUCTURE ONLY -- extern "C" trampolines onto the reference's own C core so that
// ctypes can call it without C++ name mangling (the reference compiles lwslib.cpp as C++ and
// declares no extern "C": /root/reference/lwslib/lwslib.h:6-26).  This file contains no
// algorithm; it is linked with /root/reference/lwslib/lwslib.cpp by oracle/Makefile (target
// `ref`) into oracle/_ref/liblws_ref.so.
#include "lwslib.h"

#define SWEEP_ARGS double *Sr, double *Si, double *wr, double *wi, int *wf, double *amp
extern "C" {
void ref_ExtendSpec(double *ESr, double *ESi, double *Sr, double *Si, int Nreal, int M, int L, int Q)
{ ExtendSpec(ESr, ESi, Sr, Si, Nreal, M, L, Q); }
void ref_CopySpec(double *ESr, double *ESi, double *Sr, double *Si, int Nreal, int M, int L, int Q)
{ CopySpec(ESr, ESi, Sr, Si, Nreal, M, L, Q); }
void ref_ComputeAmpSpec(double *Sr, double *Si, double *amp, int size) { ComputeAmpSpec(Sr, Si, amp, size); }

void ref_LWSQ2(SWEEP_ARGS, int Nreal, int M, int L, double thr) { LWSQ2(Sr, Si, wr, wi, wf, amp, Nreal, M, L, thr); }
void ref_LWSQ4(SWEEP_ARGS, int Nreal, int M, int L, double thr) { LWSQ4(Sr, Si, wr, wi, wf, amp, Nreal, M, L, thr); }
void ref_LWSanyQ(SWEEP_ARGS, int Nreal, int M, int L, int Q, double thr) { LWSanyQ(Sr, Si, wr, wi, wf, amp, Nreal, M, L, Q, thr); }
void ref_LWSfractionalQ(SWEEP_ARGS, int Nreal, int M, int L, int Q, double thr) { LWSfractionalQ(Sr, Si, wr, wi, wf, amp, Nreal, M, L, Q, thr); }

void ref_NoFuture_LWSQ2(SWEEP_ARGS, int Nreal, int M, int L, double thr) { NoFuture_LWSQ2(Sr, Si, wr, wi, wf, amp, Nreal, M, L, thr); }
void ref_NoFuture_LWSQ4(SWEEP_ARGS, int Nreal, int M, int L, double thr) { NoFuture_LWSQ4(Sr, Si, wr, wi, wf, amp, Nreal, M, L, thr); }
void ref_NoFuture_LWSanyQ(SWEEP_ARGS, int Nreal, int M, int L, int Q, double thr) { NoFuture_LWSanyQ(Sr, Si, wr, wi, wf, amp, Nreal, M, L, Q, thr); }

void ref_NoFuture_LWSfractionalQ(SWEEP_ARGS, int Nreal, int M, int L, int Q, double thr) { NoFuture_LWSfractionalQ(Sr, Si, wr, wi, wf, amp, Nreal, M, L, Q, thr); }
void ref_Asym_UpdatePhasefractionalQ(SWEEP_ARGS, int Nreal, int M, int M0, int L, int Q, double Qfloat, double thr, int update)
{ Asym_UpdatePhasefractionalQ(Sr, Si, wr, wi, wf, amp, Nreal, M, M0, L, Q, Qfloat, thr, update); }

void ref_Asym_UpdatePhaseQ2(SWEEP_ARGS, int Nreal, int M, int M0, int L, double thr, int update)
{ Asym_UpdatePhaseQ2(Sr, Si, wr, wi, wf, amp, Nreal, M, M0, L, thr, update); }
void ref_Asym_UpdatePhaseQ4(SWEEP_ARGS, int Nreal, int M, int M0, int L, double thr, int update)
{ Asym_UpdatePhaseQ4(Sr, Si, wr, wi, wf, amp, Nreal, M, M0, L, thr, update); }
void ref_Asym_UpdatePhaseanyQ(SWEEP_ARGS, int Nreal, int M, int M0, int L, int Q, double thr, int update)
{ Asym_UpdatePhaseanyQ(Sr, Si, wr, wi, wf, amp, Nreal, M, M0, L, Q, thr, update); }

void ref_TF_RTISI_LA(double *Sr, double *Si, double *wr, double *wi, double *wr_ai, double *wi_ai,
                     double *wr_af, double *wi_af, int *wf, int *wf_ai, int *wf_af, double *amp,
                     int iter, int LA, int Nreal, int M, int L, int Q, double Qfloat,
                     int use_summarized_weights, double *thresholds, int update)
{ TF_RTISI_LA(Sr, Si, wr, wi, wr_ai, wi_ai, wr_af, wi_af, wf, wf_ai, wf_af, amp, iter, LA, Nreal, M, L, Q,
              Qfloat, use_summarized_weights, thresholds, update); }
}
