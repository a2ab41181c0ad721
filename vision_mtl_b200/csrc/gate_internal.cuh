// Internal interface between gate.cu (entry points, streaming kernels, fp32 FFMA GEMM) and
// gate_tc.cu (tcgen05 / TMEM contraction kernels).
#pragma once

#include "vmtl_common.cuh"

namespace vmtl {

// Caller-provided workspace, carved identically by every gate entry point.
struct GateWs {
  float* coefA;  // [N] gamma*invstd
  float* coefB;  // [N] beta - mean*gamma*invstd
  float* c1;     // [N] dbeta/M   (train) or 0 (eval)
  float* c2;     // [N] dgamma/M  (train) or 0 (eval)
  float* mean;   // [N] statistics used for normalisation (batch or running)
  float* invstd; // [N]
  float* partial;       // [partial_rows][2N] per-block column partials
  int partial_rows;
  float* gemm_partial;  // [gemm_slots][N*K] split-K / per-CTA dW partials
  int gemm_slots;
  float* dz;            // [M,N] fp32 dz of the CUDA-core backward (not allocated for tensor-core shapes)
  float* hs_partial;    // [hs_rows][K] per-CTA column sums of h (tensor-core backward)
  int hs_rows;
  float* zbuf;          // [M,N] forward scratch for z when the caller passes no save_z and the shape is
                        // outside the tensor-core kernels (CUDA-core path, eval mode)
};

// shapes the tcgen05 kernels cover (everything else takes the CUDA-core path in gate.cu)
bool gate_tc_supported(int K, int N);

inline size_t gate_ws_floats(int64_t M, int K, int N, int precision, int backward, GateWs* ws,
                             float* base) {
  size_t off = 0;
  auto take = [&](size_t n) {
    float* p = base ? base + off : nullptr;
    off += (n + 63) / 64 * 64;  // keep every region 256-byte aligned
    return p;
  };
  GateWs w{};
  w.coefA = take(N);
  w.coefB = take(N);
  w.c1 = take(N);
  w.c2 = take(N);
  w.mean = take(N);
  w.invstd = take(N);
  w.partial_rows = sm_count() * 8;
  w.partial = take((size_t)w.partial_rows * 2 * N);
  w.gemm_slots = backward ? sm_count() * 2 : 0;
  w.gemm_partial = take((size_t)w.gemm_slots * N * K);
  const bool tc = precision != VMTL_GATE_FP32_FFMA && gate_tc_supported(K, N);
  w.dz = (backward && !tc) ? take((size_t)M * N) : nullptr;
  w.hs_rows = sm_count() + 16;  // grid of pass 1: <= sm_count, or the chunk count (<= 16) on a tiny device
  w.hs_partial = (backward && tc) ? take((size_t)w.hs_rows * K) : nullptr;
  w.zbuf = (!backward && !tc) ? take((size_t)M * N) : nullptr;
  if (ws) *ws = w;
  return off;
}

// tcgen05 kernels (gate_tc.cu).  Return VMTL_EUNSUPPORTED for shapes they do not cover.
// Forward phase 1: z = h @ W^T + bias -> save_z, per-CTA column partials (sum, sumsq) into
// ws.partial rows [0, *nparts).  EVAL variant applies the folded BN + sigmoid + product
// directly and writes y.
// h_coef (everywhere below): nullptr, or [A1 | B1] ([2][K]) -- then `h` is the hidden layer's PRE-activation and the
// kernels rebuild h = max(A1 c + B1, 0) while they convert the operand.
int gate_tc_fwd_gemm(const float* h, const float* h_coef, const float* W, const float* bias, int64_t M, int K, int N,
                     int split3, float* z_out, float* partial, int partial_rows, int* nparts,
                     cudaStream_t st);
int gate_tc_fwd_eval(const float* h, const float* h_coef, const float* s, const float* W, const float* bias,
                     const float* coefA, const float* coefB, int64_t M, int K, int N, int split3,
                     float* y, cudaStream_t st);
// Backward pass 1 (statistics + dW partials + ds).  CTA b of the *nparts CTAs owns the column chunk b % nch
// (nch = N <= 64 ? 1 : N / 64, chunk width Nc = N / nch); its partials land in ws.gemm_partial ([b][2][Nc][K]:
// P1 = du^T h, P2 = zhat^T h), ws.hs_partial ([b][K]: sum_r h over its units) and ws.partial ([b][3][Nc]: sum du,
// sum du*zhat, sum zhat); gate.cu's finalize turns them into the gradients.
int gate_tc_bwd_pass1(const float* dy, const float* h, const float* h_coef, const float* s, const float* z, const float* gamma,
                      const float* beta, const float* mean, const float* invstd, int64_t M, int K, int N,
                      int split3, float* ds, const GateWs& ws, int* nparts, cudaStream_t st);
// grid (= number of per-CTA partial rows) gate_tc_bwd_pass1 launches for (M, N)
int gate_tc_bwd_pass1_grid(int64_t M, int N);
// Backward pass 2: dh = dz @ W with dz rebuilt per tile from (dy, s, z) and the finalized ws.c1 / ws.c2.
int gate_tc_bwd_dh(const float* dy, const float* s, const float* z, const float* W, const GateWs& ws, int64_t M,
                   int K, int N, int split3, float* dh, cudaStream_t st);

}  // namespace vmtl
