// Host-side orchestration structs shared by api.cu (solver loops, PARALLEL / PARALLEL_L2 fields) and
// field_macaron.cu (MACARON field).  Not part of the ABI.
#pragma once
#include "internal.h"

namespace odevit {

struct Plan {
  int B, N, D, H, hid, d, M;
  int variant, precision, act;  // act = DType of activation buffers
  int dd_type;                  // DType of the cotangent buffer entering a field VJP
  float scaler;
  long long BHNN;
  // training-mode dropout (0 = off) and the per-call seed of the mask generator
  float p_attn, p_proj, p_mlp;
  uint32_t seed_lo, seed_hi;
  const uint32_t* seed_dev;   // device-resident seed (or null): keys come from `drop_keys`
  uint32_t* drop_keys;        // [kDropKeyEvals * DS_SITES] key table in the workspace (device seed only)
  bool any_drop;
  bool split_out;   // proj / mlp dropout on: out-proj and fc2 run as two GEMMs (their outputs take different masks)
  // fp32 mode on the tensor core: scratch for the bf16 hi / lo split of a GEMM's operands (api.cu::gemm_split3); set by
  // the workspace layouts (null: the FFMA kernel runs)
  mutable void* split_a = nullptr;
  mutable void* split_b = nullptr;
  mutable size_t split_a_bytes = 0, split_b_bytes = 0;
};
// mask of dropout site `site` in field evaluation `e` (e = step * stages + stage)
Drop make_drop(const Plan& p, int site, long long e);

struct Arena {
  char* base;
  size_t off = 0;
  explicit Arena(void* b) : base(reinterpret_cast<char*>(b)) {}
  void* take(size_t bytes) {
    off = (off + 1023) & ~size_t(1023);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
  float* f32(size_t n) { return reinterpret_cast<float*>(take(n * 4)); }
};

// Activation-typed copies of the weights, laid out the same way for every variant:
//   w1cat  [3D+hid, D]  = [in-proj rows (q rows carry 1/sqrt(d)) ; fc1 rows]     (+ transpose [D, 3D+hid])
//   w2cat  [D, D+hid]   = [out-proj | fc2] along K                              (+ transpose [D+hid, D])
// PARALLEL / PARALLEL_L2 fold the CenterNorm affines into w1cat and run each concatenation as ONE
// GEMM; MACARON uses the four blocks as separate operands (views with the concatenated leading dim).
struct WeightBufs {
  void *w1cat, *w1catT, *w2cat, *w2catT;
  float *b1cat, *b2;
  float* fmean;   // [3D+hid] row means of W1cat (what the transposed copy subtracts)
  const odevit_weights* user;  // the caller's fp32 parameters (biases, LayerNorm affines, res_scale)
};

// Intermediates of ONE field evaluation that its VJP needs (a tape slot or a recompute buffer).
struct StageCtx {
  void *xc, *qkv, *oh, *hpre;
  float* lse;  // [B,H,N] log2-domain row log-sum-exp of the attention (reverse sweep only)
  // MACARON only: xc = LN1(u), hpre/oh[:, D:] = first half-FFN; the rest of the chain:
  float *x0, *x1, *x2;   // [M,D] fp32: the stage input, after the first half-FFN, after attention
  void *n2, *n3;         // [M,D] LN2(x1), LN3(x2)
  void *hpre3, *h3;      // [M,hid] second half-FFN
};

struct BwdBufs {
  uint32_t* drop_keys;   // device-seeded dropout: key table (head of the workspace) or null
  WeightBufs w;
  StageCtx ctx[4];
  float *P, *dP;
  float* u;
  float* k[3];
  void *dd, *dO, *dz;
  float* mu[4];
  float* gy;
  float *delta, *dq_scratch;
  float *G1, *c1, *G2, *c2, *c3;
  size_t acc_bytes;  // G1..c3 are contiguous: one memset
  // MACARON / L2 scratch
  float* dn;    // [M,D] fp32: cotangent of a LayerNorm output
  void* ddc;    // [M,D] act: cast/scaled copy of the running cotangent (GEMM operand)
  float* sq;    // [2,B,H,N] fp32: |q|^2, |k|^2 per head (L2 attention)
  float* tmp;   // [M,D] fp32: masked fc2 output waiting for the out-proj GEMM (split_out)
  void *dd1, *dd2;  // [M,D] act: the cotangent under the fc2-output / out-proj-output masks (split_out)
};

// GEMM dispatch: tcgen05 in bf16 mode where the kernel covers the problem, FFMA otherwise.
int gemm(const Plan& p, const GemmArgs& g, cudaStream_t s);

// softmax(q k^T) v per (image, head) from the packed qkv buffer into oh[:, h*d ...] (leading dim ld_oh).
// P: [B,H,N,N] fp32 scratch (unfused path); p_copy: optional export; lse: optional row log-sum-exp.
int attention_forward(const Plan& p, const void* qkv, void* oh, long long ld_oh, float* P, float* p_copy,
                      float* lse, float* sq, Drop drop, cudaStream_t s,
                      float* jas_out = nullptr, int jas_k = 0);
// Its VJP: dO [M,D] (act) -> dq|dk|dv into dz (leading dim R, act).  g_p: optional cotangent of P.
// dq_colsum / colsum_done: the fused kernel can add the column sums of dq (the q-row bias gradient) into dq_colsum [D]
// on its way out; *colsum_done says whether it did.
int attention_vjp(const Plan& p, const void* qkv, const void* oh, long long ld_oh, const float* lse, BwdBufs& b,
                  const float* g_p, void* dz, int R, Drop drop, cudaStream_t s, float* dq_colsum = nullptr,
                  bool* colsum_done = nullptr);

// On-chip-state solver for small-token shapes (solve_resident.cu): one persistent CTA per image runs
// every step of the solve with the ODE state in shared memory.  Inference, PARALLEL field, bf16 mode.
bool solve_resident_shape_ok(const Plan& p);
bool solve_resident_supports(const Plan& p, int n_grid, bool wants_p_traj, bool has_tape);
size_t solve_resident_scratch_floats(const Plan& p);
int solve_resident(const Plan& p, const WeightBufs& wb, int S, const float (*ta)[4], const float* tbv, const float* x0,
                   const float* t_host, int n_grid, float* states, float* final_state, float* p_last, float* kbuf,
                   cudaStream_t s);

// MACARON field (field_macaron.cu)
int macaron_forward(const Plan& p, const WeightBufs& wb, const StageCtx& c, const float* u, float* P,
                    const Epi* rk, long long e, cudaStream_t s);
int macaron_vjp(const Plan& p, const WeightBufs& wb, const StageCtx& c, BwdBufs& b,
                const odevit_weight_grads* gw, const Epi& mu_epi, long long e, cudaStream_t s);

// Pre-LN encoder stack, forward only (field_macaron.cu): the distillation teacher
int encoder_forward(const Plan& p, const WeightBufs* layers, int n_layers, float ln_eps, const float* x0, float* hidden,
                    float* p_out, int p_mode, const StageCtx& c, float* P_scratch, cudaStream_t s);

}  // namespace odevit
